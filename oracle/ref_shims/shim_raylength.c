/* shim over reference src/nodes/input/raylength.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/input/raylength.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_raylength(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_RAYLENGTH };
	return true;
}
