/* shim over reference src/nodes/input/normal.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/input/normal.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_normal(const struct vectorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	*o = (struct crx_nodeinfo){ .kind = CRS_VECTOR_NORMAL };
	return true;
}
