/* shim over reference src/nodes/textures/alpha.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/textures/alpha.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_alpha(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct alphaNode *t = (const struct alphaNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_ALPHA, .in = { t->color } };
	return true;
}
