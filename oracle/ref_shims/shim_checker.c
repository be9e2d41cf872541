/* shim over reference src/nodes/textures/checker.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/textures/checker.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_checker(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct checkerTexture *t = (const struct checkerTexture *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_CHECKER, .in = { t->A, t->B, t->scale } };
	return true;
}
