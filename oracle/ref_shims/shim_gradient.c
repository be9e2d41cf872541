/* shim over reference src/nodes/textures/gradient.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/textures/gradient.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_gradient(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct gradientTexture *t = (const struct gradientTexture *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_GRADIENT, .f = { t->down.red, t->down.green, t->down.blue, t->down.alpha, t->up.red, t->up.green, t->up.blue, t->up.alpha } };
	return true;
}
