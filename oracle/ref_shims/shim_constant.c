/* shim over reference src/nodes/textures/constant.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/textures/constant.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_constant_color(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct constantTexture *t = (const struct constantTexture *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_CONSTANT, .f = { t->color.red, t->color.green, t->color.blue, t->color.alpha } };
	return true;
}
