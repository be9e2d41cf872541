/* shim over reference src/nodes/textures/image.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/textures/image.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_image(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct imageTexture *t = (const struct imageTexture *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_IMAGE, .tex = t->tex, .options = t->options };
	return true;
}
_Static_assert(SRGB_TRANSFORM == CRS_IMG_SRGB_TRANSFORM && NO_BILINEAR == CRS_IMG_NO_BILINEAR, "image options");
