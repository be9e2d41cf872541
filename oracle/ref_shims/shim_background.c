/* shim over reference src/nodes/shaders/background.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/background.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_background(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct backgroundBsdf *t = (const struct backgroundBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_BACKGROUND, .in = { t->color, t->strength, t->offset } };
	return true;
}
