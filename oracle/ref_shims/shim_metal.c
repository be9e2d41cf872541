/* shim over reference src/nodes/shaders/metal.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/metal.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_metal(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct metalBsdf *t = (const struct metalBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_METAL, .in = { t->color, t->roughness } };
	return true;
}
