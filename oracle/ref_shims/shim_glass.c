/* shim over reference src/nodes/shaders/glass.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/glass.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_glass(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct glassBsdf *t = (const struct glassBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_GLASS, .in = { t->color, t->roughness, t->IOR } };
	return true;
}
