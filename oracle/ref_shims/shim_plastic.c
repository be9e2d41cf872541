/* shim over reference src/nodes/shaders/plastic.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/plastic.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_plastic(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct plasticBsdf *t = (const struct plasticBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_PLASTIC, .in = { t->color, t->roughness, t->diffuse } };
	return true;
}
