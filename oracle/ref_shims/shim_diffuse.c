/* shim over reference src/nodes/shaders/diffuse.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/diffuse.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_diffuse(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct diffuseBsdf *t = (const struct diffuseBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_DIFFUSE, .in = { t->color } };
	return true;
}
