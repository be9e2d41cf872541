/* shim over reference src/nodes/shaders/isotropic.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/isotropic.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_isotropic(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct isotropicBsdf *t = (const struct isotropicBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_ISOTROPIC, .in = { t->color } };
	return true;
}
