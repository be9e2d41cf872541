/* shim over reference src/nodes/shaders/emission.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/emission.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_emission(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct emissiveBsdf *t = (const struct emissiveBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_EMISSIVE, .in = { t->color, t->strength } };
	return true;
}
