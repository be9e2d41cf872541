/* shim over reference src/nodes/shaders/mix.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/mix.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_mix(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct mixBsdf *t = (const struct mixBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_MIX, .in = { t->A, t->B, t->factor } };
	return true;
}
