/* shim over reference src/nodes/shaders/transparent.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/transparent.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_transparent(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct transparent *t = (const struct transparent *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_TRANSPARENT, .in = { t->color } };
	return true;
}
