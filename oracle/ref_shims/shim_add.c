/* shim over reference src/nodes/shaders/add.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/shaders/add.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_add(const struct bsdfNode *n, struct crx_nodeinfo *o) {
	if (n->sample != sample) return false;
	const struct addBsdf *t = (const struct addBsdf *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_BSDF_ADD, .in = { t->A, t->B } };
	return true;
}
