/* shim over reference src/nodes/input/fresnel.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/input/fresnel.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_fresnel(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct fresnelNode *t = (const struct fresnelNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_FRESNEL, .in = { t->IOR, t->normal } };
	return true;
}
