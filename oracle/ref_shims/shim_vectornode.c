/* shim over reference src/nodes/vectornode.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/vectornode.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_constant_vector(const struct vectorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct constantVector *t = (const struct constantVector *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VECTOR_CONSTANT, .f = { t->vector.x, t->vector.y, t->vector.z } };
	return true;
}
