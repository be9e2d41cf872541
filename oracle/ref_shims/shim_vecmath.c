/* shim over reference src/nodes/converter/vecmath.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/vecmath.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_vecmath(const struct vectorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct vecMathNode *t = (const struct vecMathNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VECTOR_VECMATH, .in = { t->A, t->B }, .options = (unsigned)t->op };
	return true;
}
