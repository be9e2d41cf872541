/* shim over reference src/nodes/converter/math.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/math.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_math(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct mathNode *t = (const struct mathNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_MATH, .in = { t->A, t->B }, .options = (unsigned)t->op };
	return true;
}
