/* shim over reference src/nodes/converter/combine.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/combine.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_combine_value(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct combineValue *t = (const struct combineValue *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_COMBINE_VALUE, .in = { t->input } };
	return true;
}
