/* shim over reference src/nodes/converter/vectocolor.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/vectocolor.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_vectocolor(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct vecToColorNode *t = (const struct vecToColorNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_VECTOCOLOR, .in = { t->vec } };
	return true;
}
