/* shim over reference src/nodes/converter/combinergb.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/combinergb.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_combine_rgb(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct combineRGB *t = (const struct combineRGB *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_COMBINE_RGB, .in = { t->R, t->G, t->B } };
	return true;
}
