/* shim over reference src/nodes/converter/blackbody.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/blackbody.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_blackbody(const struct colorNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct blackbodyNode *t = (const struct blackbodyNode *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_COLOR_BLACKBODY, .in = { t->temperature } };
	return true;
}
