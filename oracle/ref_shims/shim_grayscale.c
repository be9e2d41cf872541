/* shim over reference src/nodes/converter/grayscale.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/converter/grayscale.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_grayscale(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct grayscale *t = (const struct grayscale *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_GRAYSCALE, .in = { t->input } };
	return true;
}
