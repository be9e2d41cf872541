/* shim over reference src/nodes/valuenode.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "nodes/valuenode.c"
#include "crx.h"
#include "../../include/crscene.h"

bool crx_is_constant_value(const struct valueNode *n, struct crx_nodeinfo *o) {
	if (n->eval != eval) return false;
	const struct constantValue *t = (const struct constantValue *)n;
	*o = (struct crx_nodeinfo){ .kind = CRS_VALUE_CONSTANT, .f = { t->value } };
	return true;
}
