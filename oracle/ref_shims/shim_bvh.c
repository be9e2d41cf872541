/* shim over reference src/accelerators/bvh.c — see crx.h. The reference file is compiled in place, unmodified. */
#include "accelerators/bvh.c"
#include <string.h>
#include "crx.h"
#include "../../include/crscene.h"

unsigned crx_bvh_node_count(const struct bvh *b) { return b ? b->nodeCount : 0; }
const void *crx_bvh_nodes(const struct bvh *b) { return b ? b->nodes : NULL; }
const int *crx_bvh_prim_indices(const struct bvh *b) { return b ? b->primIndices : NULL; }
_Static_assert(sizeof(struct bvhNode) == 32, "bvhNode layout");
/* returns the last 32-bit word of a node with primCount=5, isLeaf=true: must be (5 | 1<<30) */
unsigned crx_bvh_layout_probe(void) {
	struct bvhNode n; memset(&n, 0, sizeof n);
	n.primCount = 5; n.isLeaf = true;
	unsigned w; memcpy(&w, (const char *)&n + 28, 4);
	return w;
}
