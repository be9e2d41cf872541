/*
 * cr_bvh_build.c — binned-SAH BVH construction for the scene loader (host side, runs once per scene).
 *
 * The GPU traversal kernels consume exactly the tree the reference builds (node order, leaf ranges and
 * primitive order decide which triangle wins a tie), so this builder follows the algorithm of reference
 * src/accelerators/bvh.c:96-296 decision for decision: 32 bins per axis, right-to-left then left-to-right
 * sweeps, cheapest axis with ties going to the lower axis, leaf cost area*(n - 1.5), approximate-median
 * fallback above 16 primitives, in-place two-pointer partition, children allocated as a pair, depth cap 64.
 * Float expressions keep the reference's association; float->unsigned conversions follow x86-64 cvttss2si.
 * The recursion is replaced by an explicit work stack (same visiting order: left subtree first).
 */
#include "cr_loader_int.h"
#include <stdlib.h>
#include <string.h>

#define MAX_DEPTH      64
#define MAX_LEAF_SIZE  16
#define TRAVERSAL_COST 1.5f
#define BIN_COUNT      32

struct bin { bbox3 bbox; unsigned count; float cost; };

/* (unsigned)f as GCC compiles it on x86-64: cvttss2si to 64 bits, low 32 bits kept; NaN/overflow give 0 */
static inline unsigned f2u(float f) {
	if (!(f > -9.2233720e18f && f < 9.2233720e18f)) return 0u;
	return (unsigned)(long long)f;
}

static inline unsigned bin_index(int axis, const vec3 *center, float lo, float hi) {    /* bvh.c:89-95 */
	float centerToBin = BIN_COUNT / (hi - lo);
	float coord = axis == 0 ? center->x : (axis == 1 ? center->y : center->z);
	float floatIndex = (coord - lo) * centerToBin;
	unsigned b = f2u(floatIndex < 0 ? 0 : floatIndex);
	return b >= BIN_COUNT ? BIN_COUNT - 1 : b;
}

static inline void store_bbox(struct crs_bvh_node *n, const bbox3 *b) {
	n->bounds[0] = b->min.x; n->bounds[1] = b->max.x;
	n->bounds[2] = b->min.y; n->bounds[3] = b->max.y;
	n->bounds[4] = b->min.z; n->bounds[5] = b->max.z;
}

static inline void make_leaf(struct crs_bvh_node *n, unsigned begin, unsigned count) {
	n->first_child_or_prim = begin;
	n->prim_count_leaf = CRS_BVH_LEAF_BIT | (count & CRS_BVH_COUNT_MASK);
}

struct work { unsigned node, begin, end, depth; };

int crl_build_bvh(void *user, crl_bbox_fn fn, unsigned count,
                  struct crs_bvh_node **out_nodes, uint32_t *out_count, int32_t **out_prims) {
	*out_nodes = NULL; *out_count = 0; *out_prims = NULL;
	if (count < 1) return 0;
	vec3 *centers = malloc(sizeof(vec3) * count);
	bbox3 *bboxes = malloc(sizeof(bbox3) * count);
	int32_t *prims = malloc(sizeof(int32_t) * count);
	struct crs_bvh_node *nodes = calloc((size_t)2 * count - 1, sizeof(*nodes));
	struct work *stack = malloc(sizeof(struct work) * (MAX_DEPTH + 2));
	if (!centers || !bboxes || !prims || !nodes || !stack) { free(centers); free(bboxes); free(prims); free(nodes); free(stack); return -1; }

	bbox3 root = crl_empty_bbox;
	for (unsigned i = 0; i < count; ++i) {
		fn(user, i, &bboxes[i], &centers[i]);
		prims[i] = (int32_t)i;
		root.min = v_min(root.min, bboxes[i].min);
		root.max = v_max(root.max, bboxes[i].max);
	}
	unsigned node_count = 1;
	store_bbox(&nodes[0], &root);

	static _Thread_local struct bin bins[3][BIN_COUNT];
	int sp = 0;
	stack[sp++] = (struct work){ 0, 0, count, 0 };
	while (sp) {
		struct work w = stack[--sp];
		struct crs_bvh_node *node = &nodes[w.node];
		unsigned begin = w.begin, end = w.end, primCount = end - begin;
		if (w.depth >= MAX_DEPTH || primCount < 2) { make_leaf(node, begin, primCount); continue; }

		float minCost[3] = { FLT_MAX, FLT_MAX, FLT_MAX };
		unsigned minBin[3] = { 1, 1, 1 };
		for (int axis = 0; axis < 3; ++axis) {
			for (int i = 0; i < BIN_COUNT; ++i) { bins[axis][i].bbox = crl_empty_bbox; bins[axis][i].count = 0; }
			for (unsigned i = begin; i < end; ++i) {
				int p = prims[i];
				struct bin *b = &bins[axis][bin_index(axis, &centers[p], node->bounds[axis * 2], node->bounds[axis * 2 + 1])];
				bbox_extend(&b->bbox, &bboxes[p]);
				b->count++;
			}
			bbox3 cur = crl_empty_bbox;
			unsigned curCount = 0;
			for (unsigned i = BIN_COUNT; i > 1; --i) {            /* cost of everything to the right of a split */
				struct bin *b = &bins[axis][i - 1];
				curCount += b->count;
				bbox_extend(&cur, &b->bbox);
				b->cost = curCount * bbox_half_area(&cur);
			}
			cur = crl_empty_bbox;
			curCount = 0;
			for (unsigned i = 0; i < BIN_COUNT - 1; i++) {
				struct bin *b = &bins[axis][i];
				curCount += b->count;
				bbox_extend(&cur, &b->bbox);
				float cost = curCount * bbox_half_area(&cur) + bins[axis][i + 1].cost;
				if (cost < minCost[axis]) { minBin[axis] = i + 1; minCost[axis] = cost; }
			}
		}
		unsigned minAxis = 0;
		if (minCost[1] < minCost[0]) minAxis = 1;
		if (minCost[2] < minCost[minAxis]) minAxis = 2;

		bbox3 nb = { { node->bounds[0], node->bounds[2], node->bounds[4] }, { node->bounds[1], node->bounds[3], node->bounds[5] } };
		float leafCost = bbox_half_area(&nb) * (primCount - TRAVERSAL_COST);
		if (minCost[minAxis] > leafCost) {
			if (primCount > MAX_LEAF_SIZE) {
				unsigned accum = 0, best = primCount;
				for (unsigned i = 0; i < BIN_COUNT - 1; ++i) {
					accum += bins[minAxis][i].count;
					unsigned approx = (unsigned)abs((int)primCount / 2 - (int)accum);
					if (approx < best) { best = approx; minBin[minAxis] = i + 1; }
				}
			} else {
				make_leaf(node, begin, primCount);
				continue;
			}
		}

		/* two-pointer partition: everything in a bin below the split goes left (bvh.c:97-135) */
		const unsigned split = minBin[minAxis];
		const float lo = node->bounds[minAxis * 2], hi = node->bounds[minAxis * 2 + 1];
		unsigned i = begin, j = end;
		while (i < j) {
			while (i < j && bin_index((int)minAxis, &centers[prims[i]], lo, hi) < split) i++;
			while (i < j && bin_index((int)minAxis, &centers[prims[j - 1]], lo, hi) >= split) j--;
			if (i >= j) break;
			int32_t tmp = prims[j - 1]; prims[j - 1] = prims[i]; prims[i] = tmp;
			j--; i++;
		}
		const unsigned beginRight = i;
		if (beginRight > begin) {
			unsigned left = node_count, right = left + 1;
			node_count += 2;
			bbox3 lb = crl_empty_bbox, rb = crl_empty_bbox;
			for (unsigned k = 0; k < split; ++k) bbox_extend(&lb, &bins[minAxis][k].bbox);
			for (unsigned k = split; k < BIN_COUNT; ++k) bbox_extend(&rb, &bins[minAxis][k].bbox);
			store_bbox(&nodes[left], &lb);
			store_bbox(&nodes[right], &rb);
			node->first_child_or_prim = left;
			node->prim_count_leaf = 0;
			/* the reference recurses left first, so the left subtree claims node indices first: push right, then left */
			stack[sp++] = (struct work){ right, beginRight, end, w.depth + 1 };
			stack[sp++] = (struct work){ left, begin, beginRight, w.depth + 1 };
		} else {
			make_leaf(node, begin, primCount);
		}
	}
	free(centers); free(bboxes); free(stack);
	*out_nodes = realloc(nodes, sizeof(*nodes) * node_count);
	*out_count = node_count;
	*out_prims = prims;
	return 0;
}
