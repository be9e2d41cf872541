/*
 * crgpu_bvh_build.cu — SURVEY §8(f1): the reference's binned-SAH BVH build on the device.
 *
 * Restates reference src/accelerators/bvh.c:96-296 (buildBvhRecursive + partitionPrimitives) so that the result is THE SAME tree
 * the reference (and host/loader/cr_bvh_build.c) builds — node for node, leaf range for leaf range, primitive order included —
 * because the traversal's tie-breaking depends on all three.  The recursion becomes a level-synchronous sweep: every open node of a
 * level is binned, split and partitioned by the same few kernels, whatever its size.
 *
 *   bins        the reference folds `bbox_extend` over a node's primitives in array order with the min/max MACROS (bvh.c:158-171);
 *               ties return the second operand, so the only order dependence is the SIGN OF ZERO of a bound, decided by the last
 *               tied primitive.  Here every bin bound is a 64-bit atomic min/max over (ordered value with -0 == +0, array position):
 *               the winner is the extreme value and, among ties, the last position; the float is then read back from that primitive.
 *   split       one thread per open node runs the two 32-bin sweeps of the three axes, the leaf-cost test and the approximate
 *               median fallback with the reference's float association (explicit __f*_rn: no contraction).
 *   partition   the reference's two-pointer in-place partition (bvh.c:97-135) swaps the k-th misplaced element from the left
 *               with the k-th misplaced element from the right; the ranks come from two prefix sums over the whole array.
 *   numbering   children are allocated as a pair when their parent is visited, left subtree first (bvh.c:221-239), i.e. the pair
 *               of internal node X sits at 1 + 2 * (number of internal nodes before X in left-first preorder): one bottom-up pass
 *               for subtree sizes, one top-down pass for the ranks.
 *
 * Float->unsigned conversions follow x86-64 cvttss2si like the host builder (f2u).  NaN coordinates are not supported (the
 * reference's own comparisons make the tree undefined there).  Limits: 4,194,304 primitives per call (two-level prefix sum).
 */
#include "../../include/crgpu.h"
#include "../../include/crscene.h"
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <vector>

int crg_fail(int code, const char *fmt, ...);           /* crgpu_api.cu: sets crgpu_last_error() */

#define BVB_BINS 32
#define BVB_MAX_DEPTH 64u
#define BVB_MAX_LEAF 16u
#define BVB_SCAN_BLOCK 1024u     /* elements per scan block: 256 threads x 4 */
#define BVB_FLT_MAX 3.402823466e+38f

namespace {

struct Rec {                     /* one per tree node, in creation order (levels are contiguous) */
	uint32_t begin, end, depth;
	int32_t  left;               /* record of the left child (right = left + 1); -1 leaf; -2 open (to be decided) */
	float    b[6];               /* minx,maxx,miny,maxy,minz,maxz */
	uint32_t axis, split, mid;   /* split decision of an internal node; mid = first position of the right child */
	uint32_t slot;               /* index among the open nodes of its level */
	uint32_t internals, rank, idx;
	float    lb[6], rb[6];       /* children's bounds (bvh.c:210-217) */
};

struct Bins {                    /* per open node of the current level */
	unsigned           cnt[3][BVB_BINS];
	unsigned long long kmin[3][BVB_BINS][3];
	unsigned long long kmax[3][BVB_BINS][3];
};

__device__ __forceinline__ uint32_t okey(float f) {
	f = __fadd_rn(f, 0.0f);                              /* -0 -> +0: equal values must have equal keys */
	const uint32_t b = __float_as_uint(f);
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ unsigned long long key_max(float f, uint32_t pos) { return ((unsigned long long)okey(f) << 32) | pos; }
__device__ __forceinline__ unsigned long long key_min(float f, uint32_t pos) { return ((unsigned long long)okey(f) << 32) | (uint32_t)~pos; }
#define KMIN_EMPTY 0xffffffffffffffffull
#define KMAX_EMPTY 0ull

/* (unsigned)f as GCC compiles it on x86-64 (cvttss2si to 64 bits, low half kept), cr_bvh_build.c f2u */
__device__ __forceinline__ unsigned f2u(float f) {
	if (!(f > -9.2233720e18f && f < 9.2233720e18f)) return 0u;
	return (unsigned)(unsigned long long)__float2ll_rz(f);
}
__device__ __forceinline__ unsigned bin_index(float coord, float lo, float scale) {         /* bvh.c:89-95 */
	const float fi = __fmul_rn(__fsub_rn(coord, lo), scale);
	const unsigned b = f2u(fi < 0 ? 0 : fi);
	return b >= BVB_BINS ? BVB_BINS - 1 : b;
}
__device__ __forceinline__ float bin_scale(const float *b, int axis) { return __fdiv_rn((float)BVB_BINS, __fsub_rn(b[2 * axis + 1], b[2 * axis])); }
__device__ __forceinline__ float half_area(const float *mn, const float *mx) {               /* bbox.h:26 */
	const float ex = __fsub_rn(mx[0], mn[0]), ey = __fsub_rn(mx[1], mn[1]), ez = __fsub_rn(mx[2], mn[2]);
	return __fadd_rn(__fmul_rn(ex, __fadd_rn(ey, ez)), __fmul_rn(ey, ez));
}
#define MACRO_MIN(a, b) ((a) < (b) ? (a) : (b))
#define MACRO_MAX(a, b) ((a) > (b) ? (a) : (b))

/* ---- root bounds: the fold of bvh.c:266-271 over all primitives in order ------------------------------------------------------- */
__global__ void k_root_keys(const float *__restrict__ bb, uint32_t n, unsigned long long *keys /* [6]: min xyz, max xyz */) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	for (int k = 0; k < 3; ++k) {
		atomicMin(&keys[k], key_min(bb[6 * (size_t)i + k], i));
		atomicMax(&keys[3 + k], key_max(bb[6 * (size_t)i + 3 + k], i));
	}
}
__global__ void k_root_init(const float *__restrict__ bb, uint32_t n, const unsigned long long *keys, Rec *recs, uint32_t *counters) {
	Rec r;
	memset(&r, 0, sizeof r);
	r.begin = 0u; r.end = n; r.depth = 0u; r.slot = 0u;
	for (int k = 0; k < 3; ++k) {
		const uint32_t pmin = ~(uint32_t)(keys[k] & 0xffffffffull), pmax = (uint32_t)(keys[3 + k] & 0xffffffffull);
		r.b[2 * k] = bb[6 * (size_t)pmin + k];
		r.b[2 * k + 1] = bb[6 * (size_t)pmax + 3 + k];
	}
	const bool open = n >= 2u;
	r.left = open ? -2 : -1;
	recs[0] = r;
	counters[0] = 1u;                  /* records allocated */
	counters[1] = open ? 1u : 0u;      /* open nodes of the current level */
	counters[2] = 0u;                  /* open nodes of the next level */
}

/* ---- bins (bvh.c:158-171) ------------------------------------------------------------------------------------------------------- */
__global__ void k_bin(const float *__restrict__ bb, const float *__restrict__ ctr, const int32_t *__restrict__ prims,
					  const int32_t *__restrict__ owner, const Rec *__restrict__ recs, Bins *bins, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n) return;
	const int32_t r = owner[pos];
	if (r < 0) return;
	const Rec &rec = recs[r];
	Bins &B = bins[rec.slot];
	const uint32_t p = (uint32_t)prims[pos];
	const float *pb = bb + 6 * (size_t)p;
	for (int axis = 0; axis < 3; ++axis) {
		const unsigned bi = bin_index(ctr[3 * (size_t)p + axis], rec.b[2 * axis], bin_scale(rec.b, axis));
		atomicAdd(&B.cnt[axis][bi], 1u);
		for (int k = 0; k < 3; ++k) {
			atomicMin(&B.kmin[axis][bi][k], key_min(pb[k], pos));
			atomicMax(&B.kmax[axis][bi][k], key_max(pb[3 + k], pos));
		}
	}
}
__global__ void k_bins_clear(Bins *bins, uint32_t count) {
	const size_t words = (size_t)count * (sizeof(Bins) / 8u);
	unsigned long long *w = reinterpret_cast<unsigned long long *>(bins);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
		const size_t in = i % (sizeof(Bins) / 8u);
		const size_t cnt_words = sizeof(unsigned) * 3 * BVB_BINS / 8u, key_words = 3u * BVB_BINS * 3u;
		w[i] = in < cnt_words ? 0ull : (in < cnt_words + key_words ? KMIN_EMPTY : KMAX_EMPTY);
	}
}

/* ---- split decision: bvh.c:137-207, one thread per open node --------------------------------------------------------------------- */
__global__ void k_split(const float *__restrict__ bb, const int32_t *__restrict__ prims, Rec *recs, const uint32_t *__restrict__ open_list,
						uint32_t nopen, const Bins *__restrict__ bins) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nopen) return;
	Rec &rec = recs[open_list[s]];
	const Bins &B = bins[s];
	const uint32_t primCount = rec.end - rec.begin;
	float bmin[BVB_BINS][3], bmax[BVB_BINS][3], bcost[BVB_BINS];
	float minCost[3] = { BVB_FLT_MAX, BVB_FLT_MAX, BVB_FLT_MAX };
	unsigned minBin[3] = { 1u, 1u, 1u };
	float bestLb[6], bestRb[6];
	unsigned minAxis = 0u;
	/* the sweeps of the three axes; the bins of the winning axis are needed again for the children's boxes, so the axis loop runs
	 * once to find the winner and the winner's bins are resolved a second time below (cheaper than keeping 3 x 32 boxes) */
	for (int pass = 0; pass < 2; ++pass) {
		for (int axis = 0; axis < 3; ++axis) {
			if (pass == 1 && (unsigned)axis != minAxis) continue;
			for (int i = 0; i < BVB_BINS; ++i)
				for (int k = 0; k < 3; ++k) {
					const unsigned long long a = B.kmin[axis][i][k], c = B.kmax[axis][i][k];
					bmin[i][k] = a == KMIN_EMPTY ? BVB_FLT_MAX : bb[6 * (size_t)prims[~(uint32_t)(a & 0xffffffffull)] + k];
					bmax[i][k] = c == KMAX_EMPTY ? -BVB_FLT_MAX : bb[6 * (size_t)prims[(uint32_t)(c & 0xffffffffull)] + 3 + k];
				}
			if (pass == 1) break;
			float cmin[3] = { BVB_FLT_MAX, BVB_FLT_MAX, BVB_FLT_MAX }, cmax[3] = { -BVB_FLT_MAX, -BVB_FLT_MAX, -BVB_FLT_MAX };
			unsigned curCount = 0u;
			for (unsigned i = BVB_BINS; i > 1; --i) {                  /* cost of everything to the right of a split */
				curCount += B.cnt[axis][i - 1];
				for (int k = 0; k < 3; ++k) { cmin[k] = MACRO_MIN(cmin[k], bmin[i - 1][k]); cmax[k] = MACRO_MAX(cmax[k], bmax[i - 1][k]); }
				bcost[i - 1] = __fmul_rn(__uint2float_rn(curCount), half_area(cmin, cmax));
			}
			for (int k = 0; k < 3; ++k) { cmin[k] = BVB_FLT_MAX; cmax[k] = -BVB_FLT_MAX; }
			curCount = 0u;
			for (unsigned i = 0; i < BVB_BINS - 1; ++i) {
				curCount += B.cnt[axis][i];
				for (int k = 0; k < 3; ++k) { cmin[k] = MACRO_MIN(cmin[k], bmin[i][k]); cmax[k] = MACRO_MAX(cmax[k], bmax[i][k]); }
				const float cost = __fadd_rn(__fmul_rn(__uint2float_rn(curCount), half_area(cmin, cmax)), bcost[i + 1]);
				if (cost < minCost[axis]) { minBin[axis] = i + 1; minCost[axis] = cost; }
			}
		}
		if (pass == 0) {
			minAxis = 0u;
			if (minCost[1] < minCost[0]) minAxis = 1u;
			if (minCost[2] < minCost[minAxis]) minAxis = 2u;
		}
	}
	const float nmin[3] = { rec.b[0], rec.b[2], rec.b[4] }, nmax[3] = { rec.b[1], rec.b[3], rec.b[5] };
	const float leafCost = __fmul_rn(half_area(nmin, nmax), __fsub_rn(__uint2float_rn(primCount), 1.5f));
	if (minCost[minAxis] > leafCost) {
		if (primCount <= BVB_MAX_LEAF) { rec.left = -1; return; }
		unsigned accum = 0u, best = primCount;                        /* approximate median split (bvh.c:196-205) */
		for (unsigned i = 0; i < BVB_BINS - 1; ++i) {
			accum += B.cnt[minAxis][i];
			const int dlt = (int)primCount / 2 - (int)accum;
			const unsigned approx = (unsigned)(dlt < 0 ? -dlt : dlt);
			if (approx < best) { best = approx; minBin[minAxis] = i + 1; }
		}
	}
	const unsigned split = minBin[minAxis];
	for (int k = 0; k < 3; ++k) { bestLb[2 * k] = BVB_FLT_MAX; bestLb[2 * k + 1] = -BVB_FLT_MAX; bestRb[2 * k] = BVB_FLT_MAX; bestRb[2 * k + 1] = -BVB_FLT_MAX; }
	for (unsigned i = 0; i < split; ++i)
		for (int k = 0; k < 3; ++k) { bestLb[2 * k] = MACRO_MIN(bestLb[2 * k], bmin[i][k]); bestLb[2 * k + 1] = MACRO_MAX(bestLb[2 * k + 1], bmax[i][k]); }
	for (unsigned i = split; i < BVB_BINS; ++i)
		for (int k = 0; k < 3; ++k) { bestRb[2 * k] = MACRO_MIN(bestRb[2 * k], bmin[i][k]); bestRb[2 * k + 1] = MACRO_MAX(bestRb[2 * k + 1], bmax[i][k]); }
	rec.axis = minAxis; rec.split = split;
	for (int k = 0; k < 6; ++k) { rec.lb[k] = bestLb[k]; rec.rb[k] = bestRb[k]; }
	/* rec.left stays -2: the partition decides whether anything goes left (bvh.c:229) */
}

/* ---- partition (bvh.c:97-135) ---------------------------------------------------------------------------------------------------- */
__global__ void k_flag(const float *__restrict__ ctr, const int32_t *__restrict__ prims, const int32_t *__restrict__ owner,
					   const Rec *__restrict__ recs, unsigned long long *flags, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n) return;
	unsigned long long f = 0ull;
	const int32_t r = owner[pos];
	if (r >= 0 && recs[r].left == -2) {
		const Rec &rec = recs[r];
		const int axis = (int)rec.axis;
		f = bin_index(ctr[3 * (size_t)prims[pos] + axis], rec.b[2 * axis], bin_scale(rec.b, axis)) < rec.split ? 1ull : 0ull;
	}
	flags[pos] = f;
}
__global__ void k_mid(Rec *recs, const uint32_t *__restrict__ open_list, uint32_t nopen, const unsigned long long *__restrict__ S1) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nopen) return;
	Rec &rec = recs[open_list[s]];
	if (rec.left != -2) return;
	const unsigned long long before = rec.begin ? S1[rec.begin - 1u] : 0ull;
	rec.mid = rec.begin + (uint32_t)(S1[rec.end - 1u] - before);
	if (rec.mid <= rec.begin) rec.left = -1;                           /* bvh.c:229: nothing went left -> leaf */
}
/* misplaced elements: low half = flag-0 elements left of mid, high half = flag-1 elements right of mid */
__global__ void k_misplaced(const int32_t *__restrict__ owner, const Rec *__restrict__ recs, const unsigned long long *__restrict__ flags,
							unsigned long long *mis, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n) return;
	unsigned long long m = 0ull;
	const int32_t r = owner[pos];
	if (r >= 0 && recs[r].left == -2) {
		const bool fl = flags[pos] != 0ull;
		if (pos < recs[r].mid) { if (!fl) m = 1ull; } else if (fl) m = 1ull << 32;
	}
	mis[pos] = m;
}
__global__ void k_scatter_right(const int32_t *__restrict__ owner, const Rec *__restrict__ recs, const unsigned long long *__restrict__ mis,
								const unsigned long long *__restrict__ S2, uint32_t *tmpR, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n || !(mis[pos] >> 32)) return;
	const Rec &rec = recs[owner[pos]];
	const uint32_t rankR = (uint32_t)(S2[rec.end - 1u] >> 32) - (uint32_t)(S2[pos] >> 32);     /* 0 = the rightmost one */
	tmpR[rec.begin + rankR] = pos;
}
__global__ void k_swap(const int32_t *__restrict__ owner, const Rec *__restrict__ recs, const unsigned long long *__restrict__ mis,
					   const unsigned long long *__restrict__ S2, const uint32_t *__restrict__ tmpR, int32_t *prims, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n || !(mis[pos] & 0xffffffffull)) return;
	const Rec &rec = recs[owner[pos]];
	const uint32_t before = rec.begin ? (uint32_t)(S2[rec.begin - 1u] & 0xffffffffull) : 0u;
	const uint32_t k = (uint32_t)(S2[pos] & 0xffffffffull) - before - 1u;                        /* 0 = the leftmost one */
	const uint32_t partner = tmpR[rec.begin + k];
	const int32_t a = prims[pos], b = prims[partner];
	prims[pos] = b; prims[partner] = a;
}

/* ---- children (bvh.c:221-239) ----------------------------------------------------------------------------------------------------- */
__global__ void k_children(Rec *recs, const uint32_t *__restrict__ open_list, uint32_t nopen, uint32_t *next_list, uint32_t *counters) {
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= nopen) return;
	const uint32_t ri = open_list[s];
	Rec &rec = recs[ri];
	if (rec.left != -2) return;
	const uint32_t c = atomicAdd(&counters[0], 2u);
	rec.left = (int32_t)c;
	for (int side = 0; side < 2; ++side) {
		Rec ch;
		memset(&ch, 0, sizeof ch);
		ch.begin = side ? rec.mid : rec.begin;
		ch.end = side ? rec.end : rec.mid;
		ch.depth = rec.depth + 1u;
		for (int k = 0; k < 6; ++k) ch.b[k] = side ? rec.rb[k] : rec.lb[k];
		const bool open = ch.depth < BVB_MAX_DEPTH && ch.end - ch.begin >= 2u;     /* bvh.c:138-141 */
		ch.left = open ? -2 : -1;
		if (open) { ch.slot = atomicAdd(&counters[2], 1u); next_list[ch.slot] = c + (uint32_t)side; }
		recs[c + side] = ch;
	}
}
__global__ void k_owner(int32_t *owner, const Rec *__restrict__ recs, uint32_t n) {
	const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n) return;
	const int32_t r = owner[pos];
	if (r < 0) return;
	const Rec &rec = recs[r];
	int32_t o = -1;
	if (rec.left >= 0) {
		const int32_t child = rec.left + (pos < rec.mid ? 0 : 1);
		if (recs[child].left == -2) o = child;
	}
	owner[pos] = o;
}
__global__ void k_level_advance(uint32_t *counters) { counters[1] = counters[2]; counters[2] = 0u; }

/* ---- numbering ------------------------------------------------------------------------------------------------------------------------ */
__global__ void k_count_internals(Rec *recs, uint32_t first, uint32_t count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	Rec &r = recs[first + i];
	r.internals = r.left >= 0 ? 1u + recs[r.left].internals + recs[r.left + 1].internals : 0u;
}
__global__ void k_assign(Rec *recs, uint32_t first, uint32_t count) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	const Rec &r = recs[first + i];
	if (r.left < 0) return;
	Rec &l = recs[r.left], &rr = recs[r.left + 1];
	l.rank = r.rank + 1u; rr.rank = r.rank + 1u + l.internals;
	l.idx = 1u + 2u * r.rank; rr.idx = 2u + 2u * r.rank;
}
__global__ void k_emit(const Rec *__restrict__ recs, uint32_t nrec, crs_bvh_node *out) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nrec) return;
	const Rec &r = recs[i];
	crs_bvh_node nd;
	for (int k = 0; k < 6; ++k) nd.bounds[k] = r.b[k];
	if (r.left >= 0) { nd.first_child_or_prim = 1u + 2u * r.rank; nd.prim_count_leaf = 0u; }
	else { nd.first_child_or_prim = r.begin; nd.prim_count_leaf = CRS_BVH_LEAF_BIT | ((r.end - r.begin) & CRS_BVH_COUNT_MASK); }
	out[r.idx] = nd;
}
__global__ void k_iota(int32_t *prims, int32_t *owner, uint32_t n, int32_t owner0) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { prims[i] = (int32_t)i; owner[i] = owner0; }
}

/* ---- inclusive prefix sum of 64-bit words (both halves at once: the halves never carry into each other, counts < 2^32) ------------- */
__global__ void __launch_bounds__(256) k_scan_blocks(const unsigned long long *__restrict__ in, unsigned long long *out, unsigned long long *sums, uint32_t n) {
	__shared__ unsigned long long s_warp[8];
	const uint32_t base = blockIdx.x * BVB_SCAN_BLOCK + threadIdx.x * 4u;
	unsigned long long v[4], acc = 0ull;
	for (int k = 0; k < 4; ++k) { v[k] = base + k < n ? in[base + k] : 0ull; acc += v[k]; v[k] = acc; }
	unsigned long long incl = acc;
	const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
	for (unsigned d = 1; d < 32; d <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += x; }
	if (lane == 31u) s_warp[w] = incl;
	__syncthreads();
	unsigned long long woff = 0ull;
	for (unsigned k = 0; k < w; ++k) woff += s_warp[k];
	const unsigned long long excl = woff + incl - acc;
	for (int k = 0; k < 4; ++k) if (base + k < n) out[base + k] = excl + v[k];
	if (threadIdx.x == 255u && sums) sums[blockIdx.x] = woff + incl;
}
__global__ void k_scan_add(unsigned long long *out, const unsigned long long *__restrict__ sums_incl, uint32_t n) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t blk = i / BVB_SCAN_BLOCK;
	if (i < n && blk > 0u) out[i] += sums_incl[blk - 1u];
}

struct DevBuf {
	void *p = nullptr;
	~DevBuf() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
	template <class T> T *as() { return static_cast<T *>(p); }
};

}  // namespace

#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return crg_fail(e_ == cudaErrorMemoryAllocation ? CRGPU_ERR_NOMEM : CRGPU_ERR_CUDA, "crgpu_bvh_build: %s: %s", #call, cudaGetErrorString(e_)); } while (0)

static int scan64(const unsigned long long *in, unsigned long long *out, unsigned long long *sums, unsigned long long *sums2, uint32_t n, cudaStream_t st) {
	const uint32_t nb = (n + BVB_SCAN_BLOCK - 1u) / BVB_SCAN_BLOCK;
	k_scan_blocks<<<nb, 256, 0, st>>>(in, out, sums, n);
	if (nb > 1u) {
		const uint32_t nb2 = (nb + BVB_SCAN_BLOCK - 1u) / BVB_SCAN_BLOCK;        /* <= 4 for n <= 4M */
		k_scan_blocks<<<nb2, 256, 0, st>>>(sums, sums, sums2, nb);
		if (nb2 > 1u) {
			k_scan_blocks<<<1, 256, 0, st>>>(sums2, sums2, nullptr, nb2);
			k_scan_add<<<(nb + 255u) / 256u, 256, 0, st>>>(sums, sums2, nb);
		}
		k_scan_add<<<(n + 255u) / 256u, 256, 0, st>>>(out, sums, n);
	}
	return 0;
}

extern "C" int crgpu_bvh_build(const float *bboxes, const float *centers, uint32_t n, int device,
								struct crs_bvh_node *nodes_out, uint32_t *node_count_out, int32_t *prims_out) {
	if (!nodes_out || !node_count_out || !prims_out || (n && (!bboxes || !centers))) return crg_fail(CRGPU_ERR_BAD_ARGUMENT, "crgpu_bvh_build: NULL argument");
	*node_count_out = 0u;
	if (n == 0u) return CRGPU_OK;                                           /* bvh.c:251: an empty BVH has no nodes */
	if (n > 4194304u) return crg_fail(CRGPU_ERR_UNSUPPORTED, "crgpu_bvh_build: %u primitives (limit 4194304 per call)", n);
	int ndev = 0;
	int rc = crgpu_device_count(&ndev);
	if (rc) return rc;
	if (device < 0 || device >= ndev) return crg_fail(CRGPU_ERR_BAD_ARGUMENT, "crgpu_bvh_build: device %d out of range (have %d)", device, ndev);
	CUB(cudaSetDevice(device));
	const uint32_t max_open = n / 2u + 1u, max_rec = 2u * n;
	size_t free_b = 0, total_b = 0;
	CUB(cudaMemGetInfo(&free_b, &total_b));
	const size_t need = (size_t)max_open * sizeof(Bins) + (size_t)max_rec * sizeof(Rec) + (size_t)n * 96u;
	if (need > free_b / 2u) return crg_fail(CRGPU_ERR_NOMEM, "crgpu_bvh_build: %zu bytes of scratch for %u primitives, %zu free", need, n, free_b);
	DevBuf d_bb, d_ctr, d_prims, d_owner, d_recs, d_bins, d_open[2], d_flags, d_S1, d_mis, d_S2, d_tmpR, d_sums, d_sums2, d_keys, d_cnt, d_out;
	CUB(d_bb.alloc((size_t)n * 24u)); CUB(d_ctr.alloc((size_t)n * 12u)); CUB(d_prims.alloc((size_t)n * 4u)); CUB(d_owner.alloc((size_t)n * 4u));
	CUB(d_recs.alloc((size_t)max_rec * sizeof(Rec))); CUB(d_bins.alloc((size_t)max_open * sizeof(Bins)));
	CUB(d_open[0].alloc((size_t)max_open * 4u)); CUB(d_open[1].alloc((size_t)max_open * 4u));
	CUB(d_flags.alloc((size_t)n * 8u)); CUB(d_S1.alloc((size_t)n * 8u)); CUB(d_mis.alloc((size_t)n * 8u)); CUB(d_S2.alloc((size_t)n * 8u));
	CUB(d_tmpR.alloc((size_t)n * 4u)); CUB(d_sums.alloc(8u * ((size_t)n / BVB_SCAN_BLOCK + 2u))); CUB(d_sums2.alloc(8u * 64u));
	CUB(d_keys.alloc(6u * 8u)); CUB(d_cnt.alloc(4u * 4u)); CUB(d_out.alloc((size_t)max_rec * sizeof(crs_bvh_node)));
	cudaStream_t st = nullptr;
	CUB(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{ st };
	CUB(cudaMemcpyAsync(d_bb.p, bboxes, (size_t)n * 24u, cudaMemcpyHostToDevice, st));
	CUB(cudaMemcpyAsync(d_ctr.p, centers, (size_t)n * 12u, cudaMemcpyHostToDevice, st));
	const unsigned long long init_keys[6] = { KMIN_EMPTY, KMIN_EMPTY, KMIN_EMPTY, KMAX_EMPTY, KMAX_EMPTY, KMAX_EMPTY };
	CUB(cudaMemcpyAsync(d_keys.p, init_keys, sizeof init_keys, cudaMemcpyHostToDevice, st));
	const uint32_t gn = (n + 255u) / 256u;
	Rec *recs = d_recs.as<Rec>();
	uint32_t *counters = d_cnt.as<uint32_t>();
	k_iota<<<gn, 256, 0, st>>>(d_prims.as<int32_t>(), d_owner.as<int32_t>(), n, n >= 2u ? 0 : -1);
	k_root_keys<<<gn, 256, 0, st>>>(d_bb.as<float>(), n, d_keys.as<unsigned long long>());
	k_root_init<<<1, 1, 0, st>>>(d_bb.as<float>(), n, d_keys.as<unsigned long long>(), recs, counters);
	const uint32_t zero = 0u;
	CUB(cudaMemcpyAsync(d_open[0].p, &zero, 4u, cudaMemcpyHostToDevice, st));          /* the root is open node 0 of level 0 */
	std::vector<uint32_t> level_first{ 0u }, level_count{ 1u };
	uint32_t h[4] = { 1u, n >= 2u ? 1u : 0u, 0u, 0u };
	int cur = 0;
	while (h[1] > 0u) {
		const uint32_t nopen = h[1], go = (nopen + 127u) / 128u;
		if (nopen > max_open) return crg_fail(CRGPU_ERR_CUDA, "crgpu_bvh_build: %u open nodes for %u primitives", nopen, n);
		uint32_t *open_list = d_open[cur].as<uint32_t>(), *next_list = d_open[cur ^ 1].as<uint32_t>();
		{
			const size_t blocks = ((size_t)nopen * (sizeof(Bins) / 8u) + 255u) / 256u;          /* grid-stride loop inside */
			k_bins_clear<<<(unsigned)(blocks > 65535u ? 65535u : blocks), 256, 0, st>>>(d_bins.as<Bins>(), nopen);
		}
		k_bin<<<gn, 256, 0, st>>>(d_bb.as<float>(), d_ctr.as<float>(), d_prims.as<int32_t>(), d_owner.as<int32_t>(), recs, d_bins.as<Bins>(), n);
		k_split<<<go, 128, 0, st>>>(d_bb.as<float>(), d_prims.as<int32_t>(), recs, open_list, nopen, d_bins.as<Bins>());
		k_flag<<<gn, 256, 0, st>>>(d_ctr.as<float>(), d_prims.as<int32_t>(), d_owner.as<int32_t>(), recs, d_flags.as<unsigned long long>(), n);
		scan64(d_flags.as<unsigned long long>(), d_S1.as<unsigned long long>(), d_sums.as<unsigned long long>(), d_sums2.as<unsigned long long>(), n, st);
		k_mid<<<go, 128, 0, st>>>(recs, open_list, nopen, d_S1.as<unsigned long long>());
		k_misplaced<<<gn, 256, 0, st>>>(d_owner.as<int32_t>(), recs, d_flags.as<unsigned long long>(), d_mis.as<unsigned long long>(), n);
		scan64(d_mis.as<unsigned long long>(), d_S2.as<unsigned long long>(), d_sums.as<unsigned long long>(), d_sums2.as<unsigned long long>(), n, st);
		k_scatter_right<<<gn, 256, 0, st>>>(d_owner.as<int32_t>(), recs, d_mis.as<unsigned long long>(), d_S2.as<unsigned long long>(), d_tmpR.as<uint32_t>(), n);
		k_swap<<<gn, 256, 0, st>>>(d_owner.as<int32_t>(), recs, d_mis.as<unsigned long long>(), d_S2.as<unsigned long long>(), d_tmpR.as<uint32_t>(), d_prims.as<int32_t>(), n);
		k_children<<<go, 128, 0, st>>>(recs, open_list, nopen, next_list, counters);
		k_owner<<<gn, 256, 0, st>>>(d_owner.as<int32_t>(), recs, n);
		CUB(cudaMemcpyAsync(h, counters, sizeof h, cudaMemcpyDeviceToHost, st));
		CUB(cudaStreamSynchronize(st));
		const uint32_t prev_total = level_first.back() + level_count.back();
		if (h[0] > prev_total) { level_first.push_back(prev_total); level_count.push_back(h[0] - prev_total); }
		k_level_advance<<<1, 1, 0, st>>>(counters);
		h[1] = h[2];
		cur ^= 1;
		if (level_first.size() > BVB_MAX_DEPTH + 2u) return crg_fail(CRGPU_ERR_CUDA, "crgpu_bvh_build: deeper than %u levels", BVB_MAX_DEPTH);
	}
	const uint32_t nrec = h[0];
	for (size_t l = level_first.size(); l-- > 0;)
		k_count_internals<<<(level_count[l] + 255u) / 256u, 256, 0, st>>>(recs, level_first[l], level_count[l]);
	for (size_t l = 0; l < level_first.size(); ++l)
		k_assign<<<(level_count[l] + 255u) / 256u, 256, 0, st>>>(recs, level_first[l], level_count[l]);
	k_emit<<<(nrec + 255u) / 256u, 256, 0, st>>>(recs, nrec, d_out.as<crs_bvh_node>());
	CUB(cudaMemcpyAsync(nodes_out, d_out.p, (size_t)nrec * sizeof(crs_bvh_node), cudaMemcpyDeviceToHost, st));
	CUB(cudaMemcpyAsync(prims_out, d_prims.p, (size_t)n * 4u, cudaMemcpyDeviceToHost, st));
	CUB(cudaStreamSynchronize(st));
	CUB(cudaGetLastError());
	*node_count_out = nrec;
	return CRGPU_OK;
}
