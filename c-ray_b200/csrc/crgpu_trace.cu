/*
 * crgpu_trace.cu — K1 (primary-ray generation) and K2 (two-level BVH traversal): see crgpu_wave.cuh.
 * Everything here is force-inlined; the scene descriptor travels as a by-value kernel parameter so
 * its pointers sit in the constant bank.
 */
#include "crgpu_wave.cuh"
#include "crgpu_trace.cuh"

/* ---- K1 ------------------------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) k_generate(DevScene sc, WaveBuffers wb, TileDesc td) {
	const unsigned tile_pixels = (unsigned)(td.tw * td.th);
	const unsigned n = tile_pixels * (unsigned)td.pass_count;
	for (unsigned id = blockIdx.x * blockDim.x + threadIdx.x; id < n; id += gridDim.x * blockDim.x) {
		const unsigned pl = id / tile_pixels, px = id - pl * tile_pixels;
		const int x = td.x0 + (int)(px % (unsigned)td.tw);
		const int y = td.y0 + (int)(px / (unsigned)td.tw);
		const uint32_t pixIdx = (uint32_t)(y * (int)sc.image_width + x);                        /* renderer.c:280 */
		uint64_t rng = cr_rng_init(pixIdx, (uint32_t)(td.pass_begin + (int)pl), sc.sample_count);
		v3 o, d;
		cr_camera_ray(sc.cam, x, y, rng, o, d);
		wb.stA[0][id] = make_float4(o.x, o.y, o.z, d.x);
		wb.stB[0][id] = make_float4(d.y, d.z, 1.0f, 1.0f);
		wb.stC[0][id] = make_uint4(__float_as_uint(1.0f), id, (unsigned)(rng & 0xffffffffull), (unsigned)(rng >> 32));
		wb.L[id] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) { wb.counts[0] = n; wb.counts[1] = 0u; }
}

/* ---- K2 ------------------------------------------------------------------------------------------------------------ */
template <bool COUNT>
__global__ void __launch_bounds__(256) k_trace(DevScene sc, WaveBuffers wb, int cur) {
	const unsigned n = wb.counts[cur];
	TraceCounters tc = { 0u, 0u, 0u, 0u };
	for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float4 a = wb.stA[cur][i];
		const float4 b = wb.stB[cur][i];
		const Hit h = cr_closest_hit<COUNT>(sc, v3make(a.x, a.y, a.z), v3make(a.w, b.x, b.y), &tc);
		wb.hit[i] = make_float4(h.t, h.u, h.v, __uint_as_float(h.prim));
		wb.hitInst[i] = h.inst;
	}
	if (COUNT) {
		atomicAdd(&wb.stats[1], (unsigned long long)tc.pairs);
		atomicAdd(&wb.stats[2], (unsigned long long)tc.tris);
		atomicAdd(&wb.stats[3], (unsigned long long)tc.spheres);
		atomicAdd(&wb.stats[4], (unsigned long long)tc.insts);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		wb.stats[0] += n;          /* one ray per getClosestIsect */
		wb.counts[cur ^ 1] = 0u;   /* K3 of this bounce appends survivors there */
	}
}


void crg_launch_generate(const DevScene &sc, const WaveBuffers &wb, const TileDesc &td, int grid, cudaStream_t st) {
	k_generate<<<grid, 256, 0, st>>>(sc, wb, td);
}
void crg_launch_trace(const DevScene &sc, const WaveBuffers &wb, int cur, bool count, int grid, cudaStream_t st) {
	if (count) k_trace<true><<<grid, 256, 0, st>>>(sc, wb, cur);
	else k_trace<false><<<grid, 256, 0, st>>>(sc, wb, cur);
}
