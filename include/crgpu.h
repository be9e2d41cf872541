/*
 * crgpu.h — C ABI of the B200 path-trace hot loop (libcrgpu.so).
 *
 * This is the drop-in boundary.  c-ray's tile dispatcher (`renderFrame`, reference
 * src/renderer/renderer.c:40-180) starts worker threads through a `struct crThread` slot whose
 * `threadFunc` is today `renderThread` (renderer.c:258-327) or `networkRenderThread`
 * (src/utils/protocol/server.c:215-263).  A GPU worker thread occupies the same slot
 * (c-ray_b200/host/cr_renderer.c `gpuRenderThread`; the reference-side binding is c-ray_b200/integration/gpu_thread.c) and calls ONLY the functions
 * below — plain pointers and sizes, no CUDA or torch types.
 *
 *   reference interface replaced                                  entry point here
 *   -----------------------------------------------------------   ---------------------------------
 *   (none; scene is read through pointers, scene.h:14-39)          crgpu_scene_create / _destroy
 *   renderThread inner loops, renderer.c:271-320                   crgpu_render_tile
 *     initSampler  sampler.c:41-44   getCameraRay camera.c:58-87
 *     pathTrace    pathtrace.c:32-60 traverseTopLevelBvh bvh.c:488
 *     bsdf->sample nodes/shaders     running average renderer.c:288-294
 *   colorToSRGB + setPixel(image), renderer.c:297-300              crgpu_framebuffer_to_srgb8
 *   textureGetPixel(renderBuffer), renderer.c:283                  crgpu_framebuffer_read / _write / _clear
 *   tile gather of the cluster mode, server.c:159-174              crgpu_framebuffer_device_ptr (+ NCCL in the host)
 *   buildBvhGeneric, src/accelerators/bvh.c:245-296                crgpu_bvh_build (optional: the loaders build on the host)
 *
 * All functions return 0 on success or a negative CRGPU_ERR_* code; crgpu_last_error() gives text.
 * The host turns errors into logr(warning|error, ...) like the rest of c-ray (logging.c:50-74).
 * There is NO CPU fallback: without a CUDA device every call fails with CRGPU_ERR_NO_DEVICE.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "crscene.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CRGPU_OK                 0
#define CRGPU_ERR_NO_DEVICE     -1
#define CRGPU_ERR_CUDA          -2
#define CRGPU_ERR_BAD_ARGUMENT  -3
#define CRGPU_ERR_UNSUPPORTED   -4   /* node graph deeper than the device interpreter supports, etc. */
#define CRGPU_ERR_NOMEM         -5

typedef struct crgpu_scene crgpu_scene;

/* Per-call statistics (all counted on the device). */
struct crgpu_stats {
	uint64_t paths;          /* pixel-passes started (one pathTrace call each) */
	uint64_t rays;           /* closest-hit queries, one per getClosestIsect (pathtrace.c:38) */
	uint64_t node_pairs;     /* P: BVH child-pair steps, both levels   (only with CRGPU_FLAG_COUNT) */
	uint64_t tri_tests;      /* T: triangle tests                      (only with CRGPU_FLAG_COUNT) */
	uint64_t sphere_tests;   /* S: sphere tests                        (only with CRGPU_FLAG_COUNT) */
	uint64_t inst_visits;    /* I: mesh-instance visits                (only with CRGPU_FLAG_COUNT) */
	uint64_t kernel_launches;/* CUDA kernels launched by this call */
	float    trace_ms;       /* device time inside the BVH-traversal kernel (CUDA events), if CRGPU_FLAG_TIME_KERNELS */
	float    shade_ms;       /* device time inside the shade kernel, same condition */
	float    total_ms;       /* device time of the whole call on its stream, same condition */
	float    pad;
};

#define CRGPU_FLAG_COUNT         0x1u  /* run the instrumented traversal kernel (slower; fills P/T/S/I) */
#define CRGPU_FLAG_TIME_KERNELS  0x2u  /* bracket every kernel with CUDA events (adds sync at the end only) */
#define CRGPU_FLAG_ASYNC         0x4u  /* only enqueue on the scene's stream; fetch the numbers later with crgpu_get_stats */

int crgpu_device_count(int *n);
const char *crgpu_last_error(void);

/* Upload a flattened scene (include/crscene.h) to `device`.  The device keeps its own replicas; the
 * flat arrays may be freed after the call.  Also allocates the fp32 framebuffer (W*H*3, row H-1-y:
 * the layout of state.renderBuffer, scene.c:200 + texture.c:24-28), zero-initialised. */
int crgpu_scene_create(const struct crs_scene *flat, int device, crgpu_scene **out);
int crgpu_scene_destroy(crgpu_scene *s);

/* The two halves of crgpu_scene_create, for hosts that render many frames of one scene or one frame on several GPUs (the
 * reference rebuilds nothing between frames either: loadScene builds the BVHs once, scene.c:111-213).  crgpu_prepare
 * validates the flat scene and re-lays it out for the kernels into ONE pinned host slab (device-independent, host threads);
 * crgpu_scene_create_prepared is then a single host->device copy.  Image size / sample count / bounces / camera are taken
 * from the flat scene at prepare time; crgpu_prepared_update_config re-reads them (after crscene_set_config) without
 * repacking.  crgpu_prepared_slab exposes the slab bytes (what crosses PCIe per frame; tests pin their checksum). */
typedef struct crgpu_prepared crgpu_prepared;
int  crgpu_prepare(const struct crs_scene *flat, crgpu_prepared **out);
int  crgpu_prepared_update_config(crgpu_prepared *p, const struct crs_scene *flat);
int  crgpu_prepared_slab(const crgpu_prepared *p, const void **slab, size_t *bytes);
void crgpu_prepared_free(crgpu_prepared *p);
int  crgpu_scene_create_prepared(const crgpu_prepared *p, int device, crgpu_scene **out);

/* Device memory released by crgpu_scene_destroy is kept in a per-device cache for the next scene (a frame's 35 GB of
 * wavefront state costs 10-55 ms to cudaMalloc and 0.3-0.5 s to cudaFree); crgpu_device_trim returns it to the driver. */
int crgpu_device_trim(int device);
/* Page-locked host memory for buffers that cross PCIe every frame (renderBuffer, scene.c:200); NULL on failure. */
void *crgpu_host_alloc(size_t bytes);
void  crgpu_host_free(void *p);

/* Limit on paths in flight per wavefront batch (default: what fits in 40% of the free device memory at
 * 137 B per path, at most 256M); a tile's passes are processed in batches of floor(max_paths / tile_pixels)
 * passes.  Results do not depend on it; throughput does (every batch pays a fixed latency chain). */
int crgpu_set_max_paths_in_flight(crgpu_scene *s, uint64_t max_paths);

/* Render passes [pass_begin, pass_begin+pass_count) of the tile [x0,x1) x [y0,y1) (y up, end
 * exclusive: `struct renderTile`, tile.h:28-37) into the device framebuffer, continuing the running
 * average stored there.  maxPasses and bounces come from the scene prefs (they seed the sampler,
 * sampler.c:42).  Synchronous (returns when the device is done, `stats` = this call) unless
 * CRGPU_FLAG_ASYNC is given.  `stats` may be NULL.  A tile with more pixels than the paths-in-flight budget
 * is walked in pieces internally. */
int crgpu_render_tile(crgpu_scene *s, int x0, int y0, int x1, int y1,
					  int pass_begin, int pass_count, unsigned flags, struct crgpu_stats *stats);

/* Same, over a UNION of tiles in one wavefront (rects = 4 ints per tile: x0, y0, x1, y1).  This is how a
 * rank of a multi-GPU job renders its share of the tile grid without paying the per-tile launch tails. */
int crgpu_render_tiles(crgpu_scene *s, const int *rects, int nrects,
					   int pass_begin, int pass_count, unsigned flags, struct crgpu_stats *stats);

/* Run on a caller-owned CUDA stream (a `cudaStream_t` passed as void*; NULL means the legacy default
 * stream, as everywhere in CUDA).  Lets a host that already has a stream (e.g. the one its NCCL collectives
 * use) order the kernels with its own work and time them with its own events.  crgpu_use_own_stream goes
 * back to the scene's private non-blocking stream. */
int crgpu_set_stream(crgpu_scene *s, void *cuda_stream);
int crgpu_use_own_stream(crgpu_scene *s);
/* Synchronise the stream and return the statistics accumulated since the previous fetch (by
 * crgpu_get_stats or by a synchronous crgpu_render_tile). */
int crgpu_get_stats(crgpu_scene *s, struct crgpu_stats *stats);

/* Framebuffer access.  Host variants copy rows of the tile rectangle (or the whole frame when
 * x1<=x0) between the device framebuffer and a host buffer with the SAME W*H*3 layout. */
int crgpu_framebuffer_clear(crgpu_scene *s);
int crgpu_framebuffer_read(crgpu_scene *s, float *host_rgb, int x0, int y0, int x1, int y1);
int crgpu_framebuffer_write(crgpu_scene *s, const float *host_rgb, int x0, int y0, int x1, int y1);
/* 8-bit sRGB image of the current framebuffer (renderer.c:297-300), W*H*3 bytes, same row order. */
int crgpu_framebuffer_to_srgb8(crgpu_scene *s, uint8_t *host_rgb8);
/* Raw device pointer of the fp32 framebuffer (for the NCCL gather done by the multi-GPU host). */
int crgpu_framebuffer_device_ptr(crgpu_scene *s, void **dev_ptr, size_t *bytes);
/* Device ordinal and image size of a scene (any pointer may be NULL). */
int crgpu_scene_info(crgpu_scene *s, int *device, int *width, int *height);

/* Known-answer hook used by the parity tests: camera ray → closest hit → bsdf sample for `count`
 * (x, y, pass) triples, in the 160-byte record layout of oracle/ref_harness.c `struct hit_kat`. */
int crgpu_trace_kat(crgpu_scene *s, const int32_t *xyp, int count, void *records_out);

/* the reference's binned-SAH BVH build (reference src/accelerators/bvh.c:96-296: partitionPrimitives,
 * buildBvhRecursive, buildBvhGeneric) on device `device`, level-synchronous (c-ray_b200/csrc/crgpu_bvh_build.cu).  Produces the
 * SAME tree as the reference and as the host builder (crloader_build_bvh): node order, leaf ranges, primitive order, bounds.
 *   bboxes   n x 6 floats: min.x min.y min.z max.x max.y max.z of every primitive (what getBBoxAndCenter returns, bvh.c:283-291)
 *   centers  n x 3 floats
 *   nodes_out  room for 2n-1 nodes;  prims_out  n indices (primIndices);  *node_count_out  nodes written
 * n == 0 writes nothing (an empty BVH, bvh.c:251).  At most 4,194,304 primitives per call. */
int crgpu_bvh_build(const float *bboxes, const float *centers, uint32_t n, int device,
                    struct crs_bvh_node *nodes_out, uint32_t *node_count_out, int32_t *prims_out);

#ifdef __cplusplus
}
#endif
