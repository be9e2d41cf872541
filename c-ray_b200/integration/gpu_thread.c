/*
 * gpu_thread.c — REFERENCE-SIDE integration code: the worker a c-ray maintainer adds as
 * src/renderer/gpu_thread.c.  It occupies the thread-function slot of `struct crThread` that renderFrame fills with
 * renderThread (reference src/renderer/renderer.c:92-105, src/utils/platform/thread.h:22-31) and calls ONLY the C ABI of
 * include/crgpu.h.  Compiles against the reference's headers; built by oracle/Makefile (`cray_ref_gpu`) together with the
 * unmodified reference sources, never into libcrgpu.so / libcrhost.so.
 *
 * Differences from renderThread (renderer.c:258-327), all inside this worker:
 *   - the scene is flattened once per process (flatten_world) and uploaded to the worker's device (thread_num = CUDA device);
 *   - a trip to the tile queue takes several tiles (a GPU wavefront wants ~64M paths in flight; one 64x64 tile is 4096 pixels)
 *     and renders them as ONE wavefront with crgpu_render_tiles;
 *   - pixels land in the device framebuffer and are copied into state.renderBuffer / the 8-bit output per tile afterwards, with
 *     the layout of setPixel (texture.c:19-28), so the preview window, the stats loop and the encoders see what they expect.
 */
#include "includes.h"
#include "renderer/renderer.h"
#include "datatypes/tile.h"
#include "datatypes/scene.h"
#include "datatypes/image/texture.h"
#include "utils/platform/thread.h"
#include "utils/platform/mutex.h"
#include "utils/logging.h"
#include "utils/timer.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "crgpu.h"

int flatten_world(const struct renderer *r, struct crs_scene *out);

static pthread_mutex_t g_flat_lock = PTHREAD_MUTEX_INITIALIZER;
static struct crs_scene g_flat;
static crgpu_prepared *g_prepared;
static const struct world *g_flat_of;

/* flatten + re-layout once per loaded scene; every GPU worker then uploads the same prepared scene */
static crgpu_prepared *prepared_scene(const struct renderer *r) {
	pthread_mutex_lock(&g_flat_lock);
	if (g_flat_of != r->scene || !g_prepared) {
		if (g_prepared) { crgpu_prepared_free(g_prepared); g_prepared = NULL; }
		if (flatten_world(r, &g_flat) == 0 && crgpu_prepare(&g_flat, &g_prepared) == 0) g_flat_of = r->scene;
		else g_prepared = NULL;
	} else {
		/* same world, possibly other prefs (-d / -s between frames): header only */
		g_flat.prefs.image_width = r->prefs.imageWidth; g_flat.prefs.image_height = r->prefs.imageHeight;
		g_flat.prefs.sample_count = (uint32_t)r->prefs.sampleCount; g_flat.prefs.bounces = (uint32_t)r->prefs.bounces;
		crgpu_prepared_update_config(g_prepared, &g_flat);
	}
	crgpu_prepared *p = g_prepared;
	pthread_mutex_unlock(&g_flat_lock);
	return p;
}

#define GPU_TRIP_MAX 1024

void *gpuRenderThread(void *arg) {
	struct renderThreadState *ts = (struct renderThreadState *)threadUserData(arg);      /* thread.c:15-18 */
	struct renderer *r = ts->renderer;
	struct texture *image = ts->output;
	crgpu_scene *gpu = NULL;
	crgpu_prepared *prepared = prepared_scene(r);
	if (!prepared || crgpu_scene_create_prepared(prepared, ts->thread_num /* = CUDA device */, &gpu) != CRGPU_OK) {
		logr(warning, "GPU worker %i: %s\n", ts->thread_num, crgpu_last_error());
		ts->threadComplete = true;
		ts->currentTileNum = -1;
		return 0;
	}
	const unsigned W = r->prefs.imageWidth, H = r->prefs.imageHeight;
	const double tilePaths = (double)r->prefs.tileWidth * r->prefs.tileHeight * (double)r->prefs.sampleCount;
	int *rects = malloc(sizeof(int) * 4 * GPU_TRIP_MAX);
	int *nums = malloc(sizeof(int) * GPU_TRIP_MAX);
	float *fb = malloc(sizeof(float) * 3 * (size_t)W * H);          /* same W*H*3 layout as renderBuffer; only tile rectangles are touched */
	unsigned char *fb8 = malloc((size_t)3 * W * H);
	struct timeval timer = {0};
	ts->completedSamples = 1;
	while (r->state.isRendering) {
		int want = (int)(64.0 * 1048576.0 / (tilePaths > 1.0 ? tilePaths : 1.0)) + 1;
		const int left = r->state.tileCount - r->state.finishedTileCount;
		const int share = (left + r->prefs.threadCount - 1) / (r->prefs.threadCount > 0 ? r->prefs.threadCount : 1);
		if (want > share) want = share;
		if (want > GPU_TRIP_MAX) want = GPU_TRIP_MAX;
		if (want < 1) want = 1;
		int got = 0;
		while (got < want) {
			struct renderTile tile = nextTile(r);                                        /* tile.c:22-45 */
			if (tile.tileNum == -1) break;
			rects[4 * got] = tile.begin.x; rects[4 * got + 1] = tile.begin.y; rects[4 * got + 2] = tile.end.x; rects[4 * got + 3] = tile.end.y;
			nums[got++] = tile.tileNum;
		}
		if (got == 0) break;
		ts->currentTileNum = nums[0];
		startTimer(&timer);
		struct crgpu_stats st;
		if (crgpu_render_tiles(gpu, rects, got, 0, r->prefs.sampleCount, 0u, &st) != CRGPU_OK) {
			logr(warning, "GPU worker %i: %s\n", ts->thread_num, crgpu_last_error());
			break;
		}
		/* what the stats loop of renderFrame reads (renderer.c:126-158) */
		ts->totalSamples += (uint64_t)r->prefs.sampleCount * (uint64_t)got;
		ts->completedSamples = r->prefs.sampleCount;
		ts->avgSampleTime = getUs(timer) / ((long)r->prefs.sampleCount * got);
		/* renderer.c:294-300: the float tile into state.renderBuffer, the sRGB tile into the 8-bit output */
		crgpu_framebuffer_to_srgb8(gpu, fb8);
		for (int i = 0; i < got; ++i) {
			const int *q = rects + 4 * i;
			crgpu_framebuffer_read(gpu, fb, q[0], q[1], q[2], q[3]);
			for (int y = q[1]; y < q[3]; ++y) {
				const size_t off = ((size_t)q[0] + (size_t)(H - (unsigned)(y + 1)) * W) * 3u;       /* texture.c:24-28 */
				const size_t n = (size_t)(q[2] - q[0]) * 3u;
				memcpy(r->state.renderBuffer->data.float_p + off, fb + off, n * sizeof(float));
				memcpy(image->data.byte_p + off, fb8 + off, n);
			}
			r->state.renderTiles[nums[i]].isRendering = false;                           /* renderer.c:315-316 */
			r->state.renderTiles[nums[i]].renderComplete = true;
		}
		while (ts->paused && !r->state.renderAborted) sleepMSec(100);                    /* renderer.c:310-312 */
		ts->currentTileNum = -1;
		ts->completedSamples = 1;
	}
	free(rects); free(nums); free(fb); free(fb8);
	crgpu_scene_destroy(gpu);
	ts->threadComplete = true;                                                            /* renderer.c:323 */
	ts->currentTileNum = -1;
	return 0;
}
