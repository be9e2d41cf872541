/*
 * cr_renderer.c — renderFrame and the GPU worker thread.
 *
 * renderFrame restates reference src/renderer/renderer.c:40-180: allocate the 8-bit output, start
 * prefs.threadCount workers through the thread-function slot, poll their renderThreadState every 16 ms
 * for the progress line, join, return the image.  The workers are gpuRenderThread (one per CUDA device)
 * instead of renderThread (renderer.c:258-327): same tile queue (nextTile), same published fields, but a
 * tile's passes are handed to crgpu_render_tile in one call and pixels never touch the CPU.
 * With more than one GPU the tiles are gathered on device 0 by one NCCL exchange (include/crgpu_nccl.h).
 */
#include "cr_host.h"
#include "../../include/crgpu_nccl.h"
#include <dlfcn.h>
#include "../../include/crloader.h"
#include <strings.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* The NCCL tile gather lives in its own library (libcrgpu_nccl.so) and is loaded on first use: a single-GPU render never
 * touches NCCL, and a host process that already carries another NCCL build (e.g. a Python host with torch) does not get a
 * second one mapped just by loading libcrhost.so.  Looked up next to this library, then on the default search path. */
static struct {
	void *lib;
	int (*create)(crgpu_scene **, int, crgpu_comm **);
	int (*gather)(crgpu_comm *, const int *, const int *, int, int);
	int (*destroy)(crgpu_comm *);
} g_nccl;

int crhost_load_nccl(void) {
	if (g_nccl.gather) return 0;
	char path[4096] = "";
	Dl_info info;
	if (dladdr((void *)crhost_load_nccl, &info) && info.dli_fname) {
		const char *slash = strrchr(info.dli_fname, '/');
		if (slash) snprintf(path, sizeof path, "%.*s/libcrgpu_nccl.so", (int)(slash - info.dli_fname), info.dli_fname);
	}
	void *lib = path[0] ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : NULL;
	if (!lib) lib = dlopen("libcrgpu_nccl.so", RTLD_NOW | RTLD_LOCAL);
	if (!lib) { fprintf(stderr, "renderFrame: cannot load libcrgpu_nccl.so: %s\n", dlerror()); return -1; }
	g_nccl.create = (int (*)(crgpu_scene **, int, crgpu_comm **))dlsym(lib, "crgpu_comm_create");
	g_nccl.destroy = (int (*)(crgpu_comm *))dlsym(lib, "crgpu_comm_destroy");
	void *gather = dlsym(lib, "crgpu_comm_gather_tiles");
	if (!g_nccl.create || !g_nccl.destroy || !gather) { fprintf(stderr, "renderFrame: libcrgpu_nccl.so lacks the crgpu_comm_* entry points\n"); dlclose(lib); return -1; }
	g_nccl.lib = lib;
	g_nccl.gather = (int (*)(crgpu_comm *, const int *, const int *, int, int))gather;
	return 0;
}

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

struct renderer *newRenderer(void) {                              /* renderer.c:329-343 */
	struct renderer *r = calloc(1, sizeof *r);
	if (!r) return NULL;
	pthread_mutex_init(&r->state.tileMutex, NULL);
	r->prefs.threadCount = 1;
	r->prefs.imgType = png;
	r->prefs.imgFilePath = "./";
	r->prefs.imgFileName = "rendered";
	return r;
}


int loadSceneFile(struct renderer *r, const char *path, int width, int height, int samples, int bounces) {
	size_t n = strlen(path);
	if (n > 5 && !strcasecmp(path + n - 5, ".json")) {
		/* a c-ray JSON scene: parse + build both BVH levels on the host (libcrloader.so, include/crloader.h) */
		if (crloader_load_json(&r->scene, path) != 0) {
			fprintf(stderr, "cray_b200: %s\n", crloader_last_error());
			return -1;
		}
	} else if (crscene_load(&r->scene, path) != 0) return -1;
	return applySceneConfig(r, width, height, samples, bounces);
}

int loadSceneBuf(struct renderer *r, const char *json, const char *assetPath, struct crloader_output *output,
				 int width, int height, int samples, int bounces) {
	if (crloader_load_json_buf(&r->scene, json, assetPath, output) != 0) {
		fprintf(stderr, "cray_b200: %s\n", crloader_last_error());
		return -1;
	}
	return applySceneConfig(r, width, height, samples, bounces);
}

int applySceneConfig(struct renderer *r, int width, int height, int samples, int bounces) {
	/* the part of loadScene (src/datatypes/scene.c:111-213) that follows parsing + BVH build */
	crscene_set_config(&r->scene, width, height, samples, bounces);
	r->prefs.imageWidth = r->scene.prefs.image_width;
	r->prefs.imageHeight = r->scene.prefs.image_height;
	r->prefs.sampleCount = (int)r->scene.prefs.sample_count;
	r->prefs.bounces = (int)r->scene.prefs.bounces;
	if (!r->prefs.tileWidth) r->prefs.tileWidth = r->scene.prefs.tile_width ? r->scene.prefs.tile_width : 64;
	if (!r->prefs.tileHeight) r->prefs.tileHeight = r->scene.prefs.tile_height ? r->scene.prefs.tile_height : 64;
	r->prefs.tileOrder = (enum renderOrder)r->scene.prefs.tile_order;
	free(r->state.renderTiles);
	r->state.tileCount = (int)quantizeImage(&r->state.renderTiles, r->prefs.imageWidth, r->prefs.imageHeight,
											r->prefs.tileWidth, r->prefs.tileHeight, r->prefs.tileOrder);   /* scene.c:187-192 */
	free(r->state.tileOwner);
	r->state.tileOwner = calloc((size_t)r->state.tileCount + 1, sizeof(int));
	free(r->state.renderBuffer);
	r->state.renderBuffer = calloc((size_t)r->prefs.imageWidth * r->prefs.imageHeight * 3, sizeof(float));   /* scene.c:200 */
	return r->state.renderBuffer && r->state.renderTiles ? 0 : -1;
}

/* How many tiles a GPU worker takes from the queue per trip.  A CPU thread takes one (renderer.c:265,318); a GPU
 * wavefront pays a fixed ~7 ms latency chain per batch, so a worker takes enough tiles to put ~64M paths in flight,
 * but never more than its fair share of what is left (keeps several GPUs balanced). */
static int tiles_per_trip(const struct renderer *r) {
	const double tilePaths = (double)r->prefs.tileWidth * r->prefs.tileHeight * (double)r->prefs.sampleCount;
	int want = (int)(64.0 * 1048576.0 / (tilePaths > 1.0 ? tilePaths : 1.0)) + 1;
	const int left = r->state.tileCount - r->state.finishedTileCount;
	const int share = (left + r->prefs.threadCount - 1) / (r->prefs.threadCount > 0 ? r->prefs.threadCount : 1);
	if (want > share) want = share;
	if (want > 1024) want = 1024;
	return want < 1 ? 1 : want;
}

void *gpuRenderThread(void *arg) {
	struct renderThreadState *ts = arg;
	struct renderer *r = ts->renderer;
	int *rects = malloc(sizeof(int) * 4 * 1024);
	int *nums = malloc(sizeof(int) * 1024);
	while (r->state.isRendering && !r->state.renderAborted) {
		/* take a handful of tiles from the shared queue (nextTile, tile.c:22-45) */
		const int want = tiles_per_trip(r);
		int got = 0;
		while (got < want) {
			struct renderTile tile = nextTile(r);
			if (tile.tileNum == -1) break;
			rects[4 * got] = tile.begin.x; rects[4 * got + 1] = tile.begin.y; rects[4 * got + 2] = tile.end.x; rects[4 * got + 3] = tile.end.y;
			nums[got++] = tile.tileNum;
		}
		if (got == 0) break;
		ts->currentTileNum = nums[0];
		const double t0 = now_s();
		struct crgpu_stats st;
		const int rc = crgpu_render_tiles(ts->gpu, rects, got, 0, r->prefs.sampleCount, 0u, &st);
		if (rc != CRGPU_OK) {
			fprintf(stderr, "gpuRenderThread[%d]: %s\n", ts->thread_num, crgpu_last_error());
			ts->error = rc;
			r->state.renderAborted = true;
			break;
		}
		ts->rays += st.rays;
		ts->totalSamples += (uint64_t)r->prefs.sampleCount * (uint64_t)got;      /* tile-passes done, renderer.c:307 */
		ts->completedSamples = r->prefs.sampleCount;
		ts->avgSampleTime = (long)(1e6 * (now_s() - t0) / ((double)r->prefs.sampleCount * got));
		for (int i = 0; i < got; ++i) {
			r->state.renderTiles[nums[i]].isRendering = false;                    /* renderer.c:315-316 */
			r->state.renderTiles[nums[i]].renderComplete = true;
			r->state.tileOwner[nums[i]] = ts->thread_num;
		}
		ts->currentTileNum = -1;
	}
	free(rects); free(nums);
	ts->threadComplete = true;                                    /* renderer.c:323 */
	ts->currentTileNum = -1;
	return NULL;
}

struct texture8 *renderFrame(struct renderer *r) {
	const int n = r->prefs.threadCount < 1 ? 1 : r->prefs.threadCount;
	const unsigned W = r->prefs.imageWidth, H = r->prefs.imageHeight;
	int ndev = 0;
	if (crgpu_device_count(&ndev) != CRGPU_OK || ndev < n) {
		fprintf(stderr, "renderFrame: %d GPU worker(s) requested, %d CUDA device(s) present: %s\n", n, ndev, crgpu_last_error());
		return NULL;
	}
	struct texture8 *output = calloc(1, sizeof *output);          /* renderer.c:41 */
	output->width = W; output->height = H;
	output->data = calloc((size_t)W * H * 3, 1);
	r->state.threads = calloc((size_t)n, sizeof *r->state.threads);
	r->state.threadStates = calloc((size_t)n, sizeof *r->state.threadStates);
	crgpu_scene **scenes = calloc((size_t)n, sizeof *scenes);
	for (int t = 0; t < n; ++t) {                                 /* one scene replica per GPU */
		if (crgpu_scene_create(&r->scene, t, &scenes[t]) != CRGPU_OK) {
			fprintf(stderr, "renderFrame: scene upload to device %d failed: %s\n", t, crgpu_last_error());
			for (int k = 0; k < t; ++k) crgpu_scene_destroy(scenes[k]);
			free(scenes); destroyTexture8(output);
			return NULL;
		}
	}
	r->state.isRendering = true;
	r->state.renderAborted = false;
	r->state.finishedTileCount = 0;
	for (int i = 0; i < r->state.tileCount; ++i) r->state.renderTiles[i].renderComplete = false;
	if (!r->prefs.quiet) printf("Rendering %ux%u, %d samples, %d bounces on %d GPU%s, %d tiles\n", W, H, r->prefs.sampleCount,
								r->prefs.bounces, n, n > 1 ? "s" : "", r->state.tileCount);
	const double t0 = now_s();
	for (int t = 0; t < n; ++t) {                                 /* renderer.c:97-105 */
		r->state.threadStates[t] = (struct renderThreadState){ .thread_num = t, .renderer = r, .gpu = scenes[t], .currentTileNum = -1 };
		if (pthread_create(&r->state.threads[t], NULL, gpuRenderThread, &r->state.threadStates[t]) == 0) r->state.activeThreads++;
	}
	while (r->state.isRendering) {                                /* renderer.c:122-172 */
		struct timespec ts = { 0, 16 * 1000 * 1000 };
		nanosleep(&ts, NULL);
		int done = 0;
		for (int t = 0; t < n; ++t) done += r->state.threadStates[t].threadComplete ? 1 : 0;
		if (done == n) r->state.isRendering = false;
	}
	for (int t = 0; t < n; ++t) pthread_join(r->state.threads[t], NULL);   /* renderer.c:175-177 */
	r->state.activeThreads = 0;
	int err = 0;
	r->state.totalRays = 0;
	for (int t = 0; t < n; ++t) { err |= r->state.threadStates[t].error; r->state.totalRays += r->state.threadStates[t].rays; }

	r->state.renderSeconds = now_s() - t0;
	if (!err) {
		if (n > 1) {
			/* every worker logged which tiles it rendered (state.tileOwner); gather them on device 0 */
			int *rects = malloc(sizeof(int) * 4 * (size_t)r->state.tileCount);
			int *owner = malloc(sizeof(int) * (size_t)r->state.tileCount);
			for (int i = 0; i < r->state.tileCount; ++i) {
				const struct renderTile *t = &r->state.renderTiles[i];
				rects[4 * i] = t->begin.x; rects[4 * i + 1] = t->begin.y; rects[4 * i + 2] = t->end.x; rects[4 * i + 3] = t->end.y;
				owner[i] = r->state.tileOwner[i];
			}
			crgpu_comm *comm = NULL;
			if (crhost_load_nccl() != 0 || g_nccl.create(scenes, n, &comm) != CRGPU_OK ||
				g_nccl.gather(comm, rects, owner, r->state.tileCount, 0) != CRGPU_OK) err = CRGPU_ERR_CUDA;
			if (comm) g_nccl.destroy(comm);
			free(rects); free(owner);
		}
		r->state.renderSeconds = now_s() - t0;
		if (!err && crgpu_framebuffer_read(scenes[0], r->state.renderBuffer, 0, 0, 0, 0) != CRGPU_OK) err = CRGPU_ERR_CUDA;   /* renderBuffer, scene.c:200 */
		if (!err && crgpu_framebuffer_to_srgb8(scenes[0], output->data) != CRGPU_OK) err = CRGPU_ERR_CUDA;            /* renderer.c:297-300 */
	}
	for (int t = 0; t < n; ++t) crgpu_scene_destroy(scenes[t]);
	free(scenes);
	if (!r->prefs.quiet && !err) {
		const double samples = (double)W * H * r->prefs.sampleCount;
		printf("Finished render in %.3f s: %.2f Msample/s, %.2f Mray/s (%llu rays)\n", r->state.renderSeconds,
			   samples / r->state.renderSeconds / 1e6, (double)r->state.totalRays / r->state.renderSeconds / 1e6, (unsigned long long)r->state.totalRays);
	}
	if (err) { destroyTexture8(output); return NULL; }
	return output;
}

void destroyTexture8(struct texture8 *t) { if (t) { free(t->data); free(t); } }

void destroyRenderer(struct renderer *r) {                       /* renderer.c:346-362 */
	if (!r) return;
	crscene_free(&r->scene);
	free(r->state.renderTiles); free(r->state.tileOwner); free(r->state.renderBuffer); free(r->state.threads); free(r->state.threadStates);
	pthread_mutex_destroy(&r->state.tileMutex);
	free(r);
}
