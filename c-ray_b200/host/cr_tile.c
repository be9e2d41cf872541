/*
 * cr_tile.c — image → tiles and the mutex-guarded tile queue.
 * Restates reference src/datatypes/tile.c: quantizeImage :66-117 (ragged edge tiles, end exclusive),
 * the five orderings :119-228 (fromMiddle alternates right/left of the middle, toMiddle alternates
 * last/first, topToBottom reverses, random = swap shuffle driven by PCG32 seeded 3141592 with rejection
 * sampling) and nextTile :22-45.  Built as an index permutation here instead of array copies.
 */
#include "cr_host.h"
#include <stdlib.h>
#include <string.h>

static uint32_t pcg32_next(uint64_t *state) {          /* pcg_basic.c:60-68, inc = 1 */
	uint64_t old = *state;
	*state = old * 6364136223846793005ULL + 1u;
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}

static unsigned rand_interval(unsigned lo, unsigned hi, uint64_t *rng) {   /* tile.c:131-146 */
	const unsigned range = 1 + hi - lo;
	const unsigned buckets = UINT32_MAX / range;
	const unsigned limit = buckets * range;
	unsigned r;
	do { r = pcg32_next(rng); } while (r >= limit);
	return lo + (r / buckets);
}

static void order_indices(unsigned *idx, unsigned n, enum renderOrder order) {
	for (unsigned i = 0; i < n; ++i) idx[i] = i;
	if (n == 0) return;
	unsigned *tmp = malloc(n * sizeof *tmp);
	switch (order) {
	case renderOrderTopToBottom:                          /* tile.c:119-129 */
		for (unsigned i = 0; i < n; ++i) tmp[i] = n - 1 - i;
		memcpy(idx, tmp, n * sizeof *tmp);
		break;
	case renderOrderFromMiddle: {                         /* tile.c:161-183 */
		int right = (int)(n / 2), left = right - 1;
		for (unsigned i = 0; i < n; ++i) tmp[i] = (i % 2 == 0) ? (unsigned)right++ : (unsigned)left--;
		memcpy(idx, tmp, n * sizeof *tmp);
		break;
	}
	case renderOrderToMiddle: {                           /* tile.c:185-206 */
		unsigned left = 0, right = n - 1;
		for (unsigned i = 0; i < n; ++i) tmp[i] = (i % 2 == 0) ? right-- : left++;
		memcpy(idx, tmp, n * sizeof *tmp);
		break;
	}
	case renderOrderRandom: {                             /* tile.c:148-159 */
		uint64_t rng = 0u;                                /* pcg32_srandom_r(&rng, 3141592, 0) */
		pcg32_next(&rng); rng += 3141592u; pcg32_next(&rng);
		for (unsigned i = 0; i < n; ++i) {
			unsigned j = rand_interval(0, n - 1, &rng);
			unsigned t = idx[i]; idx[i] = idx[j]; idx[j] = t;
		}
		break;
	}
	default: break;
	}
	free(tmp);
}

unsigned quantizeImage(struct renderTile **renderTiles, unsigned width, unsigned height,
					   unsigned tileWidth, unsigned tileHeight, enum renderOrder tileOrder) {
	if (tileWidth >= width) tileWidth = width;
	if (tileHeight >= height) tileHeight = height;
	if (tileWidth == 0) tileWidth = 1;
	if (tileHeight == 0) tileHeight = 1;
	const unsigned tilesX = (width + tileWidth - 1) / tileWidth;
	const unsigned tilesY = (height + tileHeight - 1) / tileHeight;
	const unsigned count = tilesX * tilesY;
	struct renderTile *grid = calloc(count ? count : 1, sizeof *grid);
	struct renderTile *out = calloc(count ? count : 1, sizeof *out);
	unsigned *idx = malloc((count ? count : 1) * sizeof *idx);
	if (!grid || !out || !idx) { free(grid); free(out); free(idx); *renderTiles = NULL; return 0; }
	for (unsigned y = 0; y < tilesY; ++y) {
		for (unsigned x = 0; x < tilesX; ++x) {
			struct renderTile *t = &grid[x + y * tilesX];
			t->begin.x = (int)(x * tileWidth);
			t->begin.y = (int)(y * tileHeight);
			t->end.x = (int)((x + 1) * tileWidth < width ? (x + 1) * tileWidth : width);
			t->end.y = (int)((y + 1) * tileHeight < height ? (y + 1) * tileHeight : height);
			t->width = (unsigned)(t->end.x - t->begin.x);
			t->height = (unsigned)(t->end.y - t->begin.y);
			t->tileNum = (int)(x + y * tilesX);
		}
	}
	order_indices(idx, count, tileOrder);
	for (unsigned i = 0; i < count; ++i) out[i] = grid[idx[i]];
	free(grid); free(idx);
	*renderTiles = out;
	return count;
}

struct renderTile nextTile(struct renderer *r) {
	struct renderTile tile;
	memset(&tile, 0, sizeof tile);
	tile.tileNum = -1;
	pthread_mutex_lock(&r->state.tileMutex);
	if (r->state.finishedTileCount < r->state.tileCount) {
		tile = r->state.renderTiles[r->state.finishedTileCount];
		r->state.renderTiles[r->state.finishedTileCount].isRendering = true;
		tile.tileNum = r->state.finishedTileCount++;
	}
	pthread_mutex_unlock(&r->state.tileMutex);
	return tile;
}
