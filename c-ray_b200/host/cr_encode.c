/*
 * cr_encode.c — 8-bit sRGB image writers (the step right after the hot path).
 * Same outputs as the reference's encoders: 24-bit bottom-up BMP (src/utils/encoders/formats/bmp.c:19-71)
 * and 8-bit RGB PNG (formats/png.c:24-75; multi-threaded zlib deflate here instead of the vendored lodepng,
 * the reference's tEXt labels included).  Pixel values come from crgpu_framebuffer_to_srgb8.
 */
#include "cr_host.h"
#include <sys/utsname.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <pthread.h>
#include <unistd.h>

int encodeBMP(const struct texture8 *img, const char *path) {
	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	const unsigned W = img->width, H = img->height;
	const unsigned rowBytes = (W * 3 + 3) & ~3u;
	const unsigned dataBytes = rowBytes * H;
	unsigned char hdr[54] = { 'B', 'M' };
	const uint32_t fileSize = 54 + dataBytes, off = 54, infoSize = 40, planesBpp = 1u | (24u << 16), ppm = 2835;
	memcpy(hdr + 2, &fileSize, 4); memcpy(hdr + 10, &off, 4); memcpy(hdr + 14, &infoSize, 4);
	memcpy(hdr + 18, &W, 4); memcpy(hdr + 22, &H, 4); memcpy(hdr + 26, &planesBpp, 4);
	memcpy(hdr + 34, &dataBytes, 4); memcpy(hdr + 38, &ppm, 4); memcpy(hdr + 42, &ppm, 4);
	fwrite(hdr, 1, 54, f);
	unsigned char *row = calloc(rowBytes, 1);
	for (unsigned y = 0; y < H; ++y) {                       /* BMP rows are bottom-up, BGR */
		const uint8_t *src = img->data + (size_t)(H - 1 - y) * W * 3;
		for (unsigned x = 0; x < W; ++x) { row[3 * x] = src[3 * x + 2]; row[3 * x + 1] = src[3 * x + 1]; row[3 * x + 2] = src[3 * x]; }
		fwrite(row, 1, rowBytes, f);
	}
	free(row);
	return fclose(f) == 0 ? 0 : -2;
}

static void put_chunk(FILE *f, const char *tag, const unsigned char *data, uint32_t len) {
	unsigned char b[4] = { (unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len };
	fwrite(b, 1, 4, f);
	fwrite(tag, 1, 4, f);
	if (len) fwrite(data, 1, len, f);
	uLong crc = crc32(0L, (const Bytef *)tag, 4);
	if (len) crc = crc32(crc, data, len);
	unsigned char c[4] = { (unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc };
	fwrite(c, 1, 4, f);
}

/* PNG, deflate in parallel.  The reference's lodepng encoder is single-threaded (an 8K frame takes seconds);
 * here the scanlines are cut into bands, every band is deflated by its own thread as a raw stream ending on a byte
 * boundary (Z_SYNC_FLUSH; the last band ends the stream with Z_FINISH), and the pieces are concatenated behind one zlib
 * header with the Adler-32 of the whole image combined from the per-band sums — a valid single zlib stream (the pigz
 * construction), written as one IDAT chunk per band. */
struct band { const struct texture8 *img; unsigned y0, y1; int last; unsigned char *out; size_t len; uLong adler; size_t raw; int rc; };
struct band_pool { struct band *bands; int n, next; };

static void deflate_band(struct band *b) {
	const unsigned W = b->img->width;
	const size_t stride = (size_t)W * 3 + 1, rawLen = stride * (b->y1 - b->y0);
	unsigned char *raw = malloc(rawLen ? rawLen : 1);
	b->rc = -3;
	if (!raw) return;
	for (unsigned y = b->y0; y < b->y1; ++y) {
		raw[(size_t)(y - b->y0) * stride] = 0;                 /* filter type 0 */
		memcpy(raw + (size_t)(y - b->y0) * stride + 1, b->img->data + (size_t)y * W * 3, (size_t)W * 3);
	}
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (deflateInit2(&zs, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { free(raw); return; }
	const size_t cap = deflateBound(&zs, (uLong)rawLen) + 64;
	b->out = malloc(cap);
	if (b->out) {
		zs.next_in = raw; zs.avail_in = (uInt)rawLen;
		zs.next_out = b->out; zs.avail_out = (uInt)cap;
		const int r = deflate(&zs, b->last ? Z_FINISH : Z_SYNC_FLUSH);
		if ((b->last && r == Z_STREAM_END) || (!b->last && r == Z_OK && zs.avail_in == 0)) {
			b->len = cap - zs.avail_out;
			b->adler = adler32(adler32(0L, Z_NULL, 0), raw, (uInt)rawLen);
			b->raw = rawLen;
			b->rc = 0;
		} else b->rc = -4;
	}
	deflateEnd(&zs);
	free(raw);
}

static void *band_worker(void *arg) {
	struct band_pool *p = arg;
	for (;;) {
		const int i = __atomic_fetch_add(&p->next, 1, __ATOMIC_RELAXED);
		if (i >= p->n) return NULL;
		deflate_band(&p->bands[i]);
	}
}

/* png.c:29-56: the reference labels its PNGs with uncompressed tEXt chunks (lodepng text_compression = 0): keyword, NUL, text */
static void put_text(FILE *f, const char *key, const char *text) {
	unsigned char buf[1600];
	const size_t k = strlen(key), t = strlen(text);
	if (k + 1 + t > sizeof buf) return;
	memcpy(buf, key, k); buf[k] = 0; memcpy(buf + k + 1, text, t);
	put_chunk(f, "tEXt", buf, (uint32_t)(k + 1 + t));
}
static void smart_time_us(double sec, char *buf, size_t n) {     /* timer.c smartTime */
	if (sec < 1.0) snprintf(buf, n, "%.0fms", 1e3 * sec);
	else if (sec < 60.0) snprintf(buf, n, "%.0fs", sec);
	else if (sec < 3600.0) snprintf(buf, n, "%dm %02ds", (int)(sec / 60.0), (int)sec % 60);
	else snprintf(buf, n, "%dh %02dm", (int)(sec / 3600.0), ((int)sec % 3600) / 60);
}

int encodePNG(const struct texture8 *img, const char *path) { return encodePNGInfo(img, path, NULL); }

int encodePNGInfo(const struct texture8 *img, const char *path, const struct renderInfo *info) {
	const unsigned W = img->width, H = img->height;
	if (!W || !H) return -1;
	/* bands of >= 1 MB of scanlines (smaller ones would cost compression ratio: each band starts with an empty window) */
	const size_t stride = (size_t)W * 3 + 1;
	unsigned rows = (unsigned)(((size_t)1 << 20) / stride) + 1;
	if ((size_t)rows * stride > 0x7ff00000u) rows = (unsigned)(0x7ff00000u / stride);   /* zlib's uInt lengths */
	if (rows < 1) return -3;
	const int n = (int)((H + rows - 1) / rows);
	struct band *bands = calloc((size_t)n, sizeof *bands);
	if (!bands) return -3;
	for (int i = 0; i < n; ++i) {
		bands[i].img = img;
		bands[i].y0 = (unsigned)i * rows;
		bands[i].y1 = bands[i].y0 + rows < H ? bands[i].y0 + rows : H;
		bands[i].last = i == n - 1;
	}
	struct band_pool pool = { bands, n, 0 };
	long cpus = sysconf(_SC_NPROCESSORS_ONLN);
	int threads = (int)(cpus < 1 ? 1 : cpus > 32 ? 32 : cpus);
	if (threads > n) threads = n;
	pthread_t th[32];
	int started = 0;
	for (int i = 1; i < threads; ++i) if (pthread_create(&th[started], NULL, band_worker, &pool) == 0) started++;
	band_worker(&pool);
	for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
	int rc = 0;
	for (int i = 0; i < n; ++i) if (bands[i].rc) rc = bands[i].rc;

	FILE *f = rc ? NULL : fopen(path, "wb");
	if (!rc && !f) rc = -1;
	if (!rc) {
		static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
		fwrite(sig, 1, 8, f);
		unsigned char ihdr[13] = { (unsigned char)(W >> 24), (unsigned char)(W >> 16), (unsigned char)(W >> 8), (unsigned char)W,
								   (unsigned char)(H >> 24), (unsigned char)(H >> 16), (unsigned char)(H >> 8), (unsigned char)H, 8, 2, 0, 0, 0 };
		put_chunk(f, "IHDR", ihdr, 13);
		if (info) {                                                /* same keywords as png.c:49-56; "Threads" counts GPU workers here */
			char num[64], sys[1400];
			put_text(f, "C-ray Version", "c-ray B200 path (libcrgpu.so / libcrhost.so), c-ray's API and file formats (c) 2015-2020 Valtteri Koskivuori");
			put_text(f, "C-ray Source", "https://github.com/vkoskiv/c-ray");
			snprintf(num, sizeof num, "%i", info->samples); put_text(f, "C-ray Samples", num);
			snprintf(num, sizeof num, "%i", info->bounces); put_text(f, "C-ray Bounces", num);
			smart_time_us(info->renderSeconds, num, sizeof num); put_text(f, "C-ray RenderTime", num);
			snprintf(num, sizeof num, "%i", info->threadCount); put_text(f, "C-ray Threads", num);
			struct utsname name;
			if (uname(&name) == 0) {
				snprintf(sys, sizeof sys, "%s %s %s %s %s", name.machine, name.nodename, name.release, name.sysname, name.version);
				put_text(f, "C-ray SysInfo", sys);
			}
		}
		uLong adler = adler32(0L, Z_NULL, 0);
		for (int i = 0; i < n; ++i) adler = i ? adler32_combine(adler, bands[i].adler, (z_off_t)bands[i].raw) : bands[i].adler;
		for (int i = 0; i < n; ++i) {
			/* first chunk carries the zlib header (deflate, 32K window, default level), the last one the Adler-32 */
			const size_t extra = (i == 0 ? 2 : 0) + (i == n - 1 ? 4 : 0);
			unsigned char *chunk = malloc(bands[i].len + extra);
			if (!chunk) { rc = -3; break; }
			size_t o = 0;
			if (i == 0) { chunk[o++] = 0x78; chunk[o++] = 0x9c; }
			memcpy(chunk + o, bands[i].out, bands[i].len); o += bands[i].len;
			if (i == n - 1) { chunk[o++] = (unsigned char)(adler >> 24); chunk[o++] = (unsigned char)(adler >> 16); chunk[o++] = (unsigned char)(adler >> 8); chunk[o++] = (unsigned char)adler; }
			put_chunk(f, "IDAT", chunk, (uint32_t)o);
			free(chunk);
		}
		put_chunk(f, "IEND", NULL, 0);
		if (fclose(f) != 0 && !rc) rc = -2;
	}
	for (int i = 0; i < n; ++i) free(bands[i].out);
	free(bands);
	return rc;
}

int writeImage(const struct texture8 *img, const char *path, enum fileType type) {   /* encoder.c:22-39 */
	return writeImageInfo(img, path, type, NULL);
}
int writeImageInfo(const struct texture8 *img, const char *path, enum fileType type, const struct renderInfo *info) {
	if (!img || !img->data || !path) return -1;
	return type == bmp ? encodeBMP(img, path) : encodePNGInfo(img, path, info);
}
void rendererInfo(const struct renderer *r, struct renderInfo *out) {            /* c-ray.c:88-95 fills struct renderInfo the same way */
	memset(out, 0, sizeof *out);
	if (!r) return;
	out->samples = r->prefs.sampleCount; out->bounces = r->prefs.bounces;
	out->threadCount = r->world > 1 ? r->world : r->prefs.threadCount;
	out->renderSeconds = r->state.renderSeconds;
}
