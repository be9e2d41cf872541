/* cr_image.c — PNG and Radiance HDR decoders (see cr_image.h); zlib does the inflate. */
#include "cr_image.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
/* the legal (colour type, bit depth) pairs of PNG (spec 11.2.2): gray 1/2/4/8/16, RGB 8/16, palette 1/2/4/8, gray+alpha 8/16, RGBA 8/16 */
static int png_depth_ok(int ctype, int depth) {
	switch (ctype) {
	case 0: return depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16;
	case 3: return depth == 1 || depth == 2 || depth == 4 || depth == 8;
	case 2: case 4: case 6: return depth == 8 || depth == 16;
	default: return 0;
	}
}


static int paeth(int a, int b, int c) {
	int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	if (pa <= pb && pa <= pc) return a;
	return pb <= pc ? b : c;
}

int cr_image_decode_png(const unsigned char *buf, size_t len, struct cr_image *out) {
	static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
	memset(out, 0, sizeof *out);
	if (len < 8 || memcmp(buf, sig, 8)) return -1;
	size_t pos = 8;
	uint32_t W = 0, H = 0;
	int depth = 0, ctype = 0, interlace = 0, have_trns = 0, pal_n = 0;
	unsigned char pal[256][4];
	unsigned char key[3][2] = { { 0 } };
	unsigned char *z = NULL;
	size_t zlen = 0;
	for (int i = 0; i < 256; ++i) { pal[i][0] = pal[i][1] = pal[i][2] = 0; pal[i][3] = 255; }
	while (pos + 12 <= len) {
		const uint32_t n = be32(buf + pos);
		const unsigned char *tag = buf + pos + 4, *d = buf + pos + 8;
		if (pos + 12 + (size_t)n > len) { free(z); return -2; }
		if (!memcmp(tag, "IHDR", 4) && n >= 13) { W = be32(d); H = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; }
		else if (!memcmp(tag, "PLTE", 4)) { pal_n = (int)(n / 3); for (int i = 0; i < pal_n && i < 256; ++i) { pal[i][0] = d[3 * i]; pal[i][1] = d[3 * i + 1]; pal[i][2] = d[3 * i + 2]; } }
		else if (!memcmp(tag, "tRNS", 4)) {
			have_trns = 1;
			if (ctype == 3) { for (uint32_t i = 0; i < n && i < 256; ++i) pal[i][3] = d[i]; }
			else if (ctype == 0 && n >= 2) { key[0][0] = d[0]; key[0][1] = d[1]; }
			else if (ctype == 2 && n >= 6) { memcpy(key, d, 6); }
		}
		else if (!memcmp(tag, "IDAT", 4)) { z = realloc(z, zlen + n + 1); memcpy(z + zlen, d, n); zlen += n; }
		else if (!memcmp(tag, "IEND", 4)) break;
		pos += 12 + (size_t)n;
	}
	if (!W || !H || !z || interlace) { free(z); return -3; }             /* Adam7 is not supported */
	const int src_n = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!src_n || !png_depth_ok(ctype, depth)) { free(z); return -4; }
	const size_t bpp_bits = (size_t)src_n * (size_t)depth;
	const size_t stride = (W * bpp_bits + 7) / 8, fbpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
	uLongf rawlen = (uLongf)((stride + 1) * H);
	unsigned char *raw = malloc(rawlen);
	if (!raw || uncompress(raw, &rawlen, z, (uLong)zlen) != Z_OK || rawlen != (stride + 1) * H) { free(z); free(raw); return -5; }
	free(z);
	/* undo the scanline filters in place */
	for (uint32_t y = 0; y < H; ++y) {
		unsigned char *cur = raw + (size_t)y * (stride + 1) + 1;
		const unsigned char *up = y ? cur - (stride + 1) : NULL;
		const int f = cur[-1];
		for (size_t i = 0; i < stride; ++i) {
			const int a = i >= fbpp ? cur[i - fbpp] : 0, b = up ? up[i] : 0, c = (up && i >= fbpp) ? up[i - fbpp] : 0;
			int v = cur[i];
			switch (f) { case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break; default: break; }
			cur[i] = (unsigned char)v;
		}
	}
	int out_n = ctype == 3 ? (have_trns ? 4 : 3) : src_n + ((have_trns && (ctype == 0 || ctype == 2)) ? 1 : 0);
	unsigned char *px = malloc((size_t)W * H * (size_t)out_n);
	if (!px) { free(raw); return -6; }
	for (uint32_t y = 0; y < H; ++y) {
		const unsigned char *row = raw + (size_t)y * (stride + 1) + 1;
		unsigned char *dst = px + (size_t)y * W * (size_t)out_n;
		for (uint32_t x = 0; x < W; ++x) {
			unsigned s16[4] = { 0, 0, 0, 0 };
			unsigned char s8[4] = { 0, 0, 0, 0 };
			for (int c = 0; c < src_n; ++c) {
				if (depth == 16) { const unsigned char *p = row + ((size_t)x * src_n + c) * 2; s16[c] = ((unsigned)p[0] << 8) | p[1]; s8[c] = p[0]; }
				else if (depth == 8) { s8[c] = row[(size_t)x * src_n + c]; s16[c] = s8[c]; }
				else {
					const size_t bit = (size_t)x * (size_t)depth;
					const unsigned v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
					s16[c] = v;
					static const unsigned char scale[9] = { 0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01 };
					s8[c] = ctype == 3 ? (unsigned char)v : (unsigned char)(v * scale[depth]);
				}
			}
			if (ctype == 3) {
				const unsigned char *pe = pal[s8[0]];
				dst[0] = pe[0]; dst[1] = pe[1]; dst[2] = pe[2];
				if (out_n == 4) dst[3] = pe[3];
			} else {
				for (int c = 0; c < src_n; ++c) dst[c] = s8[c];
				if (out_n == src_n + 1) {
					int match = 1;
					for (int c = 0; c < src_n; ++c) {
						const unsigned k = depth == 16 ? (((unsigned)key[c][0] << 8) | key[c][1]) : (depth == 8 ? key[c][1] : key[c][1]);
						if (s16[c] != k) match = 0;
					}
					dst[src_n] = match ? 0 : 255;
				}
			}
			dst += out_n;
		}
	}
	free(raw);
	out->width = W; out->height = H; out->channels = (unsigned)out_n; out->is_float = 0; out->data = px;
	return 0;
}

/* ---- Radiance RGBE ------------------------------------------------------------------------------------------- */
int cr_path_is_hdr(const unsigned char *buf, size_t len) {
	return (len >= 11 && !memcmp(buf, "#?RADIANCE\n", 11)) || (len >= 7 && !memcmp(buf, "#?RGBE\n", 7));
}

static void rgbe_to_float(float *o, const unsigned char *rgbe) {          /* stbi__hdr_convert with 3 components */
	if (rgbe[3] != 0) {
		const float f1 = (float)ldexp(1.0f, rgbe[3] - (int)(128 + 8));
		o[0] = rgbe[0] * f1; o[1] = rgbe[1] * f1; o[2] = rgbe[2] * f1;
	} else { o[0] = o[1] = o[2] = 0.0f; }
}

static const unsigned char *read_line(const unsigned char *p, const unsigned char *end, char *line, size_t cap) {
	size_t n = 0;
	while (p < end && *p != '\n') { if (n + 1 < cap) line[n++] = (char)*p; p++; }
	line[n] = 0;
	return p < end ? p + 1 : p;
}

int cr_image_decode_hdr(const unsigned char *buf, size_t len, struct cr_image *out) {
	memset(out, 0, sizeof *out);
	const unsigned char *p = buf, *end = buf + len;
	char line[1024];
	p = read_line(p, end, line, sizeof line);
	if (strcmp(line, "#?RADIANCE") && strcmp(line, "#?RGBE")) return -1;
	int fmt = 0;
	for (;;) {
		if (p >= end) return -2;
		p = read_line(p, end, line, sizeof line);
		if (line[0] == 0) break;
		if (!strcmp(line, "FORMAT=32-bit_rle_rgbe")) fmt = 1;
	}
	if (!fmt) return -3;
	p = read_line(p, end, line, sizeof line);
	if (strncmp(line, "-Y ", 3)) return -4;
	char *tok = line + 3;
	const long H = strtol(tok, &tok, 10);
	while (*tok == ' ') tok++;
	if (strncmp(tok, "+X ", 3)) return -4;
	const long W = strtol(tok + 3, NULL, 10);
	if (W <= 0 || H <= 0) return -5;
	float *px = malloc(sizeof(float) * 3 * (size_t)W * (size_t)H);
	if (!px) return -6;
	int flat = (W < 8 || W >= 32768);
	unsigned char *scan = NULL;
	for (long j = 0; j < H && !flat; ++j) {
		if (end - p < 4) { free(px); free(scan); return -7; }
		const int c1 = p[0], c2 = p[1], l = p[2];
		if (c1 != 2 || c2 != 2 || (l & 0x80)) {
			if (j != 0) { free(px); free(scan); return -8; }           /* stb only falls back to flat data on the first scanline */
			flat = 1;
			break;
		}
		if (((l << 8) | p[3]) != W) { free(px); free(scan); return -9; }
		p += 4;
		if (!scan) scan = malloc((size_t)W * 4);
		for (int k = 0; k < 4; ++k) {
			long i = 0;
			while (i < W) {
				if (p >= end) { free(px); free(scan); return -7; }
				int count = *p++;
				if (count > 128) {
					count -= 128;
					if (p >= end || i + count > W) { free(px); free(scan); return -10; }
					const unsigned char v = *p++;
					for (int z = 0; z < count; ++z) scan[(i++) * 4 + k] = v;
				} else {
					if (end - p < count || i + count > W) { free(px); free(scan); return -10; }
					for (int z = 0; z < count; ++z) scan[(i++) * 4 + k] = *p++;
				}
			}
		}
		for (long i = 0; i < W; ++i) rgbe_to_float(px + ((size_t)j * (size_t)W + (size_t)i) * 3, scan + i * 4);
	}
	free(scan);
	if (flat) {
		if ((size_t)(end - p) < (size_t)W * (size_t)H * 4) { free(px); return -7; }
		for (size_t i = 0; i < (size_t)W * (size_t)H; ++i) rgbe_to_float(px + i * 3, p + i * 4);
	}
	out->width = (unsigned)W; out->height = (unsigned)H; out->channels = 3; out->is_float = 1; out->data = px;
	return 0;
}

/* Cheap plausibility check used before a decode is handed to a background thread: 1 when the header looks like something
 * the decoders accept (a full decode can still fail later, e.g. on a corrupt deflate stream). */
int cr_image_probe(const unsigned char *buf, size_t len) {
	if (cr_path_is_hdr(buf, len)) return 1;
	static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
	if (len < 8 + 12 + 13 || memcmp(buf, sig, 8) || memcmp(buf + 12, "IHDR", 4) || be32(buf + 8) < 13) return 0;
	const unsigned char *d = buf + 16;
	const int depth = d[8], ctype = d[9], interlace = d[12];
	const int src_n = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!be32(d) || !be32(d + 4) || interlace || !src_n) return 0;
	return png_depth_ok(ctype, depth);
}

int cr_image_decode(const unsigned char *buf, size_t len, struct cr_image *out) {
	return cr_path_is_hdr(buf, len) ? cr_image_decode_hdr(buf, len, out) : cr_image_decode_png(buf, len, out);
}

int cr_image_load(const char *path, struct cr_image *out) {
	memset(out, 0, sizeof *out);
	FILE *f = fopen(path, "rb");
	if (!f) return -100;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	rewind(f);
	unsigned char *buf = malloc(n > 0 ? (size_t)n : 1);
	if (!buf || fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return -101; }
	fclose(f);
	int rc = cr_image_decode(buf, (size_t)n, out);
	free(buf);
	return rc;
}
