/*
 * cr_encode.c — 8-bit sRGB image writers (the step right after the hot path).
 * Same outputs as the reference's encoders: 24-bit bottom-up BMP (src/utils/encoders/formats/bmp.c:19-71)
 * and 8-bit RGB PNG (formats/png.c:24-75; zlib deflate here instead of the vendored lodepng, tEXt
 * metadata omitted).  Pixel values come from crgpu_framebuffer_to_srgb8.
 */
#include "cr_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

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

int encodePNG(const struct texture8 *img, const char *path) {
	const unsigned W = img->width, H = img->height;
	const size_t rawLen = (size_t)H * ((size_t)W * 3 + 1);
	unsigned char *raw = malloc(rawLen);
	if (!raw) return -3;
	for (unsigned y = 0; y < H; ++y) {
		raw[(size_t)y * (W * 3 + 1)] = 0;                    /* filter type 0 */
		memcpy(raw + (size_t)y * (W * 3 + 1) + 1, img->data + (size_t)y * W * 3, (size_t)W * 3);
	}
	uLongf zlen = compressBound(rawLen);
	unsigned char *z = malloc(zlen);
	if (!z || compress2(z, &zlen, raw, rawLen, 6) != Z_OK) { free(raw); free(z); return -4; }
	free(raw);
	FILE *f = fopen(path, "wb");
	if (!f) { free(z); return -1; }
	static const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
	fwrite(sig, 1, 8, f);
	unsigned char ihdr[13] = { (unsigned char)(W >> 24), (unsigned char)(W >> 16), (unsigned char)(W >> 8), (unsigned char)W,
							   (unsigned char)(H >> 24), (unsigned char)(H >> 16), (unsigned char)(H >> 8), (unsigned char)H, 8, 2, 0, 0, 0 };
	put_chunk(f, "IHDR", ihdr, 13);
	put_chunk(f, "IDAT", z, (uint32_t)zlen);
	put_chunk(f, "IEND", NULL, 0);
	free(z);
	return fclose(f) == 0 ? 0 : -2;
}

int writeImage(const struct texture8 *img, const char *path, enum fileType type) {   /* encoder.c:22-39 */
	if (!img || !img->data || !path) return -1;
	return type == bmp ? encodeBMP(img, path) : encodePNG(img, path);
}
