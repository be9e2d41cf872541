/*
 * cr_image.h — texture file decoding for the scene loader: PNG (8/16-bit, gray / gray+alpha / RGB / RGBA / palette,
 * non-interlaced) and Radiance HDR (.hdr, RLE and flat).  Output conventions are the ones the reference gets from the
 * vendored stb_image 2.23 (src/utils/loaders/textureloader.c:34-87): channels as stored in the file (palette → RGB or
 * RGBA, tRNS colour key → extra alpha channel, 16-bit → high byte), rows top to bottom; HDR → 3 fp32 channels with
 * value = mantissa * 2^(exponent - 136).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

struct cr_image {
	unsigned width, height, channels;
	int is_float;
	void *data;        /* uint8_t[] or float[] */
};

/* returns 0 on success; *out->data is malloc'ed */
int cr_image_load(const char *path, struct cr_image *out);
int cr_image_decode_png(const unsigned char *buf, size_t len, struct cr_image *out);
int cr_image_decode_hdr(const unsigned char *buf, size_t len, struct cr_image *out);
int cr_path_is_hdr(const unsigned char *buf, size_t len);
int cr_image_decode(const unsigned char *buf, size_t len, struct cr_image *out);   /* picks the decoder from the content */
int cr_image_probe(const unsigned char *buf, size_t len);                          /* 1 = header plausible */
