/*
 * ref_harness.c — drives the UNMODIFIED c-ray reference (compiled from /root/reference by
 * oracle/Makefile into oracle/_ref/) as the parity oracle and CPU baseline.  TEST INFRASTRUCTURE ONLY:
 * nothing under c-ray_b200/ links or calls this.
 *
 * It replaces reference src/main.c:14-42 (same call sequence: crParseArgs → crInitRenderer →
 * crReadFile → crLoadSceneFromBuf → crStartRenderer) and adds what the reference cannot do itself:
 *   render  : dump the fp32 renderBuffer (the reference only writes 8-bit sRGB, c-ray.c:85-103)
 *   export  : flatten the loaded `struct world` into a .crscene file (include/crscene.h)
 *   hits    : known-answer vectors — camera ray, closest hit, bsdf sample for a grid of pixels
 *   kat     : known-answer vectors for initSampler/getDimension (hash64 + PCG32)
 */
#include "includes.h"
#include "c-ray.h"
#include "renderer/renderer.h"
#include "renderer/pathtrace.h"
#include "renderer/samplers/sampler.h"
#include "datatypes/scene.h"
#include "datatypes/camera.h"
#include "datatypes/mesh.h"
#include "datatypes/sphere.h"
#include "datatypes/poly.h"
#include "datatypes/instance.h"
#include "datatypes/vertexbuffer.h"
#include "datatypes/image/texture.h"
#include "datatypes/hitrecord.h"
#include "accelerators/bvh.h"
#include "nodes/bsdfnode.h"
#include "utils/logging.h"
#include "utils/args.h"

#include <stdio.h>
#include <string.h>
#include <float.h>
#include <time.h>

#include "ref_shims/crx.h"
#include "../include/crscene.h"

extern struct renderer *g_renderer;

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- node types no JSON scene can reach (math, vecmath, fresnel, raylength, normal, vectocolor, combine) --------
 * With CRAY_REF_F4 set, the materials of the loaded scene's spheres (and of its first mesh) are replaced by graphs built with the
 * reference's own constructors, so that the golden fixture g_f4 pins every mathOp / vecOp against the unmodified eval() code. */
#include "nodes/vectornode.h"
#include "nodes/converter/math.h"
#include "nodes/converter/vecmath.h"
#include "nodes/converter/vectocolor.h"
#include "nodes/converter/combine.h"
#include "nodes/converter/combinergb.h"
#include "nodes/converter/grayscale.h"
#include "nodes/input/fresnel.h"
#include "nodes/input/raylength.h"
#include "nodes/input/normal.h"
static void patch_f4_nodes(struct world *w) {
#define CV(x) newConstantValue(w, (x))
#define M2(a, b, op) newMath(w, (a), (b), (op))
#define M1(a, op) newMath(w, (a), NULL, (op))
#define VC(x, y, z) newConstantVector(w, (struct vector){ (x), (y), (z) })
#define V2(a, b, op) newVecMath(w, (a), (b), (op))
#define V1(a, op) newVecMath(w, (a), NULL, (op))
	const struct vectorNode *N = newNormal(w);
	const struct valueNode *RL = newRayLength(w);
	const struct valueNode *FR = newFresnel(w, CV(1.45f), N);
	const struct bsdfNode *g[8];
	g[0] = newDiffuse(w, newCombineRGB(w, M1(M1(M2(RL, CV(3.0f), Multiply), Sine), Absolute), FR, M2(CV(0.8f), CV(0.25f), Subtract)));
	g[1] = newMetal(w, newVecToColor(w, V1(N, VecAbs)), M2(FR, CV(0.3f), Multiply));
	g[2] = newDiffuse(w, newCombineValue(w, M2(M2(M2(M1(M1(RL, Cosine), Absolute), CV(2.0f), Power), CV(0.1f), Max), CV(0.95f), Min)));
	g[3] = newDiffuse(w, newCombineRGB(w, M2(M1(M1(M1(M2(RL, CV(0.5f), Multiply), Tangent), Absolute), SquareRoot), CV(1.0f), Min),
										 M1(CV(40.0f), ToRadians), M1(CV(0.01f), ToDegrees)));
	g[4] = newDiffuse(w, newVecToColor(w, V1(V1(V2(V2(N, VC(0.0f, 1.0f, 0.0f), VecCross), V2(N, VC(0.5f, 0.5f, 0.5f), VecMultiply), VecAdd), VecNormalize), VecAbs)));
	g[5] = newDiffuse(w, newVecToColor(w, V1(V2(V2(N, VC(0.0f, 1.0f, 0.0f), VecReflect), V2(N, VC(0.2f, 0.1f, 0.3f), VecSubtract), VecAverage), VecAbs)));
	g[6] = newDiffuse(w, newVecToColor(w, V2(V2(N, N, VecDot), V2(V1(N, VecLength), VC(0.3f, 0.6f, 0.9f), VecAdd), VecAdd)));
	g[7] = newGlass(w, newCombineValue(w, CV(0.95f)), M1(CV(1.02f), Log), M2(CV(3.0f), CV(2.0f), Divide));
	for (int i = 0; i < w->sphereCount && i < 8; ++i) w->spheres[i].material.bsdf = g[i];
	if (w->meshCount > 0)
		for (int m = 0; m < w->meshes[0].materialCount; ++m)
			w->meshes[0].materials[m].bsdf = newDiffuse(w, newCombineRGB(w, M2(CV(0.2f), M2(FR, CV(0.5f), Multiply), Add), CV(0.4f),
				newGrayscaleConverter(w, newVecToColor(w, V1(N, VecAbs)))));
#undef CV
#undef M2
#undef M1
#undef VC
#undef V2
#undef V1
}

/* ---- load through the reference's own API, with the CLI overrides of args.c:95-142 ------------- */
static void load_scene(const char *json, int W, int H, int spp, int bounces, int threads, int tw, int th) {
	char dims[64], tdims[64], sppstr[32], thr[32];
	snprintf(dims, sizeof dims, "%ix%i", W, H);
	snprintf(tdims, sizeof tdims, "%ix%i", tw, th);
	snprintf(sppstr, sizeof sppstr, "%i", spp);
	snprintf(thr, sizeof thr, "%i", threads);
	char a0[] = "c-ray", f1[] = "-d", f2[] = "-s", f3[] = "-j", f4[] = "-t";
	char *path = strdup(json);
	char *argv[16]; int argc = 0;
	argv[argc++] = a0; argv[argc++] = path;
	if (W > 0 && H > 0) { argv[argc++] = f1; argv[argc++] = dims; }
	if (spp > 0) { argv[argc++] = f2; argv[argc++] = sppstr; }
	if (threads > 0) { argv[argc++] = f3; argv[argc++] = thr; }
	if (tw > 0 && th > 0) { argv[argc++] = f4; argv[argc++] = tdims; }
	argv[argc] = NULL;
	crParseArgs(argc, argv);
	crInitRenderer();
	size_t bytes = 0;
	char *input = crReadFile(&bytes);
	if (!input) { fprintf(stderr, "ref_harness: cannot read %s\n", json); exit(2); }
	if (crLoadSceneFromBuf(input) != 0) { fprintf(stderr, "ref_harness: scene load failed\n"); exit(3); }
	free(input);
	if (bounces > 0) g_renderer->prefs.bounces = bounces; /* no CLI flag for bounces (args.c:28-44) */
	if (getenv("CRAY_REF_F4")) patch_f4_nodes(g_renderer->scene);
}

/* ---- flatten: c-ray_b200/integration/flatten_world.c (the reference-side integration code; linked into this harness) ---- */
int flatten_world(const struct renderer *r, struct crs_scene *s);
static void flatten(const struct renderer *r, struct crs_scene *s) {
	if (flatten_world(r, s) != 0) { fprintf(stderr, "ref_harness: scene not exportable\n"); exit(4); }
}

/* ---- known-answer vectors ------------------------------------------------------------------------- */
struct hit_kat {          /* 40 floats/ints, little endian */
	int32_t x, y, pixIdx, instIndex;
	int32_t polyIndex;    /* global poly index (mesh poly_offset + index in mesh) or -1 */
	float o[3], d[3];     /* camera ray (world) */
	float distance, uv[2];
	float hitPoint[3], normal[3];
	float emission[3];
	float out[3], color[4]; /* bsdf sample (or background sample on a miss) */
	float nextDraw;       /* one getDimension() after the bsdf sample: pins the number of draws consumed */
	float pad[9];
};

static int global_poly_index(const struct world *w, const struct poly *p) {
	if (!p) return -1;
	int off = 0;
	for (int m = 0; m < w->meshCount; ++m) {
		const struct mesh *mesh = &w->meshes[m];
		if (p >= mesh->polygons && p < mesh->polygons + mesh->polyCount) return off + (int)(p - mesh->polygons);
		off += mesh->polyCount;
	}
	return -2;
}

static void cmd_hits(const char *out, int n) {
	struct renderer *r = g_renderer;
	const struct world *w = r->scene;
	FILE *f = fopen(out, "wb");
	if (!f) { perror(out); exit(5); }
	int W = (int)r->prefs.imageWidth, H = (int)r->prefs.imageHeight;
	sampler *smp = newSampler();
	for (int i = 0; i < n; ++i) {
		/* low-discrepancy-ish pixel walk that covers the whole frame */
		int x = (int)(((uint64_t)i * 2654435761u) % (uint64_t)W);
		int y = (int)(((uint64_t)i * 40503u + 7u) % (uint64_t)H);
		struct hit_kat k; memset(&k, 0, sizeof k);
		k.x = x; k.y = y; k.pixIdx = y * W + x;
		initSampler(smp, Random, i % r->prefs.sampleCount, r->prefs.sampleCount, (uint32_t)k.pixIdx);
		struct lightRay ray = getCameraRay(w->camera, x, y, smp);
		k.o[0] = ray.start.x; k.o[1] = ray.start.y; k.o[2] = ray.start.z;
		k.d[0] = ray.direction.x; k.d[1] = ray.direction.y; k.d[2] = ray.direction.z;
		struct hitRecord isect = { .incident = ray, .instIndex = -1, .distance = FLT_MAX, .polygon = NULL };
		traverseTopLevelBvh(w->instances, w->topLevel, &ray, &isect, smp);
		k.instIndex = isect.instIndex;
		struct bsdfSample s;
		if (isect.instIndex < 0) {
			k.polyIndex = -1;
			s = w->background->sample(w->background, smp, &isect);
		} else {
			k.polyIndex = global_poly_index(w, isect.polygon);
			k.distance = isect.distance; k.uv[0] = isect.uv.x; k.uv[1] = isect.uv.y;
			k.hitPoint[0] = isect.hitPoint.x; k.hitPoint[1] = isect.hitPoint.y; k.hitPoint[2] = isect.hitPoint.z;
			k.normal[0] = isect.surfaceNormal.x; k.normal[1] = isect.surfaceNormal.y; k.normal[2] = isect.surfaceNormal.z;
			k.emission[0] = isect.material.emission.red; k.emission[1] = isect.material.emission.green; k.emission[2] = isect.material.emission.blue;
			s = isect.material.bsdf->sample(isect.material.bsdf, smp, &isect);
		}
		k.out[0] = s.out.x; k.out[1] = s.out.y; k.out[2] = s.out.z;
		k.color[0] = s.color.red; k.color[1] = s.color.green; k.color[2] = s.color.blue; k.color[3] = s.color.alpha;
		k.nextDraw = getDimension(smp);
		fwrite(&k, sizeof k, 1, f);
	}
	destroySampler(smp);
	fclose(f);
}

static void cmd_kat(const char *out) {
	/* (pixIdx, pass, maxPasses) triples incl. the 32-bit wrap of sampler.c:42 */
	static const uint32_t cases[][3] = {
		{0, 0, 1}, {1, 0, 1}, {12345, 7, 16}, {63999, 15, 16}, {2073599, 999, 1000},
		{2073599, 2499, 2500}, {33177599, 3999, 4000}, {4294967295u, 1, 2}, {777, 3, 250}, {1, 1, 100} };
	FILE *f = fopen(out, "wb");
	if (!f) { perror(out); exit(5); }
	sampler *smp = newSampler();
	for (size_t c = 0; c < sizeof cases / sizeof cases[0]; ++c) {
		initSampler(smp, Random, (int)cases[c][1], (int)cases[c][2], cases[c][0]);
		fwrite(cases[c], sizeof(uint32_t), 3, f);
		for (int i = 0; i < 13; ++i) { float v = getDimension(smp); fwrite(&v, 4, 1, f); }
	}
	destroySampler(smp);
	fclose(f);
}

static void usage(void) {
	fprintf(stderr,
		"usage: cray_ref render <json> <W> <H> <spp> <bounces> <threads> <tileW> <tileH> <out.f32>\n"
		"       cray_ref export <json> <W> <H> <spp> <bounces> <out.crscene>\n"
		"       cray_ref hits   <json> <W> <H> <spp> <bounces> <count> <out.bin>\n"
		"       cray_ref kat    <out.bin>\n"
		"(W/H/spp/bounces/threads/tile <= 0 keep the JSON value)\n");
	exit(1);
}

int main(int argc, char **argv) {
	if (argc < 2) usage();
	if (!strcmp(argv[1], "kat") && argc == 3) { cmd_kat(argv[2]); return 0; }
	if (!strcmp(argv[1], "render") && argc == 11) {
		load_scene(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]));
		double t0 = now_s();
		crStartRenderer();
		double t1 = now_s();
		const struct texture *rb = g_renderer->state.renderBuffer;
		FILE *f = fopen(argv[10], "wb");
		if (!f) { perror(argv[10]); return 5; }
		fwrite(rb->data.float_p, sizeof(float), rb->width * rb->height * rb->channels, f);
		fclose(f);
		printf("\nREF_RENDER width=%zu height=%zu spp=%d bounces=%d threads=%d seconds=%.6f\n",
			   rb->width, rb->height, g_renderer->prefs.sampleCount, g_renderer->prefs.bounces,
			   g_renderer->prefs.threadCount, t1 - t0);
		return 0;
	}
	if (!strcmp(argv[1], "export") && argc == 8) {
		load_scene(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), 1, 0, 0);
		struct crs_scene s;
		flatten(g_renderer, &s);
		if (crscene_save(&s, argv[7]) != 0) { fprintf(stderr, "ref_harness: cannot write %s\n", argv[7]); return 5; }
		printf("\nREF_EXPORT instances=%u spheres=%u meshes=%u polys=%u bvh_nodes=%u nodes=%u textures=%u texbytes=%llu\n",
			   s.instance_count, s.sphere_count, s.mesh_count, s.poly_count, s.bvh_node_count, s.node_count,
			   s.texture_count, (unsigned long long)s.texdata_bytes);
		return 0;
	}
	if (!strcmp(argv[1], "hits") && argc == 9) {
		load_scene(argv[2], atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), 1, 0, 0);
		cmd_hits(argv[8], atoi(argv[7]));
		return 0;
	}
	usage();
	return 1;
}
