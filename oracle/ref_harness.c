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
}

/* ---- flatten ------------------------------------------------------------------------------------ */
struct ptrmap { const void **keys; int count, cap; };
static int map_find(struct ptrmap *m, const void *k) {
	for (int i = 0; i < m->count; ++i) if (m->keys[i] == k) return i;
	return -1;
}
static int map_add(struct ptrmap *m, const void *k) {
	if (m->count == m->cap) { m->cap = m->cap ? m->cap * 2 : 64; m->keys = realloc(m->keys, m->cap * sizeof(*m->keys)); }
	m->keys[m->count] = k;
	return m->count++;
}

static struct ptrmap g_nodes, g_texs;
static struct crs_node *g_flat_nodes;

static int flat_tex(const struct texture *t) {
	if (!t) return -1;
	int i = map_find(&g_texs, t);
	return i >= 0 ? i : map_add(&g_texs, t);
}

static int flat_value(const struct valueNode *n);
static int flat_color(const struct colorNode *n);

static int node_slot(const void *n) {
	int idx = map_add(&g_nodes, n);
	g_flat_nodes = realloc(g_flat_nodes, (size_t)g_nodes.count * sizeof(*g_flat_nodes));
	memset(&g_flat_nodes[idx], 0, sizeof(*g_flat_nodes));
	g_flat_nodes[idx].in[0] = g_flat_nodes[idx].in[1] = g_flat_nodes[idx].in[2] = -1;
	g_flat_nodes[idx].tex = -1;
	return idx;
}
static void node_fill(int idx, const struct crx_nodeinfo *o, int in0, int in1, int in2) {
	struct crs_node *d = &g_flat_nodes[idx];
	d->kind = o->kind;
	d->in[0] = in0; d->in[1] = in1; d->in[2] = in2;
	memcpy(d->f, o->f, sizeof(d->f));
	d->tex = flat_tex(o->tex);
	d->options = o->options;
}

static int flat_color(const struct colorNode *n) {
	if (!n) return -1;
	int idx = map_find(&g_nodes, n);
	if (idx >= 0) return idx;
	struct crx_nodeinfo o;
	idx = node_slot(n);
	if (crx_is_constant_color(n, &o) || crx_is_image(n, &o) || crx_is_gradient(n, &o)) {
		node_fill(idx, &o, -1, -1, -1);
	} else if (crx_is_checker(n, &o)) {
		int a = flat_color(o.in[0]), b = flat_color(o.in[1]), s = flat_value(o.in[2]);
		node_fill(idx, &o, a, b, s);
	} else if (crx_is_blackbody(n, &o)) {
		int t = flat_value(o.in[0]);
		node_fill(idx, &o, t, -1, -1);
	} else {
		fprintf(stderr, "ref_harness: color node type not exportable\n"); exit(4);
	}
	return idx;
}

static int flat_value(const struct valueNode *n) {
	if (!n) return -1;
	int idx = map_find(&g_nodes, n);
	if (idx >= 0) return idx;
	struct crx_nodeinfo o;
	idx = node_slot(n);
	if (crx_is_constant_value(n, &o)) {
		node_fill(idx, &o, -1, -1, -1);
	} else if (crx_is_grayscale(n, &o) || crx_is_alpha(n, &o)) {
		int c = flat_color(o.in[0]);
		node_fill(idx, &o, c, -1, -1);
	} else {
		fprintf(stderr, "ref_harness: value node type not exportable\n"); exit(4);
	}
	return idx;
}

static int flat_bsdf(const struct bsdfNode *n) {
	if (!n) return -1;
	int idx = map_find(&g_nodes, n);
	if (idx >= 0) return idx;
	struct crx_nodeinfo o;
	idx = node_slot(n);
	if (crx_is_diffuse(n, &o) || crx_is_transparent(n, &o) || crx_is_isotropic(n, &o)) {
		int c = flat_color(o.in[0]);
		node_fill(idx, &o, c, -1, -1);
	} else if (crx_is_metal(n, &o) || crx_is_emission(n, &o)) {
		int c = flat_color(o.in[0]), v = flat_value(o.in[1]);
		node_fill(idx, &o, c, v, -1);
	} else if (crx_is_glass(n, &o) || crx_is_background(n, &o)) {
		int c = flat_color(o.in[0]), v = flat_value(o.in[1]), w = flat_value(o.in[2]);
		node_fill(idx, &o, c, v, w);
	} else if (crx_is_plastic(n, &o)) {
		int c = flat_color(o.in[0]), r = flat_color(o.in[1]), d = flat_bsdf(o.in[2]);
		node_fill(idx, &o, c, r, d);
	} else if (crx_is_mix(n, &o)) {
		int a = flat_bsdf(o.in[0]), b = flat_bsdf(o.in[1]), f = flat_value(o.in[2]);
		node_fill(idx, &o, a, b, f);
	} else if (crx_is_add(n, &o)) {
		int a = flat_bsdf(o.in[0]), b = flat_bsdf(o.in[1]);
		node_fill(idx, &o, a, b, -1);
	} else {
		fprintf(stderr, "ref_harness: bsdf node type not exportable\n"); exit(4);
	}
	return idx;
}

static void flat_material(struct crs_material *d, const struct material *m) {
	d->emission[0] = m->emission.red; d->emission[1] = m->emission.green;
	d->emission[2] = m->emission.blue; d->emission[3] = m->emission.alpha;
	d->IOR = m->IOR;
	d->bsdf = flat_bsdf(m->bsdf);
}

static void flat_bvh(const struct bvh *b, struct crs_bvh *d, struct crs_bvh_node **nodes, uint32_t *nodeCount,
					 int32_t **prims, uint32_t *primCount, uint32_t nprims) {
	unsigned nc = crx_bvh_node_count(b);
	d->node_offset = *nodeCount; d->node_count = nc;
	d->prim_offset = *primCount; d->prim_count = nprims;
	*nodes = realloc(*nodes, ((size_t)*nodeCount + nc + 1) * sizeof(**nodes));
	/* struct bvhNode is {float bounds[6]; unsigned firstChildOrPrim; unsigned primCount:30; bool isLeaf:1;}
	   (bvh.c:37-42): with GCC's bit-field layout isLeaf lands in bit 30 of the last word, which is exactly
	   crs_bvh_node.prim_count_leaf, so a byte copy is the flattening. */
	if (nc) memcpy(*nodes + *nodeCount, crx_bvh_nodes(b), (size_t)nc * sizeof(**nodes));
	*nodeCount += nc;
	*prims = realloc(*prims, ((size_t)*primCount + nprims + 1) * sizeof(**prims));
	if (nprims) memcpy(*prims + *primCount, crx_bvh_prim_indices(b), (size_t)nprims * sizeof(**prims));
	*primCount += nprims;
}

static void flatten(const struct renderer *r, struct crs_scene *s) {
	const struct world *w = r->scene;
	memset(s, 0, sizeof(*s));
	s->prefs = (struct crs_prefs){
		.image_width = r->prefs.imageWidth, .image_height = r->prefs.imageHeight,
		.sample_count = (uint32_t)r->prefs.sampleCount, .bounces = (uint32_t)r->prefs.bounces,
		.tile_width = r->prefs.tileWidth, .tile_height = r->prefs.tileHeight,
		.tile_order = (uint32_t)r->prefs.tileOrder, .thread_count = (uint32_t)r->prefs.threadCount };
	const struct camera *c = w->camera;
	s->camera = (struct crs_camera){
		.sensor_x = c->sensorSize.x, .sensor_y = c->sensorSize.y, .aperture = c->aperture,
		.focal_distance = c->focalDistance,
		.forward = { c->forward.x, c->forward.y, c->forward.z },
		.right = { c->right.x, c->right.y, c->right.z },
		.up = { c->up.x, c->up.y, c->up.z },
		.width = c->width, .height = c->height };
	memcpy(s->camera.A, c->composite.A.mtx, sizeof(s->camera.A));

	/* materials: every mesh's material set, then one per sphere */
	uint32_t matCount = 0;
	for (int m = 0; m < w->meshCount; ++m) matCount += (uint32_t)w->meshes[m].materialCount;
	matCount += (uint32_t)w->sphereCount;
	s->materials = calloc(matCount + 1, sizeof(*s->materials));
	s->material_count = matCount;

	s->meshes = calloc((size_t)w->meshCount + 1, sizeof(*s->meshes));
	s->mesh_count = (uint32_t)w->meshCount;
	s->bvhs = calloc((size_t)w->meshCount + 2, sizeof(*s->bvhs));
	uint32_t polyCount = 0;
	for (int m = 0; m < w->meshCount; ++m) polyCount += (uint32_t)w->meshes[m].polyCount;
	s->polys = calloc((size_t)polyCount + 1, sizeof(*s->polys));
	s->poly_count = polyCount;

	uint32_t mat = 0, poly = 0;
	for (int m = 0; m < w->meshCount; ++m) {
		const struct mesh *mesh = &w->meshes[m];
		struct crs_mesh *d = &s->meshes[m];
		d->poly_offset = poly; d->poly_count = (uint32_t)mesh->polyCount;
		d->material_offset = mat; d->material_count = (uint32_t)mesh->materialCount;
		d->texcoord_count = (uint32_t)mesh->textureCoordCount;
		d->ray_offset = mesh->rayOffset;
		d->bvh = s->bvh_count;
		flat_bvh(mesh->bvh, &s->bvhs[s->bvh_count++], &s->bvh_nodes, &s->bvh_node_count,
				 &s->prim_indices, &s->prim_index_count, (uint32_t)mesh->polyCount);
		for (int p = 0; p < mesh->polyCount; ++p) {
			const struct poly *sp = &mesh->polygons[p];
			struct crs_poly *dp = &s->polys[poly + (uint32_t)p];
			for (int k = 0; k < 3; ++k) { dp->v[k] = sp->vertexIndex[k]; dp->n[k] = sp->normalIndex[k]; dp->t[k] = sp->textureIndex[k]; }
			dp->material = sp->materialIndex;
			dp->has_normals = sp->hasNormals ? 1u : 0u;
		}
		for (int k = 0; k < mesh->materialCount; ++k) flat_material(&s->materials[mat + (uint32_t)k], &mesh->materials[k]);
		poly += (uint32_t)mesh->polyCount;
		mat += (uint32_t)mesh->materialCount;
	}

	s->spheres = calloc((size_t)w->sphereCount + 1, sizeof(*s->spheres));
	s->sphere_count = (uint32_t)w->sphereCount;
	for (int i = 0; i < w->sphereCount; ++i) {
		s->spheres[i] = (struct crs_sphere){ .radius = w->spheres[i].radius, .ray_offset = w->spheres[i].rayOffset, .material = mat };
		flat_material(&s->materials[mat++], &w->spheres[i].material);
	}

	s->instances = calloc((size_t)w->instanceCount + 1, sizeof(*s->instances));
	s->instance_count = (uint32_t)w->instanceCount;
	for (int i = 0; i < w->instanceCount; ++i) {
		const struct instance *in = &w->instances[i];
		struct crs_instance *d = &s->instances[i];
		memcpy(d->A, in->composite.A.mtx, sizeof(d->A));
		memcpy(d->Ainv, in->composite.Ainv.mtx, sizeof(d->Ainv));
		if (isMesh(in)) {
			d->kind = CRS_INST_MESH;
			d->object = (uint32_t)((const struct mesh *)in->object - w->meshes);
		} else {
			/* the JSON loader only creates solid spheres and solid meshes (sceneloader.c:928,1086) */
			d->kind = CRS_INST_SPHERE;
			d->object = (uint32_t)((const struct sphere *)in->object - w->spheres);
			if (d->object >= (uint32_t)w->sphereCount) { fprintf(stderr, "ref_harness: volume instances are not exportable\n"); exit(4); }
		}
	}
	s->top_bvh = s->bvh_count;
	flat_bvh(w->topLevel, &s->bvhs[s->bvh_count++], &s->bvh_nodes, &s->bvh_node_count,
			 &s->prim_indices, &s->prim_index_count, (uint32_t)w->instanceCount);

	s->background = flat_bsdf(w->background);

	s->vertex_count = (uint32_t)vertexCount; s->normal_count = (uint32_t)normalCount; s->texcoord_count = (uint32_t)textureCount;
	s->vertices = (float *)g_vertices; s->normals = (float *)g_normals; s->texcoords = (float *)g_textureCoords;

	s->nodes = g_flat_nodes; s->node_count = (uint32_t)g_nodes.count;

	s->texture_count = (uint32_t)g_texs.count;
	s->textures = calloc((size_t)g_texs.count + 1, sizeof(*s->textures));
	uint64_t off = 0;
	for (int i = 0; i < g_texs.count; ++i) {
		const struct texture *t = g_texs.keys[i];
		uint64_t bytes = (uint64_t)t->width * t->height * t->channels * (t->precision == float_p ? 4u : 1u);
		s->textures[i] = (struct crs_texture){ .width = (uint32_t)t->width, .height = (uint32_t)t->height,
			.channels = (uint32_t)t->channels, .is_float = t->precision == float_p, .has_alpha = t->hasAlpha, .data_offset = off };
		off += (bytes + 15u) & ~(uint64_t)15u;
	}
	s->texdata_bytes = off;
	s->texdata = calloc(off + 16, 1);
	for (int i = 0; i < g_texs.count; ++i) {
		const struct texture *t = g_texs.keys[i];
		uint64_t bytes = (uint64_t)t->width * t->height * t->channels * (t->precision == float_p ? 4u : 1u);
		memcpy(s->texdata + s->textures[i].data_offset, t->data.byte_p, bytes);
	}
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
