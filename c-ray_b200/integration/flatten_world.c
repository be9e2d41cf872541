/*
 * flatten_world.c — REFERENCE-SIDE integration code: the file a c-ray maintainer adds to the reference
 * tree next to gpu_thread.c.  It compiles ONLY against the reference's own headers (src/datatypes/scene.h, renderer.h, ...):
 * it is not part of libcrgpu.so / libcrhost.so, and this repository builds it only inside oracle/Makefile's
 * `cray_ref_gpu` target, which compiles the unmodified reference sources where they lie.
 *
 * flatten_world(r, &flat): the loaded `struct world` (pointer-linked, src/datatypes/scene.h:14-39) re-indexed into the
 * pointer-free `struct crs_scene` of include/crscene.h.  Re-indexing ONLY: every float is copied bit for bit, BVHs node for
 * node, so GPU and CPU start from identical inputs.  The reference keeps its BVH and node structs file-private; the
 * accessors declared in crx.h (oracle/ref_shims/: one five-line getter per struct, appended to the reference's .c file through
 * an #include shim) stand where an in-tree integration would add the same getters next to each struct.
 */
#include "includes.h"
#include "renderer/renderer.h"
#include "datatypes/scene.h"
#include "datatypes/camera.h"
#include "datatypes/mesh.h"
#include "datatypes/sphere.h"
#include "datatypes/poly.h"
#include "datatypes/instance.h"
#include "datatypes/vertexbuffer.h"
#include "datatypes/image/texture.h"
#include "datatypes/material.h"
#include "accelerators/bvh.h"
#include "nodes/bsdfnode.h"
#include "nodes/vectornode.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crx.h"
#include "crscene.h"

static int g_flat_error;

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
static int flat_vector(const struct vectorNode *n);

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
	} else if (crx_is_blackbody(n, &o) || crx_is_combine_value(n, &o)) {
		int t = flat_value(o.in[0]);
		node_fill(idx, &o, t, -1, -1);
	} else if (crx_is_combine_rgb(n, &o)) {
		int r = flat_value(o.in[0]), g = flat_value(o.in[1]), b = flat_value(o.in[2]);
		node_fill(idx, &o, r, g, b);
	} else if (crx_is_vectocolor(n, &o)) {
		int v = flat_vector(o.in[0]);
		node_fill(idx, &o, v, -1, -1);
	} else {
		fprintf(stderr, "flatten_world: color node type not exportable\n"); g_flat_error = 1; return -1;
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
	} else if (crx_is_math(n, &o)) {
		int a = flat_value(o.in[0]), b = flat_value(o.in[1]);
		node_fill(idx, &o, a, b, -1);
	} else if (crx_is_fresnel(n, &o)) {
		int ior = flat_value(o.in[0]), nv = flat_vector(o.in[1]);
		node_fill(idx, &o, ior, nv, -1);
	} else if (crx_is_raylength(n, &o)) {
		node_fill(idx, &o, -1, -1, -1);
	} else {
		fprintf(stderr, "flatten_world: value node type not exportable\n"); g_flat_error = 1; return -1;
	}
	return idx;
}

static int flat_vector(const struct vectorNode *n) {
	if (!n) return -1;
	int idx = map_find(&g_nodes, n);
	if (idx >= 0) return idx;
	struct crx_nodeinfo o;
	idx = node_slot(n);
	if (crx_is_constant_vector(n, &o) || crx_is_normal(n, &o)) {
		node_fill(idx, &o, -1, -1, -1);
	} else if (crx_is_vecmath(n, &o)) {
		int a = flat_vector(o.in[0]), b = flat_vector(o.in[1]);
		node_fill(idx, &o, a, b, -1);
	} else {
		fprintf(stderr, "flatten_world: vector node type not exportable\n"); g_flat_error = 1; return -1;
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
		fprintf(stderr, "flatten_world: bsdf node type not exportable\n"); g_flat_error = 1; return -1;
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

int flatten_world(const struct renderer *r, struct crs_scene *s) {
	const struct world *w = r->scene;
	memset(s, 0, sizeof(*s));
	g_flat_error = 0;
	g_nodes.count = 0; g_texs.count = 0;
	g_flat_nodes = NULL;          /* ownership of the previous table went to the previous crs_scene */
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
			if (d->object >= (uint32_t)w->sphereCount) { fprintf(stderr, "flatten_world: volume instances are not exportable\n"); g_flat_error = 1; return -1; }
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
	return g_flat_error ? -1 : 0;
}
