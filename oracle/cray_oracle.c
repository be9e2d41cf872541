/*
 * cray_oracle.c — CPU restatement of c-ray's hot loop (see cray_oracle.h).  Plain C99, compiled with
 * -ffp-contract=off so every a*b+c rounds twice like the strict reference build.  Each function cites
 * the reference file:line it restates.  Written against the flat scene (include/crscene.h), not the
 * reference's structs; no reference source is included or linked.
 */
#include "cray_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <pthread.h>

#define PI 3.141592653589793238462643383279502f /* includes.h:13 */

typedef struct { float x, y, z; } v3;
typedef struct { float x, y; } v2;
typedef struct { float r, g, b, a; } col;

/* includes.h:20-21 — ternary macros; operand order and NaN behaviour matter */
#define MIN_(a, b) (((a) < (b)) ? (a) : (b))
#define MAX_(a, b) (((a) > (b)) ? (a) : (b))

/* ---- vector.h ------------------------------------------------------------------------------------ */
static inline v3 v3add(v3 a, v3 b) { return (v3){ a.x + b.x, a.y + b.y, a.z + b.z }; }          /* :65 */
static inline v3 v3sub(v3 a, v3 b) { return (v3){ a.x - b.x, a.y - b.y, a.z - b.z }; }          /* :76 */
static inline v3 v3mul(v3 a, v3 b) { return (v3){ a.x * b.x, a.y * b.y, a.z * b.z }; }          /* :80 */
static inline float v3dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             /* :91 */
static inline v3 v3scale(v3 v, float c) { return (v3){ v.x * c, v.y * c, v.z * c }; }           /* :102 */
static inline v3 v3cross(v3 a, v3 b) {                                                          /* :121 */
	return (v3){ (a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x) };
}
static inline float v3len(v3 v) { return sqrtf(v3dot(v, v)); }                                  /* :162 */
static inline v3 v3norm(v3 v) { float l = v3len(v); return (v3){ v.x / l, v.y / l, v.z / l }; } /* :173 */
static inline v3 v3neg(v3 v) { return (v3){ -v.x, -v.y, -v.z }; }                               /* :200 */
static inline v3 v3reflect(v3 I, v3 N) { return v3sub(I, v3scale(N, v3dot(N, I) * 2.0f)); }     /* :211 */
static inline float clampf(float value, float lo, float hi) { return MIN_(MAX_(value, lo), hi); } /* :55 */
static inline float wrapMax(float x, float max) { return fmodf(max + fmodf(x, max), max); }     /* :215 */
static inline float wrapMinMax(float x, float min, float max) { return min + wrapMax(x - min, max - min); } /* :219 */

/* ---- color.h ------------------------------------------------------------------------------------- */
static inline col cmul(col a, col b) { return (col){ a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a }; }   /* :33 */
static inline col cadd(col a, col b) { return (col){ a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a }; }   /* :38 */
static inline col ccoef(float k, col c) { return (col){ c.r * k, c.g * k, c.b * k, c.a * k }; }        /* :49 */
static inline col cmix(col a, col b, float t) { return cadd(ccoef(1.0f - t, a), ccoef(t, b)); }        /* :54 */
static inline float srgb_to_linear(float c) {                                                           /* :68 */
	if (c <= 0.04045f) return c / 12.92f;
	return powf(((c + 0.055f) / 1.055f), 2.4f);
}
static inline float linear_to_srgb(float c) {                                                           /* :60 */
	if (c <= 0.0031308f) return 12.92f * c;
	return (1.055f * powf(c, 0.4166666667f)) - 0.055f;
}
static inline col to_grayscale(col c) {                                                                 /* :43 — note the double constants */
	float b = sqrtf(0.299f * powf(c.r, 2) + 0.587 * powf(c.g, 2) + 0.114 * powf(c.b, 2));
	return (col){ b, b, b, c.a };
}
static col color_for_kelvin(float kelvin) {                                                             /* color.c:28-70 */
	col ret = { 0 };
	float temp = kelvin >= 40000.0f ? 40000.0f : kelvin;
	temp = temp / 100.0f;
	if (temp <= 66.0f) {
		ret.r = 255.0f;
	} else {
		ret.r = temp - 60.0f;
		ret.r = 329.698727446f * powf(ret.r, -0.1332047592f);
		ret.r = ret.r < 0.0f ? 0.0f : ret.r;
		ret.r = ret.r > 255.0f ? 255.0f : ret.r;
	}
	if (temp <= 66.0f) {
		ret.g = temp;
		ret.g = 99.4708025861f * logf(ret.g) - 161.1195681661f;
		ret.g = ret.g < 0.0f ? 0.0f : ret.g;
		ret.g = ret.g > 255.0f ? 255.0f : ret.g;
	} else {
		ret.g = temp - 60.0f;
		ret.g = 288.1221695283f * powf(ret.g, -0.0755148492f);
		ret.g = ret.g < 0.0f ? 0.0f : ret.g;
		ret.g = ret.g > 255.0f ? 255.0f : ret.g;
	}
	if (temp >= 66.0f) {
		ret.b = 255.0f;
	} else {
		if (temp <= 19.0f) {
			ret.b = 0.0f;
		} else {
			ret.b = temp - 10.0f;
			ret.b = 138.5177312231f * logf(ret.b) - 305.0447927307f;
			ret.b = ret.b < 0.0f ? 0.0f : ret.b;
			ret.b = ret.b > 255.0f ? 255.0f : ret.b;
		}
	}
	return (col){ ret.r / 255.0f, ret.g / 255.0f, ret.b / 255.0f, 0 };
}

/* ---- sampler: common.h:22-27, pcg_basic.c:42-68, random.c:16-21 ------------------------------------ */
struct rng { uint64_t state; struct cro_counters *ctr; }; /* inc is always 1 (initseq 0) */

static inline uint64_t hash64(uint64_t x) {
	x = (x ^ (x >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
	x = (x ^ (x >> 27)) * UINT64_C(0x94d049bb133111eb);
	x = x ^ (x >> 31);
	return x;
}
static inline uint32_t pcg32(struct rng *r) {
	uint64_t old = r->state;
	r->state = old * 6364136223846793005ULL + 1u;
	uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
	uint32_t rot = (uint32_t)(old >> 59u);
	return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}
static inline void rng_init(struct rng *r, uint32_t pixIdx, int pass, int maxPasses) { /* sampler.c:41-44 */
	uint32_t seed32 = pixIdx * (uint32_t)maxPasses + (uint32_t)pass; /* 32-bit wrap is reference behaviour */
	uint64_t seed = hash64((uint64_t)seed32);
	r->state = 0u;
	pcg32(r);
	r->state += seed;
	pcg32(r);
}
static inline float draw(struct rng *r) {
	if (r->ctr) r->ctr->draws++;
	return (1.0f / (1ull << 32)) * pcg32(r); /* may equal 1.0f */
}

static inline v3 random_on_unit_sphere(struct rng *r) {                                       /* vector.h:243 */
	const float sx = draw(r);
	const float sy = draw(r);
	const float a = sx * (2.0f * PI);
	const float s = 2.0f * sqrtf(MAX_(0.0f, sy * (1.0f - sy)));
	return (v3){ cosf(a) * s, sinf(a) * s, 1.0f - 2.0f * sy };
}

/* ---- transforms.c:76-116 ----------------------------------------------------------------------------- */
static inline v3 xf_point(const float *m, v3 v) {
	return (v3){ (m[0] * v.x) + (m[1] * v.y) + (m[2] * v.z) + m[3],
				 (m[4] * v.x) + (m[5] * v.y) + (m[6] * v.z) + m[7],
				 (m[8] * v.x) + (m[9] * v.y) + (m[10] * v.z) + m[11] };
}
static inline v3 xf_vector(const float *m, v3 v) {
	return (v3){ (m[0] * v.x) + (m[1] * v.y) + (m[2] * v.z),
				 (m[4] * v.x) + (m[5] * v.y) + (m[6] * v.z),
				 (m[8] * v.x) + (m[9] * v.y) + (m[10] * v.z) };
}
static inline v3 xf_vector_transpose(const float *m, v3 v) {                                   /* :106-111 */
	return (v3){ (m[0] * v.x) + (m[4] * v.y) + (m[8] * v.z),
				 (m[1] * v.x) + (m[5] * v.y) + (m[9] * v.z),
				 (m[2] * v.x) + (m[6] * v.y) + (m[10] * v.z) };
}

/* ---- hit record (hitrecord.h:14-23, flattened) ------------------------------------------------------ */
struct ray { v3 o, d; };
struct hit {
	struct ray incident;
	int material;      /* global material index (the reference copies the 144-byte struct) */
	v3 p, n;
	v2 uv;
	float dist;
	int poly;          /* global poly index or -1 */
	int inst;
};

/* ---- camera.c:50-87 ------------------------------------------------------------------------------------ */
static inline float signf_(float v) { return (v >= 0.0f) ? 1.0f : -1.0f; }
static inline float triangle_distribution(float v) {
	const float orig = v * 2.0f - 1.0f;
	v = orig / sqrtf(fabsf(orig));
	v = clampf(v, -1.0f, 1.0f);
	v = v - signf_(orig);
	return v;
}
static struct ray camera_ray(const struct crs_camera *cam, int x, int y, struct rng *r) {
	struct ray ray = { { 0, 0, 0 }, { 0, 0, 0 } };
	const v3 right = { cam->right[0], cam->right[1], cam->right[2] };
	const v3 up = { cam->up[0], cam->up[1], cam->up[2] };
	const v3 forward = { cam->forward[0], cam->forward[1], cam->forward[2] };
	const float jitterX = triangle_distribution(draw(r));
	const float jitterY = triangle_distribution(draw(r));
	v3 pixX = v3scale(right, (cam->sensor_x / cam->width));
	v3 pixY = v3scale(up, (cam->sensor_y / cam->height));
	v3 pixV = v3add(forward, v3add(v3scale(pixX, x - cam->width * 0.5f + jitterX + 0.5f),
								   v3scale(pixY, y - cam->height * 0.5f + jitterY + 0.5f)));
	ray.d = v3norm(pixV);
	if (cam->aperture > 0.0f) {
		float ft = cam->focal_distance / v3dot(ray.d, forward);
		v3 focus = v3add(ray.o, v3scale(ray.d, ft));
		float rr = sqrtf(draw(r));                           /* randomCoordOnUnitDisc, vector.h:194-198 */
		float theta = ((draw(r)) * (2.0f * PI - 0.0f)) + 0.0f; /* rndFloatRange(0, 2π) */
		v2 lens = { (rr * cosf(theta)) * cam->aperture, (rr * sinf(theta)) * cam->aperture };
		ray.o = v3add(ray.o, v3add(v3scale(right, lens.x), v3scale(up, lens.y)));
		ray.d = v3norm(v3sub(focus, ray.o));
	}
	ray.o = xf_point(cam->A, ray.o);
	ray.d = xf_vector(cam->A, ray.d);
	return ray;
}

/* ---- poly.c:17-53 ------------------------------------------------------------------------------------------ */
static inline v3 vtx(const float *a, int i) { return (v3){ a[3 * (size_t)i], a[3 * (size_t)i + 1], a[3 * (size_t)i + 2] }; }

static bool ray_triangle(const struct crs_scene *s, const struct ray *ray, int polyIdx, struct hit *isect) {
	const struct crs_poly *p = &s->polys[polyIdx];
	v3 v0 = vtx(s->vertices, p->v[0]), v1 = vtx(s->vertices, p->v[1]), v2_ = vtx(s->vertices, p->v[2]);
	v3 e1 = v3sub(v0, v1);
	v3 e2 = v3sub(v2_, v0);
	v3 n = v3cross(e1, e2);
	v3 c = v3sub(v0, ray->o);
	v3 r = v3cross(ray->d, c);
	float invDet = 1.0f / v3dot(n, ray->d);
	float u = v3dot(r, e2) * invDet;
	float v = v3dot(r, e1) * invDet;
	float w = 1.0f - u - v;
	if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) {
		float t = v3dot(n, c) * invDet;
		if (t >= 0.0f && t < isect->dist) {
			isect->uv = (v2){ u, v };
			isect->dist = t;
			if (p->has_normals) {
				v3 up = v3scale(vtx(s->normals, p->n[1]), u);
				v3 vp = v3scale(vtx(s->normals, p->n[2]), v);
				v3 wp = v3scale(vtx(s->normals, p->n[0]), w);
				isect->n = v3add(v3add(up, vp), wp);
			} else {
				isect->n = n;
			}
			isect->p = v3add(ray->o, v3scale(ray->d, t));
			return true;
		}
	}
	return false;
}

/* ---- sphere.c:20-61 ------------------------------------------------------------------------------------------ */
static bool ray_sphere(const struct ray *ray, float radius, struct hit *isect) {
	float A = v3dot(ray->d, ray->d);
	float B = 2.0f * v3dot(ray->d, ray->o);
	float C = v3dot(ray->o, ray->o) - (radius * radius);
	float disc = B * B - 4.0f * A * C;
	if (disc < 0.0f) return false;
	float sq = sqrtf(disc);
	float t0 = (-B + sq) / 2.0f;
	float t1 = (-B - sq) / 2.0f;
	if (t0 > t1 && t1 > 0.0f) t0 = t1;
	if (t0 < 0.00001f || t0 > isect->dist) return false;
	isect->dist = t0;
	isect->p = v3add(ray->o, v3scale(ray->d, isect->dist));
	isect->n = v3norm(isect->p);
	isect->poly = -1;
	return true;
}

/* ---- bvh.c:318-441 -------------------------------------------------------------------------------------------- */
static inline bool node_test(const struct crs_bvh_node *node, v3 invDir, v3 scaledStart, const int *oct,
							 float maxDist, float *tEntry) {
	/* strict reference build: FP_FAST_FMAF undefined → plain a*b+c (bvh.c:318-324) */
	float tMinX = node->bounds[0 + oct[0]] * invDir.x + scaledStart.x;
	float tMaxX = node->bounds[0 + 1 - oct[0]] * invDir.x + scaledStart.x;
	float tMinY = node->bounds[2 + oct[1]] * invDir.y + scaledStart.y;
	float tMaxY = node->bounds[2 + 1 - oct[1]] * invDir.y + scaledStart.y;
	float tMinZ = node->bounds[4 + oct[2]] * invDir.z + scaledStart.z;
	float tMaxZ = node->bounds[4 + 1 - oct[2]] * invDir.z + scaledStart.z;
	float tMin = tMinX > tMinY ? tMinX : tMinY;
	float tMax = tMaxX < tMaxY ? tMaxX : tMaxY;
	tMin = tMin > tMinZ ? tMin : tMinZ;
	tMax = tMax < tMaxZ ? tMax : tMaxZ;
	tMin = tMin > 0 ? tMin : 0;
	tMax = tMax < maxDist ? tMax : maxDist;
	*tEntry = tMin;
	return tMin <= tMax;
}

/* optional event trace of the traversal, at the granularity of the GPU kernel's state machine (crgpu_trace.cuh): used by
 * tools/k2_warp_model.py to study warp scheduling policies offline.  Event bytes: 0 = top-level child-pair step,
 * 1 = mesh instance step, 2 = sphere instance step, 64+k = bottom-level child-pair step that tested k triangles (k <= 63). */
struct cro_trace { uint8_t *buf; size_t cap, len; size_t pair_at; uint16_t x, y, pass; };
struct trav_ctx { const struct crs_scene *s; struct cro_counters *ctr; struct cro_trace *tr; };
static inline void tr_event(const struct trav_ctx *c, uint8_t code) {
	if (c->tr && c->tr->len < c->tr->cap) { c->tr->pair_at = c->tr->len; c->tr->buf[c->tr->len++] = code; }
}
static inline void tr_tri(const struct trav_ctx *c, size_t at) {
	if (c->tr && at < c->tr->cap && c->tr->buf[at] < 127) c->tr->buf[at]++;
}

static bool leaf_bottom(const struct trav_ctx *c, const struct crs_mesh *mesh, const struct crs_bvh *bvh,
						const struct crs_bvh_node *leaf, const struct ray *ray, struct hit *isect) { /* bvh.c:443-462 */
	bool found = false;
	int count = (int)(leaf->prim_count_leaf & CRS_BVH_COUNT_MASK);
	for (int i = 0; i < count; ++i) {
		int local = c->s->prim_indices[bvh->prim_offset + leaf->first_child_or_prim + (uint32_t)i];
		int polyIdx = (int)mesh->poly_offset + local;
		if (c->ctr) c->ctr->tri_tests++;
		if (c->tr) tr_tri(c, c->tr->pair_at);
		if (ray_triangle(c->s, ray, polyIdx, isect)) {
			isect->poly = polyIdx;
			found = true;
		}
	}
	return found;
}

static bool traverse(const struct trav_ctx *c, uint32_t bvhIdx, int meshIdx /* -1: top level */,
					 const struct ray *ray, struct hit *isect);

static v2 texmap_sphere(v3 ud) {                                                              /* instance.c:33-43 */
	float phi = atan2f(ud.z, ud.x);
	float theta = asinf(ud.y);
	float v = (theta + PI / 2.0f) / PI;
	float u = 1.0f - (phi + PI) / (PI * 2.0f);
	u = wrapMinMax(u, 0.0f, 1.0f);
	v = wrapMinMax(v, 0.0f, 1.0f);
	return (v2){ u, v };
}

static bool intersect_sphere_inst(const struct trav_ctx *c, const struct crs_instance *inst,
								  const struct ray *ray, struct hit *isect) {                  /* instance.c:45-60 */
	const struct crs_sphere *sp = &c->s->spheres[inst->object];
	struct ray copy = { xf_point(inst->Ainv, ray->o), xf_vector(inst->Ainv, ray->d) };
	copy.o = v3add(copy.o, v3scale(copy.d, sp->ray_offset));
	if (c->ctr) c->ctr->sphere_tests++;
	tr_event(c, 2);
	if (ray_sphere(&copy, sp->radius, isect)) {
		isect->uv = texmap_sphere(isect->n);
		isect->poly = -1;
		isect->material = (int)sp->material;
		isect->p = xf_point(inst->A, isect->p);
		isect->n = xf_vector_transpose(inst->Ainv, isect->n);
		return true;
	}
	return false;
}

static v2 texmap_mesh(const struct crs_scene *s, const struct crs_mesh *mesh, const struct hit *isect) { /* instance.c:150-167 */
	if (mesh->texcoord_count == 0) return (v2){ -1.0f, -1.0f };
	const struct crs_poly *p = &s->polys[isect->poly];
	if (p->t[0] == -1) return (v2){ -1.0f, -1.0f };
	const float u = isect->uv.x;
	const float v = isect->uv.y;
	const float w = 1.0f - u - v;
	const float *tc = s->texcoords;
	v2 uc = { tc[2 * (size_t)p->t[1]] * u, tc[2 * (size_t)p->t[1] + 1] * u };
	v2 vc = { tc[2 * (size_t)p->t[2]] * v, tc[2 * (size_t)p->t[2] + 1] * v };
	v2 wc = { tc[2 * (size_t)p->t[0]] * w, tc[2 * (size_t)p->t[0] + 1] * w };
	return (v2){ (uc.x + vc.x) + wc.x, (uc.y + vc.y) + wc.y };
}

static bool intersect_mesh_inst(const struct trav_ctx *c, const struct crs_instance *inst,
								const struct ray *ray, struct hit *isect) {                    /* instance.c:169-185 */
	const struct crs_mesh *mesh = &c->s->meshes[inst->object];
	struct ray copy = { xf_point(inst->Ainv, ray->o), xf_vector(inst->Ainv, ray->d) };
	copy.o = v3add(copy.o, v3scale(copy.d, mesh->ray_offset));
	if (c->ctr) c->ctr->inst_visits++;
	tr_event(c, 1);
	if (traverse(c, mesh->bvh, (int)inst->object, &copy, isect)) {
		isect->uv = texmap_mesh(c->s, mesh, isect);
		isect->material = (int)(mesh->material_offset + c->s->polys[isect->poly].material);
		isect->p = xf_point(inst->A, isect->p);
		isect->n = xf_vector_transpose(inst->Ainv, isect->n);
		isect->n = v3norm(isect->n);
		return true;
	}
	return false;
}

static bool leaf_top(const struct trav_ctx *c, const struct crs_bvh *bvh, const struct crs_bvh_node *leaf,
					 const struct ray *ray, struct hit *isect) {                               /* bvh.c:468-486 */
	bool found = false;
	int count = (int)(leaf->prim_count_leaf & CRS_BVH_COUNT_MASK);
	for (int i = 0; i < count; ++i) {
		int cur = c->s->prim_indices[bvh->prim_offset + leaf->first_child_or_prim + (uint32_t)i];
		const struct crs_instance *inst = &c->s->instances[cur];
		bool h = inst->kind == CRS_INST_MESH ? intersect_mesh_inst(c, inst, ray, isect)
											 : intersect_sphere_inst(c, inst, ray, isect);
		if (h) {
			isect->inst = cur;
			found = true;
		}
	}
	return found;
}

#define MAX_BVH_DEPTH 64

static bool traverse(const struct trav_ctx *c, uint32_t bvhIdx, int meshIdx, const struct ray *ray, struct hit *isect) {
	const struct crs_bvh *bvh = &c->s->bvhs[bvhIdx];
	const struct crs_bvh_node *nodes = c->s->bvh_nodes + bvh->node_offset;
	const struct crs_mesh *mesh = meshIdx >= 0 ? &c->s->meshes[meshIdx] : NULL;
#define LEAF(nd) (mesh ? leaf_bottom(c, mesh, bvh, (nd), ray, isect) : leaf_top(c, bvh, (nd), ray, isect))
	if (bvh->node_count < 1) {                                                              /* bvh.c:362-365 */
		isect->inst = -1;
		return false;
	}
	const struct crs_bvh_node *stack[MAX_BVH_DEPTH + 1];
	int stackSize = 0;
	int oct[3] = { signbit(ray->d.x) ? 1 : 0, signbit(ray->d.y) ? 1 : 0, signbit(ray->d.z) ? 1 : 0 };
	v3 invDir = { 1.0f / ray->d.x, 1.0f / ray->d.y, 1.0f / ray->d.z };
	v3 scaledStart = v3scale(v3mul(ray->o, invDir), -1.0f);
	float maxDist = isect->dist;

	if (bvh->node_count == 1) {                                                             /* bvh.c:382-387 */
		float tEntry;
		if (mesh) tr_event(c, 64);                               /* single-leaf BVH: one step that tests the leaf's triangles */
		if (node_test(nodes, invDir, scaledStart, oct, maxDist, &tEntry)) return LEAF(nodes);
		return false;
	}
	const struct crs_bvh_node *node = nodes;
	bool hasHit = false;
	while (true) {
		unsigned firstChild = node->first_child_or_prim;
		const struct crs_bvh_node *left = &nodes[firstChild];
		const struct crs_bvh_node *right = &nodes[firstChild + 1];
		if (c->ctr) c->ctr->node_pairs++;
		tr_event(c, mesh ? 64 : 0);
		float tL, tR;
		bool hitL = node_test(left, invDir, scaledStart, oct, maxDist, &tL);
		bool hitR = node_test(right, invDir, scaledStart, oct, maxDist, &tR);
		if (hitL) {
			if (left->prim_count_leaf & CRS_BVH_LEAF_BIT) {
				if (LEAF(left)) { maxDist = isect->dist; hasHit = true; }
				left = NULL;
			}
		} else left = NULL;
		if (hitR) {
			if (right->prim_count_leaf & CRS_BVH_LEAF_BIT) {
				if (LEAF(right)) { maxDist = isect->dist; hasHit = true; }
				right = NULL;
			}
		} else right = NULL;
		if ((right != NULL) & (left != NULL)) {
			if (tL > tR) { node = left; left = right; right = node; }
			node = left;
			stack[stackSize++] = right;
			if (c->ctr && (uint64_t)stackSize > c->ctr->max_stack) c->ctr->max_stack = (uint64_t)stackSize;
		} else if ((right != NULL) ^ (left != NULL)) {
			node = right != NULL ? right : left;
		} else {
			if (stackSize == 0) break;
			node = stack[--stackSize];
		}
	}
#undef LEAF
	return hasHit;
}

/* ---- texture.c:32-79 ---------------------------------------------------------------------------------------------- */
/* (size_t)float as GCC emits it on x86-64 (cvttss2si): negative values wrap modulo 2^64 */
static inline uint64_t f2sz(float x) {
	if (x < 9223372036854775808.0f) return (uint64_t)(int64_t)x;
	return ((uint64_t)(int64_t)(x - 9223372036854775808.0f)) ^ (UINT64_C(1) << 63);
}

static col texel(const struct crs_scene *s, const struct crs_texture *t, uint64_t x, uint64_t y) {
	col o = { 0, 0, 0, 0 };
	x = x % t->width;
	y = y % t->height;
	const uint8_t *base = s->texdata + t->data_offset;
	size_t idx = (size_t)((x + ((t->height - 1) - y) * t->width) * t->channels);
	if (t->channels == 1) {
		if (t->is_float) o.r = ((const float *)base)[idx];
		else o.r = base[idx] / 255.0f;
		o.g = o.r; o.b = o.r; o.a = 1.0f;
	} else if (t->is_float) {
		const float *f = (const float *)base;
		o.r = f[idx]; o.g = f[idx + 1]; o.b = f[idx + 2];
		o.a = t->has_alpha ? f[idx + 3] : 1.0f;
	} else {
		o.r = base[idx] / 255.0f; o.g = base[idx + 1] / 255.0f; o.b = base[idx + 2] / 255.0f;
		o.a = t->has_alpha ? base[idx + 3] / 255.0f : 1.0f;
	}
	return o;
}

static col texture_get(const struct crs_scene *s, const struct crs_texture *t, float x, float y, bool filtered) {
	if (!filtered) return texel(s, t, f2sz(x), f2sz(y));
	x = x * t->width;
	y = y * t->height;
	float xcopy = x - 0.5f;
	float ycopy = y - 0.5f;
	int xint = (int)xcopy;
	int yint = (int)ycopy;
	col tl = texel(s, t, (uint64_t)(int64_t)xint, (uint64_t)(int64_t)yint);
	col tr = texel(s, t, (uint64_t)(int64_t)(xint + 1), (uint64_t)(int64_t)yint);
	col bl = texel(s, t, (uint64_t)(int64_t)xint, (uint64_t)(int64_t)(yint + 1));
	col br = texel(s, t, (uint64_t)(int64_t)(xint + 1), (uint64_t)(int64_t)(yint + 1));
	return cmix(cmix(tl, tr, xcopy - xint), cmix(bl, br, xcopy - xint), ycopy - yint);
}

/* ---- nodes ------------------------------------------------------------------------------------------------------------ */
static float eval_value(const struct crs_scene *s, int node, const struct hit *rec);
static inline float schlick(float cosine, float IOR);

/* vector nodes (vectornode.c:38-42, input/normal.c:36-40, converter/vecmath.c:43-83): struct vectorValue's .v member; the ops that
 * produce the scalar .f (dot, length) leave .v zero-initialised, and no node of the reference reads .f */
static v3 eval_vector(const struct crs_scene *s, int node, const struct hit *rec) {
	const struct crs_node *n = &s->nodes[node];
	switch (n->kind) {
	case CRS_VECTOR_CONSTANT: return (v3){ n->f[0], n->f[1], n->f[2] };
	case CRS_VECTOR_NORMAL: return rec->n;
	case CRS_VECTOR_VECMATH: {
		const v3 a = eval_vector(s, n->in[0], rec);
		const v3 b = eval_vector(s, n->in[1], rec);
		switch (n->options) {
		case CRS_VEC_ADD: return v3add(a, b);
		case CRS_VEC_SUBTRACT: return v3sub(a, b);
		case CRS_VEC_MULTIPLY: return (v3){ a.x * b.x, a.y * b.y, a.z * b.z };
		case CRS_VEC_AVERAGE: return v3scale(v3add(a, b), 0.5f);
		case CRS_VEC_CROSS: return v3cross(a, b);
		case CRS_VEC_NORMALIZE: return v3norm(a);
		case CRS_VEC_REFLECT: return v3reflect(a, b);
		case CRS_VEC_ABS: return (v3){ fabsf(a.x), fabsf(a.y), fabsf(a.z) };
		default: return (v3){ 0, 0, 0 };                     /* VecDot, VecLength: .f only */
		}
	}
	default: return (v3){ 0, 0, 0 };
	}
}

static col eval_color(const struct crs_scene *s, int node, const struct hit *rec) {
	const struct crs_node *n = &s->nodes[node];
	switch (n->kind) {
	case CRS_COLOR_CONSTANT:                                                          /* constant.c:39-42 */
		return (col){ n->f[0], n->f[1], n->f[2], n->f[3] };
	case CRS_COLOR_IMAGE: {                                                           /* image.c:31-48 */
		if (n->tex < 0) return (col){ 1.0f, 0.0f, 0.5f, 1.0f };
		const struct crs_texture *t = &s->textures[n->tex];
		col out;
		if (n->options & CRS_IMG_NO_BILINEAR) {
			float x = rec->uv.x * t->width;
			float y = rec->uv.y * t->height;
			out = texture_get(s, t, x, y, false);
		} else {
			out = texture_get(s, t, rec->uv.x, rec->uv.y, true);
		}
		if (n->options & CRS_IMG_SRGB_TRANSFORM)
			out = (col){ srgb_to_linear(out.r), srgb_to_linear(out.g), srgb_to_linear(out.b), out.a };
		return out;
	}
	case CRS_COLOR_CHECKER: {                                                         /* checker.c:31-54 */
		const float coef = eval_value(s, n->in[2], rec);
		float sines;
		if (rec->uv.x >= 0) sines = sinf(coef * rec->uv.x) * sinf(coef * rec->uv.y);
		else sines = sinf(coef * rec->p.x) * sinf(coef * rec->p.y) * sinf(coef * rec->p.z);
		return sines < 0.0f ? eval_color(s, n->in[0], rec) : eval_color(s, n->in[1], rec);
	}
	case CRS_COLOR_GRADIENT: {                                                        /* gradient.c:40-45 */
		v3 unit = v3norm(rec->incident.d);
		float t = 0.5f * (unit.y + 1.0f);
		col down = { n->f[0], n->f[1], n->f[2], n->f[3] }, up = { n->f[4], n->f[5], n->f[6], n->f[7] };
		return cadd(ccoef(1.0f - t, down), ccoef(t, up));
	}
	case CRS_COLOR_BLACKBODY:                                                         /* blackbody.c:38-42 */
		return color_for_kelvin(eval_value(s, n->in[0], rec));
	case CRS_COLOR_VECTOCOLOR: {                                                      /* vectocolor.c:38-43 */
		const v3 v = eval_vector(s, n->in[0], rec);
		return (col){ v.x, v.y, v.z, 0.0f };
	}
	case CRS_COLOR_COMBINE_VALUE: {                                                   /* combine.c:38-43 */
		const float v = eval_value(s, n->in[0], rec);
		return (col){ v, v, v, 1.0f };
	}
	case CRS_COLOR_COMBINE_RGB: {                                                     /* combinergb.c:44-53: R, G, B in this order */
		const float r = eval_value(s, n->in[0], rec);
		const float g = eval_value(s, n->in[1], rec);
		const float b = eval_value(s, n->in[2], rec);
		return (col){ r, g, b, 1.0f };
	}
	default:
		return (col){ 0, 0, 0, 1 };
	}
}

static float eval_value(const struct crs_scene *s, int node, const struct hit *rec) {
	const struct crs_node *n = &s->nodes[node];
	switch (n->kind) {
	case CRS_VALUE_CONSTANT: return n->f[0];
	case CRS_VALUE_GRAYSCALE: return to_grayscale(eval_color(s, n->in[0], rec)).r;     /* grayscale.c:40-43 */
	case CRS_VALUE_ALPHA: return eval_color(s, n->in[0], rec).a;                        /* alpha.c:38-41 */
	case CRS_VALUE_MATH: {                                                              /* math.c:44-97 */
		const float a = eval_value(s, n->in[0], rec);
		const float b = eval_value(s, n->in[1], rec);
		switch (n->options) {
		case CRS_MATH_ADD: return a + b;
		case CRS_MATH_SUBTRACT: return a - b;
		case CRS_MATH_MULTIPLY: return a * b;
		case CRS_MATH_DIVIDE: return a / b;
		case CRS_MATH_POWER: return powf(a, b);
		case CRS_MATH_LOG: return log10f(a);
		case CRS_MATH_SQRT: return sqrtf(a);
		case CRS_MATH_ABS: return fabsf(a);
		case CRS_MATH_MIN: return MIN_(a, b);
		case CRS_MATH_MAX: return MAX_(a, b);
		case CRS_MATH_SINE: return sinf(a);
		case CRS_MATH_COSINE: return cosf(a);
		case CRS_MATH_TANGENT: return tanf(a);
		case CRS_MATH_TO_RADIANS: return (a * PI) / 180.0f;                             /* transforms.c:18-20 */
		case CRS_MATH_TO_DEGREES: return a * (180.0f / PI);                             /* transforms.c:22-24 */
		default: return 0.0f;
		}
	}
	case CRS_VALUE_FRESNEL: {                                                           /* fresnel.c:43-55: IOR is evaluated twice */
		const float IOR = eval_value(s, n->in[0], rec);
		float cosine;
		const float dn = v3dot(rec->incident.d, rec->n);
		if (dn > 0.0f) cosine = IOR * v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d);
		else cosine = -(v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d));
		return schlick(cosine, eval_value(s, n->in[0], rec));
	}
	case CRS_VALUE_RAYLENGTH: return rec->dist;                                         /* raylength.c:36-40 */
	default: return 0.0f;
	}
}

struct bsdf_sample { v3 out; col color; };

static bool refract_(v3 in, v3 normal, float niOverNt, v3 *refracted) {                 /* vector.h:252-266 */
	const v3 uv = v3norm(in);
	const float dt = v3dot(uv, normal);
	const float disc = 1.0f - niOverNt * niOverNt * (1.0f - dt * dt);
	if (disc > 0.0f) {
		const v3 A = v3scale(normal, dt);
		const v3 B = v3sub(uv, A);
		const v3 C = v3scale(B, niOverNt);
		const v3 D = v3scale(normal, sqrtf(disc));
		*refracted = v3sub(C, D);
		return true;
	}
	return false;
}
static inline float schlick(float cosine, float IOR) {                                  /* vector.h:268-272 */
	float r0 = (1.0f - IOR) / (1.0f + IOR);
	r0 = r0 * r0;
	return r0 + (1.0f - r0) * powf((1.0f - cosine), 5.0f);
}

static struct bsdf_sample sample_bsdf(const struct crs_scene *s, int node, struct rng *r, struct hit *rec) {
	const struct crs_node *n = &s->nodes[node];
	switch (n->kind) {
	case CRS_BSDF_DIFFUSE: {                                                          /* diffuse.c:40-47 */
		const v3 dir = v3norm(v3add(rec->n, random_on_unit_sphere(r)));
		return (struct bsdf_sample){ dir, eval_color(s, n->in[0], rec) };
	}
	case CRS_BSDF_ISOTROPIC: {                                                        /* isotropic.c:40-47 */
		const v3 dir = v3norm(random_on_unit_sphere(r));
		return (struct bsdf_sample){ dir, eval_color(s, n->in[0], rec) };
	}
	case CRS_BSDF_EMISSIVE: {                                                         /* emission.c:42-49 */
		const v3 dir = v3norm(v3add(rec->n, random_on_unit_sphere(r)));
		float strength = eval_value(s, n->in[1], rec);
		return (struct bsdf_sample){ dir, ccoef(strength, eval_color(s, n->in[0], rec)) };
	}
	case CRS_BSDF_METAL: {                                                            /* metal.c:40-55 */
		const v3 nd = v3norm(rec->incident.d);
		v3 reflected = v3reflect(nd, rec->n);
		float rough = eval_value(s, n->in[1], rec);
		if (rough > 0.0f) reflected = v3add(reflected, v3scale(random_on_unit_sphere(r), rough));
		return (struct bsdf_sample){ reflected, eval_color(s, n->in[0], rec) };
	}
	case CRS_BSDF_GLASS: {                                                            /* glass.c:41-87 */
		/* glass.c:47 declares `refracted` uninitialised and reads it (:76-80) when refract() failed (total internal reflection)
		 * and the draw is exactly 1.0f (129 of 2^32 draws): the compiled reference then scatters along (0, 0, <stale stack word>),
		 * which is not a function of the path.  DEFINED here (DESIGN.md deviation #3, identical in crgpu_shade.cuh): total
		 * internal reflection always reflects, i.e. `refracted` starts out as `reflected`. */
		v3 outward;
		v3 reflected = v3reflect(rec->incident.d, rec->n);
		v3 refracted = reflected;
		float niOverNt, prob, cosine;
		float IOR = eval_value(s, n->in[2], rec);
		if (v3dot(rec->incident.d, rec->n) > 0.0f) {
			outward = v3neg(rec->n);
			niOverNt = IOR;
			cosine = IOR * v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d);
		} else {
			outward = rec->n;
			niOverNt = 1.0f / IOR;
			cosine = -(v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d));
		}
		if (refract_(rec->incident.d, outward, niOverNt, &refracted)) prob = schlick(cosine, IOR);
		else prob = 1.0f;
		float rough = eval_value(s, n->in[1], rec);
		if (rough > 0.0f) {
			v3 fuzz = v3scale(random_on_unit_sphere(r), rough);
			reflected = v3add(reflected, fuzz);
			refracted = v3add(refracted, fuzz);
		}
		v3 dir = draw(r) < prob ? reflected : refracted;
		return (struct bsdf_sample){ dir, eval_color(s, n->in[0], rec) };
	}
	case CRS_BSDF_PLASTIC: {                                                          /* plastic.c:42-87 */
		v3 outward, refracted;
		float niOverNt, prob, cosine;
		const float IOR = s->materials[rec->material].IOR;
		if (v3dot(rec->incident.d, rec->n) > 0.0f) {
			outward = v3neg(rec->n);
			niOverNt = IOR;
			cosine = IOR * v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d);
		} else {
			outward = rec->n;
			niOverNt = 1.0f / IOR;
			cosine = -(v3dot(rec->incident.d, rec->n) / v3len(rec->incident.d));
		}
		if (refract_(rec->incident.d, outward, niOverNt, &refracted)) prob = schlick(cosine, IOR);
		else prob = 1.0f;
		if (draw(r) < prob) {                                                         /* sampleShiny :42-56 */
			v3 reflected = v3reflect(rec->incident.d, rec->n);
			float rough = eval_color(s, n->in[1], rec).r;
			if (rough > 0.0f) reflected = v3add(reflected, v3scale(random_on_unit_sphere(r), rough));
			return (struct bsdf_sample){ reflected, (col){ 1.0f, 1.0f, 1.0f, 1.0f } };
		}
		return sample_bsdf(s, n->in[2], r, rec);
	}
	case CRS_BSDF_MIX: {                                                              /* mix.c:42-50 */
		const float lerp = eval_value(s, n->in[2], rec);
		if (draw(r) > lerp) return sample_bsdf(s, n->in[0], r, rec);
		return sample_bsdf(s, n->in[1], r, rec);
	}
	case CRS_BSDF_ADD: {                                                              /* add.c:42-49 */
		struct bsdf_sample A = sample_bsdf(s, n->in[0], r, rec);
		struct bsdf_sample B = sample_bsdf(s, n->in[1], r, rec);
		return (struct bsdf_sample){ v3add(A.out, B.out), cadd(A.color, B.color) };
	}
	case CRS_BSDF_TRANSPARENT:                                                        /* transparent.c:40-44 */
		return (struct bsdf_sample){ rec->incident.d, eval_color(s, n->in[0], rec) };
	case CRS_BSDF_BACKGROUND: {                                                       /* background.c:39-66 */
		v3 ud = v3norm(rec->incident.d);
		float rr = 1.0f;
		float phi = (atan2f(ud.z, ud.x) / 4.0f) + eval_value(s, n->in[2], rec);
		float theta = acosf((-ud.y / rr));
		float u = theta / PI;
		float v = (phi / (PI / 2.0f));
		u = wrapMinMax(u, 0.0f, 1.0f);
		v = wrapMinMax(v, 0.0f, 1.0f);
		rec->uv = (v2){ v, u };
		float strength = eval_value(s, n->in[1], rec);
		return (struct bsdf_sample){ (v3){ 0, 0, 0 }, ccoef(strength, eval_color(s, n->in[0], rec)) };
	}
	default:
		return (struct bsdf_sample){ (v3){ 0, 0, 0 }, (col){ 0, 0, 0, 1 } };
	}
}

/* ---- pathtrace.c:26-60 ---------------------------------------------------------------------------------------------------- */
static struct hit closest_isect(const struct trav_ctx *c, const struct ray *ray) {
	struct hit isect;
	memset(&isect, 0, sizeof isect);
	isect.incident = *ray;
	isect.inst = -1;
	isect.dist = FLT_MAX;
	isect.poly = -1;
	if (c->ctr) c->ctr->rays++;
	const uint64_t before = c->ctr ? c->ctr->node_pairs : 0;
	traverse(c, c->s->top_bvh, -1, ray, &isect);
	if (c->ctr && c->ctr->node_pairs - before > c->ctr->max_pairs_ray) c->ctr->max_pairs_ray = c->ctr->node_pairs - before;
	return isect;
}

static col path_trace(const struct trav_ctx *c, struct ray ray, int maxDepth, struct rng *r) {
	const struct crs_scene *s = c->s;
	col weight = { 1, 1, 1, 1 };
	col final = { 0, 0, 0, 1 };
	for (int depth = 0; depth < maxDepth; ++depth) {
		size_t hdr = 0;
		const size_t HDR = 36;
		int have_hdr = 0;
		if (c->tr && c->tr->len + HDR <= c->tr->cap) {           /* record header: x, y, pass, depth (u16), event count (u32), o, d (6 x f32) */
			hdr = c->tr->len;
			have_hdr = 1;
			const uint16_t h[4] = { c->tr->x, c->tr->y, c->tr->pass, (uint16_t)depth };
			const float od[6] = { ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z };
			memcpy(c->tr->buf + hdr, h, 8);
			memcpy(c->tr->buf + hdr + 12, od, 24);
			c->tr->len += HDR;
		}
		struct hit isect = closest_isect(c, &ray);
		if (have_hdr) { const uint32_t n = (uint32_t)(c->tr->len - hdr - HDR); memcpy(c->tr->buf + hdr + 8, &n, 4); }
		if (c->ctr && (uint64_t)(depth + 1) > c->ctr->max_depth) c->ctr->max_depth = (uint64_t)(depth + 1);
		if (isect.inst < 0) {
			final = cadd(final, cmul(weight, sample_bsdf(s, s->background, r, &isect).color));
			break;
		}
		const struct crs_material *m = &s->materials[isect.material];
		final = cadd(final, cmul(weight, (col){ m->emission[0], m->emission[1], m->emission[2], m->emission[3] }));
		const struct bsdf_sample smp = sample_bsdf(s, m->bsdf, r, &isect);
		ray = (struct ray){ isect.p, smp.out };
		const col att = smp.color;
		float probability = 1.0f;
		if (depth >= 4) {
			probability = MAX_(att.r, MAX_(att.g, att.b));
			if (draw(r) > probability) break;
		}
		weight = ccoef(1.0f / probability, cmul(att, weight));
	}
	return final;
}

/* ---- renderer.c:271-320 ------------------------------------------------------------------------------------------------------ */
struct job {
	const struct crs_scene *s;
	int x0, x1, ybegin, yend, pass_begin, pass_count;
	float *rgb;
	struct cro_counters ctr;
	bool count;
};

static void *render_rows(void *arg) {
	struct job *j = arg;
	const struct crs_scene *s = j->s;
	struct trav_ctx c = { s, j->count ? &j->ctr : NULL, NULL };
	const int W = (int)s->prefs.image_width, H = (int)s->prefs.image_height;
	const int maxPasses = (int)s->prefs.sample_count, bounces = (int)s->prefs.bounces;
	for (int y = j->yend - 1; y > j->ybegin - 1; --y) {
		for (int x = j->x0; x < j->x1; ++x) {
			uint32_t pixIdx = (uint32_t)(y * W + x);
			float *px = j->rgb + ((size_t)x + (size_t)(H - (y + 1)) * (size_t)W) * 3;
			for (int pass = j->pass_begin; pass < j->pass_begin + j->pass_count; ++pass) {
				struct rng r = { 0, c.ctr };
				rng_init(&r, pixIdx, pass, maxPasses);
				col out = { px[0], px[1], px[2], 1.0f };
				struct ray ray = camera_ray(&s->camera, x, y, &r);
				if (c.ctr) c.ctr->paths++;
				col smp = path_trace(&c, ray, bounces, &r);
				const int completed = pass + 1;
				out = ccoef((float)(completed - 1), out);
				out = cadd(out, smp);
				float t = 1.0f / completed;
				out = ccoef(t, out);
				px[0] = out.r; px[1] = out.g; px[2] = out.b;
			}
		}
	}
	return NULL;
}

struct runner { struct job *jobs; int first, step, n; };

static void *run_chunks(void *arg) {
	struct runner *r = arg;
	for (int i = r->first; i < r->n; i += r->step) render_rows(&r->jobs[i]);
	return NULL;
}

/* Traversal event trace of every ray of a region (single thread; pixels row-major y up, passes innermost), see struct cro_trace.
 * Returns the number of bytes written (records are dropped, not truncated, once the buffer is full). */
size_t cro_trace_region(const struct crs_scene *s, int x0, int y0, int x1, int y1, int pass_begin, int pass_count, uint8_t *buf, size_t cap) {
	struct cro_trace tr = { buf, cap, 0, 0, 0, 0, 0 };
	struct trav_ctx c = { s, NULL, &tr };
	const int W = (int)s->prefs.image_width;
	const int maxPasses = (int)s->prefs.sample_count, bounces = (int)s->prefs.bounces;
	for (int y = y0; y < y1; ++y)
		for (int x = x0; x < x1; ++x)
			for (int pass = pass_begin; pass < pass_begin + pass_count; ++pass) {
				struct rng r = { 0, NULL };
				rng_init(&r, (uint32_t)(y * W + x), pass, maxPasses);
				tr.x = (uint16_t)x; tr.y = (uint16_t)y; tr.pass = (uint16_t)pass;
				struct ray ray = camera_ray(&s->camera, x, y, &r);
				if (tr.len + 4096 > tr.cap) return tr.len;
				path_trace(&c, ray, bounces, &r);
			}
	return tr.len;
}

int cro_render(const struct crs_scene *s, int x0, int y0, int x1, int y1, int pass_begin, int pass_count,
			   float *rgb, int threads, struct cro_counters *counters) {
	if (threads < 1) threads = 1;
	int rows = y1 - y0;
	if (rows < 1 || x1 <= x0) return 0;
	if (threads > rows) threads = rows;
	/* interleaved row chunks so that threads finish together */
	int chunks = threads * 8;
	if (chunks > rows) chunks = rows;
	struct job *jobs = calloc((size_t)chunks, sizeof *jobs);
	for (int i = 0; i < chunks; ++i) {
		jobs[i] = (struct job){ .s = s, .x0 = x0, .x1 = x1, .ybegin = y0 + (int)((int64_t)rows * i / chunks),
			.yend = y0 + (int)((int64_t)rows * (i + 1) / chunks), .pass_begin = pass_begin, .pass_count = pass_count,
			.rgb = rgb, .count = counters != NULL };
	}
	/* simple static round-robin over threads: each thread runs chunks t, t+threads, ... */
	pthread_t *tids = calloc((size_t)threads, sizeof *tids);
	struct runner *rs = calloc((size_t)threads, sizeof *rs);
	for (int t = 0; t < threads; ++t) {
		rs[t] = (struct runner){ jobs, t, threads, chunks };
		pthread_create(&tids[t], NULL, run_chunks, &rs[t]);
	}
	for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
	if (counters) {
		memset(counters, 0, sizeof *counters);
		for (int i = 0; i < chunks; ++i) {
			const struct cro_counters *k = &jobs[i].ctr;
			counters->paths += k->paths; counters->rays += k->rays; counters->node_pairs += k->node_pairs;
			counters->tri_tests += k->tri_tests; counters->sphere_tests += k->sphere_tests;
			counters->inst_visits += k->inst_visits; counters->draws += k->draws;
			if (k->max_depth > counters->max_depth) counters->max_depth = k->max_depth;
			if (k->max_stack > counters->max_stack) counters->max_stack = k->max_stack;
			if (k->max_pairs_ray > counters->max_pairs_ray) counters->max_pairs_ray = k->max_pairs_ray;
		}
	}
	free(rs); free(tids); free(jobs);
	return 0;
}

/* ---- known-answer helpers ------------------------------------------------------------------------------------------------------ */
void cro_sampler_kat(uint32_t pixIdx, int pass, int maxPasses, int n, float *out) {
	struct rng r = { 0, NULL };
	rng_init(&r, pixIdx, pass, maxPasses);
	for (int i = 0; i < n; ++i) out[i] = draw(&r);
}

void cro_trace_kat(const struct crs_scene *s, int x, int y, int pass, struct cro_hit_kat *k) {
	memset(k, 0, sizeof *k);
	const int W = (int)s->prefs.image_width;
	struct trav_ctx c = { s, NULL, NULL };
	k->x = x; k->y = y; k->pixIdx = y * W + x;
	struct rng r = { 0, NULL };
	rng_init(&r, (uint32_t)k->pixIdx, pass, (int)s->prefs.sample_count);
	struct ray ray = camera_ray(&s->camera, x, y, &r);
	k->o[0] = ray.o.x; k->o[1] = ray.o.y; k->o[2] = ray.o.z;
	k->d[0] = ray.d.x; k->d[1] = ray.d.y; k->d[2] = ray.d.z;
	struct hit isect = closest_isect(&c, &ray);
	k->instIndex = isect.inst;
	struct bsdf_sample smp;
	if (isect.inst < 0) {
		k->polyIndex = -1;
		smp = sample_bsdf(s, s->background, &r, &isect);
	} else {
		k->polyIndex = isect.poly;
		k->distance = isect.dist; k->uv[0] = isect.uv.x; k->uv[1] = isect.uv.y;
		k->hitPoint[0] = isect.p.x; k->hitPoint[1] = isect.p.y; k->hitPoint[2] = isect.p.z;
		k->normal[0] = isect.n.x; k->normal[1] = isect.n.y; k->normal[2] = isect.n.z;
		const struct crs_material *m = &s->materials[isect.material];
		k->emission[0] = m->emission[0]; k->emission[1] = m->emission[1]; k->emission[2] = m->emission[2];
		smp = sample_bsdf(s, m->bsdf, &r, &isect);
	}
	k->out[0] = smp.out.x; k->out[1] = smp.out.y; k->out[2] = smp.out.z;
	k->color[0] = smp.color.r; k->color[1] = smp.color.g; k->color[2] = smp.color.b; k->color[3] = smp.color.a;
	k->nextDraw = draw(&r);
}

void cro_to_srgb8(const float *rgb, uint8_t *out, size_t pixels) {
	for (size_t i = 0; i < pixels * 3; ++i) {
		float c = linear_to_srgb(rgb[i]);
		out[i] = (unsigned char)MIN_(c * 255.0f, 255.0f);
	}
}
