/*
 * crgpu_trace.cuh — camera ray generation and two-level BVH closest-hit query (device functions).
 *
 * Restates, over the device layout of crgpu_scene.cuh:
 *   getCameraRay            reference src/datatypes/camera.c:50-87
 *   traverseBvhGeneric      reference src/accelerators/bvh.c:354-441 (+ intersectNode :326-352)
 *   intersectTopLevelLeaf   bvh.c:468-486      intersectBottomLevelLeaf  bvh.c:443-462
 *   intersectSphere/Mesh    src/datatypes/instance.c:45-60, 169-185 (closest-hit part)
 *   rayIntersectsWithPolygon src/datatypes/poly.c:17-53      intersect  src/datatypes/sphere.c:20-50
 *
 * The traversal visits nodes and primitives in EXACTLY the reference order (nearer child first,
 * farther pushed; both children tested with the maxDist valid before either leaf is processed), so
 * the closest hit — including exact-tie winners — is the reference's.
 */
#pragma once
#include "crgpu_math.cuh"
#include "crgpu_scene.cuh"

struct Hit {
	float t, u, v;
	int   inst;          /* -1 = miss */
	uint32_t prim;       /* global prim slot of the winning triangle */
};

struct TraceCounters { unsigned pairs, tris, spheres, insts; };

/* ---- camera.c:50-87 ---------------------------------------------------------------------------------------- */
CRD float cr_triangle_distribution(float v) {
	const float orig = v * 2.0f - 1.0f;
	v = cr_div(orig, cr_sqrtf(fabsf(orig)));
	v = cr_clamp(v, -1.0f, 1.0f);
	v = v - ((orig >= 0.0f) ? 1.0f : -1.0f);
	return v;
}

CRD void cr_camera_ray(const DevCamera &cam, int x, int y, uint64_t &rng, v3 &o, v3 &d) {
	const v3 right = v3make(cam.right[0], cam.right[1], cam.right[2]);
	const v3 up = v3make(cam.up[0], cam.up[1], cam.up[2]);
	const v3 forward = v3make(cam.forward[0], cam.forward[1], cam.forward[2]);
	const float jitterX = cr_triangle_distribution(cr_draw(rng));
	const float jitterY = cr_triangle_distribution(cr_draw(rng));
	const v3 pixX = v3scale(right, cr_div(cam.sensor_x, (float)cam.width));
	const v3 pixY = v3scale(up, cr_div(cam.sensor_y, (float)cam.height));
	const v3 pixV = v3add(forward, v3add(v3scale(pixX, (float)x - (float)cam.width * 0.5f + jitterX + 0.5f),
										 v3scale(pixY, (float)y - (float)cam.height * 0.5f + jitterY + 0.5f)));
	o = v3make(0.0f, 0.0f, 0.0f);
	d = v3norm(pixV);
	if (cam.aperture > 0.0f) {
		const float ft = cr_div(cam.focal_distance, v3dot(d, forward));
		const v3 focus = v3add(o, v3scale(d, ft));
		const float rr = cr_sqrtf(cr_draw(rng));                               /* vector.h:194-198 */
		const float theta = (cr_draw(rng) * (2.0f * CR_PI - 0.0f)) + 0.0f;
		float sn, cs;
		cr_sincosf(theta, &sn, &cs);
		const float lx = (rr * cs) * cam.aperture, ly = (rr * sn) * cam.aperture;
		o = v3add(o, v3add(v3scale(right, lx), v3scale(up, ly)));
		d = v3norm(v3sub(focus, o));
	}
	o = xf_point(cam.A, o);
	d = xf_vector(cam.A, d);
}

/* ---- bvh.c:326-352 -------------------------------------------------------------------------------------------- */
struct RaySetup {
	v3 invDir, scaledStart;
	bool ox, oy, oz;
};

CRD RaySetup cr_ray_setup(v3 o, v3 d) {                                                /* bvh.c:369-376 */
	RaySetup s;
	s.ox = (__float_as_uint(d.x) >> 31) != 0u;
	s.oy = (__float_as_uint(d.y) >> 31) != 0u;
	s.oz = (__float_as_uint(d.z) >> 31) != 0u;
	s.invDir = v3make(cr_div(1.0f, d.x), cr_div(1.0f, d.y), cr_div(1.0f, d.z));
	s.scaledStart = v3scale(v3mul(o, s.invDir), -1.0f);
	return s;
}

CRD bool cr_node_test(const float *b, const RaySetup &r, float maxDist, float &tEntry) {
	/* strict reference: a*b+c with two roundings (FP_FAST_FMAF undefined without -march, bvh.c:318-324) */
	const float tMinX = (r.ox ? b[1] : b[0]) * r.invDir.x + r.scaledStart.x;
	const float tMaxX = (r.ox ? b[0] : b[1]) * r.invDir.x + r.scaledStart.x;
	const float tMinY = (r.oy ? b[3] : b[2]) * r.invDir.y + r.scaledStart.y;
	const float tMaxY = (r.oy ? b[2] : b[3]) * r.invDir.y + r.scaledStart.y;
	const float tMinZ = (r.oz ? b[5] : b[4]) * r.invDir.z + r.scaledStart.z;
	const float tMaxZ = (r.oz ? b[4] : b[5]) * r.invDir.z + r.scaledStart.z;
	float tMin = tMinX > tMinY ? tMinX : tMinY;
	float tMax = tMaxX < tMaxY ? tMaxX : tMaxY;
	tMin = tMin > tMinZ ? tMin : tMinZ;
	tMax = tMax < tMaxZ ? tMax : tMaxZ;
	tMin = tMin > 0 ? tMin : 0;
	tMax = tMax < maxDist ? tMax : maxDist;
	tEntry = tMin;
	return tMin <= tMax;
}

/* Generic ordered traversal; Leaf(first, count) -> bool tests a leaf and may shrink best.t. */
template <class Leaf, bool COUNT>
CRD bool cr_traverse(const DevBvh &bvh, const PairNode *__restrict__ pairs, v3 o, v3 d, Hit &best,
					 Leaf &leaf, uint32_t *stack, TraceCounters *ctr) {
	if (bvh.node_count < 1) {                                                      /* bvh.c:362-365 */
		best.inst = -1;
		return false;
	}
	const RaySetup rs = cr_ray_setup(o, d);
	float maxDist = best.t;
	if (bvh.node_count == 1) {                                                     /* bvh.c:382-387 */
		float tEntry;
		if (cr_node_test(bvh.root_bounds, rs, maxDist, tEntry)) return leaf(bvh.root_first, bvh.root_count);
		return false;
	}
	const PairNode *base = pairs + bvh.pair_offset;
	uint32_t node = 0;
	int sp = 0;
	bool hasHit = false;
	while (true) {
		const float4 *p4 = reinterpret_cast<const float4 *>(base + node);
		const float4 q0 = __ldg(p4 + 0), q1 = __ldg(p4 + 1), q2 = __ldg(p4 + 2);
		const uint4 q3 = __ldg(reinterpret_cast<const uint4 *>(p4 + 3));
		const float lb[6] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y };
		const float rb[6] = { q1.z, q1.w, q2.x, q2.y, q2.z, q2.w };
		if (COUNT) ctr->pairs++;
		float tL, tR;
		const bool hitL = cr_node_test(lb, rs, maxDist, tL);
		const bool hitR = cr_node_test(rb, rs, maxDist, tR);
		bool goL = false, goR = false;
		if (hitL) {
			if (q3.z & CRG_LEAF_BIT) {
				if (leaf(q3.x, q3.z & ~CRG_LEAF_BIT)) { maxDist = best.t; hasHit = true; }
			} else goL = true;
		}
		if (hitR) {
			if (q3.w & CRG_LEAF_BIT) {
				if (leaf(q3.y, q3.w & ~CRG_LEAF_BIT)) { maxDist = best.t; hasHit = true; }
			} else goR = true;
		}
		if (goL & goR) {
			const bool swap = tL > tR;                                             /* bvh.c:424-431 */
			node = swap ? q3.y : q3.x;
			stack[sp++] = swap ? q3.x : q3.y;
		} else if (goL ^ goR) {
			node = goL ? q3.x : q3.y;
		} else {
			if (sp == 0) break;
			node = stack[--sp];
		}
	}
	return hasHit;
}

/* ---- leaves ------------------------------------------------------------------------------------------------------ */
template <bool COUNT>
struct BottomLeaf {                                                               /* bvh.c:443-462 + poly.c:17-53 */
	const PackedTri *__restrict__ tris;   /* already offset to this BVH's first slot */
	uint32_t slot_base;
	v3 o, d;
	Hit *best;
	TraceCounters *ctr;
	CRD bool operator()(uint32_t first, uint32_t count) {
		bool found = false;
		for (uint32_t i = 0; i < count; ++i) {
			const float4 *t4 = reinterpret_cast<const float4 *>(tris + first + i);
			const float4 a = __ldg(t4 + 0), b = __ldg(t4 + 1), c4 = __ldg(t4 + 2);
			if (COUNT) ctr->tris++;
			const v3 v0 = v3make(a.x, a.y, a.z), e1 = v3make(a.w, b.x, b.y), e2 = v3make(b.z, b.w, c4.x);
			const v3 n = v3make(c4.y, c4.z, c4.w);
			const v3 c = v3sub(v0, o);
			const v3 r = v3cross(d, c);
			const float invDet = cr_div(1.0f, v3dot(n, d));
			const float u = v3dot(r, e2) * invDet;
			const float v = v3dot(r, e1) * invDet;
			if (u >= 0.0f && v >= 0.0f && u + v <= 1.0f) {
				const float t = v3dot(n, c) * invDet;
				if (t >= 0.0f && t < best->t) {
					best->t = t; best->u = u; best->v = v;
					best->prim = slot_base + first + i;
					found = true;
				}
			}
		}
		return found;
	}
};

CRD bool cr_sphere_test(v3 o, v3 d, float radius, float &dist) {                   /* sphere.c:20-50 */
	const float A = v3dot(d, d);
	const float B = 2.0f * v3dot(d, o);
	const float C = v3dot(o, o) - (radius * radius);
	const float disc = B * B - 4.0f * A * C;
	if (disc < 0.0f) return false;
	const float sq = cr_sqrtf(disc);
	float t0 = cr_div(-B + sq, 2.0f);
	const float t1 = cr_div(-B - sq, 2.0f);
	if (t0 > t1 && t1 > 0.0f) t0 = t1;
	if (t0 < 0.00001f || t0 > dist) return false;
	dist = t0;
	return true;
}

/* object-space ray of an instance: transformRay(Ainv) + rayOffset advance (instance.c:47-50, 170-174) */
CRD void cr_object_ray(const float *Ainv, float ray_offset, v3 o, v3 d, v3 &oo, v3 &od) {
	oo = xf_point(Ainv, o);
	od = xf_vector(Ainv, d);
	oo = v3add(oo, v3scale(od, ray_offset));
}

template <bool COUNT>
struct TopLeaf {                                                                   /* bvh.c:468-486 */
	const DevScene *sc;
	v3 o, d;
	Hit *best;
	uint32_t *stack2;
	TraceCounters *ctr;
	CRD bool operator()(uint32_t first, uint32_t count) {
		bool found = false;
		for (uint32_t i = 0; i < count; ++i) {
			const int cur = __ldg(sc->top_prims + sc->top.slot_offset + first + i);
			const DevInstance *inst = sc->instances + cur;
			const float4 *m4 = reinterpret_cast<const float4 *>(inst->Ainv);
			const float4 r0 = __ldg(m4 + 0), r1 = __ldg(m4 + 1), r2 = __ldg(m4 + 2);
			const float Ainv[12] = { r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w };
			const uint4 meta = __ldg(reinterpret_cast<const uint4 *>(&inst->kind)); /* kind, bvh, ray_offset, radius */
			v3 oo, od;
			cr_object_ray(Ainv, __uint_as_float(meta.z), o, d, oo, od);
			bool h;
			if (meta.x == CRS_INST_MESH) {                                         /* instance.c:169-175 */
				if (COUNT) ctr->insts++;
				const DevBvh bvh = sc->bvhs[meta.y];
				BottomLeaf<COUNT> leaf = { sc->tris + bvh.slot_offset, bvh.slot_offset, oo, od, best, ctr };
				h = cr_traverse<BottomLeaf<COUNT>, COUNT>(bvh, sc->pairs, oo, od, *best, leaf, stack2, ctr);
			} else {                                                               /* instance.c:45-51 */
				if (COUNT) ctr->spheres++;
				h = cr_sphere_test(oo, od, __uint_as_float(meta.w), best->t);
			}
			if (h) {
				best->inst = cur;
				found = true;
			}
		}
		return found;
	}
};

/* getClosestIsect, pathtrace.c:26-30 */
template <bool COUNT>
CRD Hit cr_closest_hit(const DevScene &sc, v3 o, v3 d, TraceCounters *ctr) {
	Hit best;
	best.t = CR_FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.inst = -1; best.prim = 0u;
	uint32_t stack1[CRG_MAX_STACK + 1];
	uint32_t stack2[CRG_MAX_STACK + 1];
	TopLeaf<COUNT> leaf = { &sc, o, d, &best, stack2, ctr };
	cr_traverse<TopLeaf<COUNT>, COUNT>(sc.top, sc.pairs, o, d, best, leaf, stack1, ctr);
	return best;
}
