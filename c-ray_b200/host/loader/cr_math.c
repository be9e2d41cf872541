/*
 * cr_math.c — 4x4 transforms for the scene loader (see cr_math.h).
 * Behaviour follows reference src/datatypes/transforms.c: rotation/translate/scale constructors (:107-200),
 * adjugate-over-determinant inverse (:202-283), row-by-column product (:293-313), point / direction transforms
 * (:74-104) and the abs-matrix bounding-box transform (:86-94).  The inverse is expanded along the first row with
 * the same association of products as the reference's recursive cofactor expansion, so results match bit for bit.
 */
#include "cr_math.h"
#include <string.h>

mat4 crl_identity(void) {
	mat4 r;
	memset(&r, 0, sizeof(r));
	r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f;
	return r;
}

mat4 crl_mul(const mat4 *A, const mat4 *B) {
	mat4 r;
	for (int i = 0; i < 4; ++i)
		for (int j = 0; j < 4; ++j)
			r.m[i][j] = A->m[i][0] * B->m[0][j] + A->m[i][1] * B->m[1][j] + A->m[i][2] * B->m[2][j] + A->m[i][3] * B->m[3][j];
	return r;
}

/* Determinant of the n x n minor of m picked by rows[] / cols[], Laplace expansion along its first row:
 * det = 0 + (s*a0)*D0 + (s*a1)*D1 ..., sign alternating from +1 (transforms.c:240-258). */
static float minor_det(const mat4 *m, const int *rows, const int *cols, int n) {
	if (n == 1) return m->m[rows[0]][cols[0]];
	float det = 0.0f, sign = 1.0f;
	int sub[3];
	for (int f = 0; f < n; ++f) {
		for (int c = 0, k = 0; c < n; ++c) if (c != f) sub[k++] = cols[c];
		det += sign * m->m[rows[0]][cols[f]] * minor_det(m, rows + 1, sub, n - 1);
		sign = -sign;
	}
	return det;
}

/* closed-form 4x4 determinant in the association of transforms.c:232-238 */
static float det4(const mat4 *M) {
	const float (*A)[4] = M->m;
	float s2233 = (A[2][2] * A[3][3]) - (A[2][3] * A[3][2]);
	float s2133 = (A[2][1] * A[3][3]) - (A[2][3] * A[3][1]);
	float s2132 = (A[2][1] * A[3][2]) - (A[2][2] * A[3][1]);
	float s2033 = (A[2][0] * A[3][3]) - (A[2][3] * A[3][0]);
	float s2032 = (A[2][0] * A[3][2]) - (A[2][2] * A[3][0]);
	float s2031 = (A[2][0] * A[3][1]) - (A[2][1] * A[3][0]);
	float c0 = A[0][0] * ((A[1][1] * s2233) - (A[1][2] * s2133) + (A[1][3] * s2132));
	float c1 = A[0][1] * ((A[1][0] * s2233) - (A[1][2] * s2033) + (A[1][3] * s2032));
	float c2 = A[0][2] * ((A[1][0] * s2133) - (A[1][1] * s2033) + (A[1][3] * s2031));
	float c3 = A[0][3] * ((A[1][0] * s2132) - (A[1][1] * s2032) + (A[1][2] * s2031));
	return c0 - c1 + c2 - c3;
}

mat4 crl_inverse(const mat4 *m) {
	float det = det4(m);
	mat4 inv;
	for (int i = 0; i < 4; ++i) {
		int rows[3];
		for (int r = 0, k = 0; r < 4; ++r) if (r != i) rows[k++] = r;
		for (int j = 0; j < 4; ++j) {
			int cols[3];
			for (int c = 0, k = 0; c < 4; ++c) if (c != j) cols[k++] = c;
			int sign = ((i + j) % 2 == 0) ? 1 : -1;
			float adj = (sign) * (minor_det(m, rows, cols, 3));
			inv.m[j][i] = adj / det;      /* the reference divides, then transposes (transforms.c:275-283) */
		}
	}
	return inv;
}

struct xform crl_xf_identity(void) {
	struct xform t;
	t.type = XF_IDENTITY;
	t.A = crl_identity();
	t.Ainv = t.A;
	return t;
}

static struct xform finish(struct xform t, enum xf_type type) {
	t.type = type;
	t.Ainv = crl_inverse(&t.A);
	return t;
}

struct xform crl_xf_rotate_x(float rads) {
	struct xform t = crl_xf_identity();
	float c = cosf(rads), s = sinf(rads);
	t.A.m[1][1] = c; t.A.m[1][2] = -s;
	t.A.m[2][1] = s; t.A.m[2][2] = c;
	return finish(t, XF_ROTATE_X);
}

struct xform crl_xf_rotate_y(float rads) {
	struct xform t = crl_xf_identity();
	float c = cosf(rads), s = sinf(rads);
	t.A.m[0][0] = c;  t.A.m[0][2] = s;
	t.A.m[2][0] = -s; t.A.m[2][2] = c;
	return finish(t, XF_ROTATE_Y);
}

struct xform crl_xf_rotate_z(float rads) {
	struct xform t = crl_xf_identity();
	float c = cosf(rads), s = sinf(rads);
	t.A.m[0][0] = c; t.A.m[0][1] = -s;
	t.A.m[1][0] = s; t.A.m[1][1] = c;
	return finish(t, XF_ROTATE_Z);
}

struct xform crl_xf_translate(float x, float y, float z) {
	struct xform t = crl_xf_identity();
	t.A.m[0][3] = x; t.A.m[1][3] = y; t.A.m[2][3] = z;
	return finish(t, XF_TRANSLATE);
}

struct xform crl_xf_scale(float x, float y, float z) {
	struct xform t = crl_xf_identity();
	t.A.m[0][0] = x; t.A.m[1][1] = y; t.A.m[2][2] = z;
	return finish(t, XF_SCALE);
}

vec3 crl_point(vec3 v, const mat4 *m) {
	vec3 r;
	r.x = (m->m[0][0] * v.x) + (m->m[0][1] * v.y) + (m->m[0][2] * v.z) + m->m[0][3];
	r.y = (m->m[1][0] * v.x) + (m->m[1][1] * v.y) + (m->m[1][2] * v.z) + m->m[1][3];
	r.z = (m->m[2][0] * v.x) + (m->m[2][1] * v.y) + (m->m[2][2] * v.z) + m->m[2][3];
	return r;
}

vec3 crl_vector(vec3 v, const mat4 *m) {
	vec3 r;
	r.x = (m->m[0][0] * v.x) + (m->m[0][1] * v.y) + (m->m[0][2] * v.z);
	r.y = (m->m[1][0] * v.x) + (m->m[1][1] * v.y) + (m->m[1][2] * v.z);
	r.z = (m->m[2][0] * v.x) + (m->m[2][1] * v.y) + (m->m[2][2] * v.z);
	return r;
}

void crl_transform_bbox(bbox3 *b, const mat4 *m) {
	mat4 a = *m;
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) a.m[i][j] = fabsf(m->m[i][j]);
	vec3 center = v_scale(v_add(b->min, b->max), 0.5f);
	vec3 half = v_scale(v_sub(b->max, b->min), 0.5f);
	half = crl_vector(half, &a);
	center = crl_point(center, &a);
	b->min = v_sub(center, half);
	b->max = v_add(center, half);
}
