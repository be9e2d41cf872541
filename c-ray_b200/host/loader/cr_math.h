/*
 * cr_math.h — host-side math of the scene loader: 4x4 transforms, bounding boxes, camera set-up.
 * Restates the parts of reference src/datatypes/transforms.c, bbox.h and camera.c that run ONCE while a
 * scene is built (never in the hot path).  Compiled with -ffp-contract=off and written with the reference's
 * operation order so that matrices, ray offsets and camera vectors come out bit-identical to the strict
 * reference build (checked against its exported scenes in tests/test_loader.py).
 */
#pragma once
#include <math.h>
#include <float.h>
#include <stdbool.h>

#define CRL_PI 3.141592653589793238462643383279502f   /* includes.h:13 */

typedef struct { float x, y, z; } vec3;
typedef struct { float m[4][4]; } mat4;
typedef struct { vec3 min, max; } bbox3;

enum xf_type { XF_ROTATE_X, XF_ROTATE_Y, XF_ROTATE_Z, XF_TRANSLATE, XF_SCALE, XF_IDENTITY, XF_COMPOSITE };
struct xform { enum xf_type type; mat4 A, Ainv; };

static inline float crl_to_radians(float degrees) { return (degrees * CRL_PI) / 180.0f; }   /* transforms.c:17 */

static inline vec3 v_add(vec3 a, vec3 b) { return (vec3){ a.x + b.x, a.y + b.y, a.z + b.z }; }
static inline vec3 v_sub(vec3 a, vec3 b) { return (vec3){ a.x - b.x, a.y - b.y, a.z - b.z }; }
static inline vec3 v_scale(vec3 v, float c) { return (vec3){ v.x * c, v.y * c, v.z * c }; }
static inline vec3 v_cross(vec3 a, vec3 b) { return (vec3){ (a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x) }; }
static inline float v_len(vec3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
static inline vec3 v_norm(vec3 v) { float l = v_len(v); return (vec3){ v.x / l, v.y / l, v.z / l }; }
#define CRL_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define CRL_MAX(a, b) (((a) > (b)) ? (a) : (b))
static inline vec3 v_min(vec3 a, vec3 b) { return (vec3){ CRL_MIN(a.x, b.x), CRL_MIN(a.y, b.y), CRL_MIN(a.z, b.z) }; }
static inline vec3 v_max(vec3 a, vec3 b) { return (vec3){ CRL_MAX(a.x, b.x), CRL_MAX(a.y, b.y), CRL_MAX(a.z, b.z) }; }

static const bbox3 crl_empty_bbox = { { FLT_MAX, FLT_MAX, FLT_MAX }, { -FLT_MAX, -FLT_MAX, -FLT_MAX } };
static inline float bbox_half_area(const bbox3 *b) { vec3 e = v_sub(b->max, b->min); return e.x * (e.y + e.z) + e.y * e.z; }   /* bbox.h:26 */
static inline void bbox_extend(bbox3 *d, const bbox3 *s) { d->min = v_min(d->min, s->min); d->max = v_max(d->max, s->max); }
static inline vec3 bbox_center(const bbox3 *b) { return v_scale(v_add(b->max, b->min), 0.5f); }
static inline float bbox_ray_offset(bbox3 b) { return 0.0001f * v_len(v_sub(b.max, b.min)); }                                  /* bbox.h:44-46 */

mat4 crl_identity(void);
mat4 crl_mul(const mat4 *A, const mat4 *B);
mat4 crl_inverse(const mat4 *m);
struct xform crl_xf_identity(void);
struct xform crl_xf_rotate_x(float rads);
struct xform crl_xf_rotate_y(float rads);
struct xform crl_xf_rotate_z(float rads);
struct xform crl_xf_translate(float x, float y, float z);
struct xform crl_xf_scale(float x, float y, float z);
vec3 crl_point(vec3 v, const mat4 *m);
vec3 crl_vector(vec3 v, const mat4 *m);
void crl_transform_bbox(bbox3 *b, const mat4 *m);
