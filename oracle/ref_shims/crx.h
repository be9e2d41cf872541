/*
 * crx.h — introspection hooks into the UNMODIFIED reference sources (test infrastructure only).
 *
 * The reference keeps its BVH and every node struct file-private (`struct bvh` bvh.c:44-48,
 * `struct mixBsdf` mix.c:21-26, ...), each with a file-static sample()/eval().  To flatten a loaded
 * `struct world` into include/crscene.h without copying or editing any reference source, each shim
 * translation unit in this directory does `#include "<reference .c file>"` (compiled from where it
 * lies under /root/reference) and appends one accessor that can see the private struct.  The shim
 * object replaces the plain object of that file at link time; the reference code itself is unchanged.
 */
#pragma once
#include <stdbool.h>
#include <stdint.h>

struct bvh; struct bvhNode; struct bsdfNode; struct colorNode; struct valueNode; struct vectorNode; struct texture;

struct crx_nodeinfo {
	int kind;                 /* enum crs_node_kind */
	const void *in[3];
	float f[8];
	const struct texture *tex;
	unsigned options;
};

/* bvh.c */
unsigned crx_bvh_node_count(const struct bvh *b);
const void *crx_bvh_nodes(const struct bvh *b);          /* array of 32-byte struct bvhNode */
const int *crx_bvh_prim_indices(const struct bvh *b);

/* each returns true and fills *o when `n` is a node of that file's type */
bool crx_is_diffuse(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_metal(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_glass(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_plastic(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_mix(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_add(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_transparent(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_emission(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_background(const struct bsdfNode *n, struct crx_nodeinfo *o);
bool crx_is_isotropic(const struct bsdfNode *n, struct crx_nodeinfo *o);

bool crx_is_constant_color(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_image(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_checker(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_gradient(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_blackbody(const struct colorNode *n, struct crx_nodeinfo *o);

bool crx_is_constant_value(const struct valueNode *n, struct crx_nodeinfo *o);
bool crx_is_grayscale(const struct valueNode *n, struct crx_nodeinfo *o);
bool crx_is_alpha(const struct valueNode *n, struct crx_nodeinfo *o);

/* the node types without a JSON path */
bool crx_is_vectocolor(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_combine_value(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_combine_rgb(const struct colorNode *n, struct crx_nodeinfo *o);
bool crx_is_math(const struct valueNode *n, struct crx_nodeinfo *o);
bool crx_is_fresnel(const struct valueNode *n, struct crx_nodeinfo *o);
bool crx_is_raylength(const struct valueNode *n, struct crx_nodeinfo *o);
bool crx_is_constant_vector(const struct vectorNode *n, struct crx_nodeinfo *o);
bool crx_is_normal(const struct vectorNode *n, struct crx_nodeinfo *o);
bool crx_is_vecmath(const struct vectorNode *n, struct crx_nodeinfo *o);
