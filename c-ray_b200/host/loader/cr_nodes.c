/*
 * cr_nodes.c — the shading-node store of the scene loader: constructors with the reference's defaults and its
 * hash-consing rule, plus the legacy material -> node-graph mapping.
 *
 * The reference interns every node in a hash table (src/nodes/nodebase.h:21-33, src/utils/hashtable.c:116-126): a
 * new node is dropped in favour of an existing one when the two have the same type, the same 32-bit FNV hash over the
 * type's hashed fields AND the type's compare() says equal.  Which fields take part differs per type (e.g. glass
 * ignores its IOR input, background looks at the colour input only, an image node hashes its options but compares
 * the texture only); since sharing decides how many nodes the flat scene has and which constants survive, the same
 * rule is applied here.  Inputs are pointers in the reference, handles here: "same bytes" <=> "same handle".
 * (A 32-bit hash collision between different inputs would make the reference keep both nodes; equal inputs always
 * collide, which is the only case modelled.)
 */
#include "cr_loader_int.h"
#include <stdlib.h>
#include <string.h>

const struct crl_color crl_black = { 0.0f, 0.0f, 0.0f, 1.0f };      /* color.c:16-18 */
const struct crl_color crl_gray  = { 0.5f, 0.5f, 0.5f, 1.0f };
const struct crl_color crl_white = { 1.0f, 1.0f, 1.0f, 1.0f };

/* how many leading inputs / constant floats identify a node of each kind, and whether tex+options do */
struct key_rule { int kind, inputs, floats, image; };
static const struct key_rule rules[] = {
	{ CRS_COLOR_CONSTANT, 0, 4, 0 },   /* constant.c:21-33 */
	{ CRS_COLOR_IMAGE,    0, 0, 1 },   /* image.c:16-29 */
	{ CRS_COLOR_CHECKER,  3, 0, 0 },
	{ CRS_COLOR_GRADIENT, 0, 8, 0 },
	{ CRS_COLOR_BLACKBODY,1, 0, 0 },
	{ CRS_VALUE_CONSTANT, 0, 1, 0 },
	{ CRS_VALUE_GRAYSCALE,1, 0, 0 },
	{ CRS_VALUE_ALPHA,    1, 0, 0 },
	{ CRS_BSDF_DIFFUSE,   1, 0, 0 },
	{ CRS_BSDF_METAL,     2, 0, 0 },
	{ CRS_BSDF_GLASS,     2, 0, 0 },   /* IOR input not part of the identity: glass.c:23-38 */
	{ CRS_BSDF_PLASTIC,   2, 0, 0 },   /* diffuse input not part of the identity: plastic.c:24-39 */
	{ CRS_BSDF_MIX,       3, 0, 0 },
	{ CRS_BSDF_ADD,       2, 0, 0 },
	{ CRS_BSDF_TRANSPARENT,1,0, 0 },
	{ CRS_BSDF_EMISSIVE,  2, 0, 0 },
	{ CRS_BSDF_BACKGROUND,1, 0, 0 },   /* strength and offset ignored: background.c:21-36 */
	{ CRS_BSDF_ISOTROPIC, 1, 0, 0 },
};

static const struct key_rule *rule_for(int kind) {
	for (size_t i = 0; i < sizeof(rules) / sizeof(rules[0]); ++i) if (rules[i].kind == kind) return &rules[i];
	return NULL;
}

static int same_node(const struct key_rule *r, const struct crl_node *a, const struct crl_node *b) {
	if (a->kind != b->kind) return 0;
	for (int i = 0; i < r->inputs; ++i) if (a->in[i] != b->in[i]) return 0;
	if (r->floats) {
		if (memcmp(a->f, b->f, sizeof(float) * (size_t)r->floats)) return 0;      /* same hash <= same bytes */
		for (int i = 0; i < r->floats; ++i) if (!(a->f[i] == b->f[i])) return 0;  /* compare(): NaN never equal */
	}
	if (r->image && (a->tex != b->tex || a->options != b->options)) return 0;
	return 1;
}

static int intern(struct crl_ctx *c, int kind, int in0, int in1, int in2, const float *f, int nf, int tex, uint32_t options) {
	struct crl_node n;
	memset(&n, 0, sizeof(n));
	n.kind = kind;
	n.in[0] = in0; n.in[1] = in1; n.in[2] = in2;
	if (nf) memcpy(n.f, f, sizeof(float) * (size_t)nf);
	n.tex = tex;
	n.options = options;
	const struct key_rule *r = rule_for(kind);
	for (int i = 0; i < c->node_count; ++i) if (same_node(r, &c->nodes[i], &n)) return i;
	if (c->node_count == c->node_cap) {
		c->node_cap = c->node_cap ? c->node_cap * 2 : 64;
		c->nodes = realloc(c->nodes, sizeof(*c->nodes) * (size_t)c->node_cap);
	}
	c->nodes[c->node_count] = n;
	return c->node_count++;
}

/* ---- colour nodes ------------------------------------------------------------------------------------ */
int crl_const_color(struct crl_ctx *c, struct crl_color col) {
	const float f[4] = { col.r, col.g, col.b, col.a };
	return intern(c, CRS_COLOR_CONSTANT, -1, -1, -1, f, 4, -1, 0);
}
int crl_image(struct crl_ctx *c, int tex, uint32_t options) {
	if (tex < 0) return -1;
	return intern(c, CRS_COLOR_IMAGE, -1, -1, -1, NULL, 0, tex, options);
}
int crl_checker(struct crl_ctx *c, int A, int B, int scale) {              /* checker.c:60-71 */
	if (A < 0) A = crl_const_color(c, crl_black);
	if (B < 0) B = crl_const_color(c, crl_white);
	if (scale < 0) scale = crl_const_value(c, 5.0f);
	return intern(c, CRS_COLOR_CHECKER, A, B, scale, NULL, 0, -1, 0);
}
int crl_gradient(struct crl_ctx *c, struct crl_color d, struct crl_color u) {
	const float f[8] = { d.r, d.g, d.b, d.a, u.r, u.g, u.b, u.a };
	return intern(c, CRS_COLOR_GRADIENT, -1, -1, -1, f, 8, -1, 0);
}
int crl_blackbody(struct crl_ctx *c, int t) {
	if (t < 0) t = crl_const_value(c, 4000.0f);
	return intern(c, CRS_COLOR_BLACKBODY, t, -1, -1, NULL, 0, -1, 0);
}

/* ---- value nodes ------------------------------------------------------------------------------------- */
int crl_const_value(struct crl_ctx *c, float v) { return intern(c, CRS_VALUE_CONSTANT, -1, -1, -1, &v, 1, -1, 0); }
int crl_grayscale(struct crl_ctx *c, int color) {
	if (color < 0) color = crl_const_color(c, crl_black);
	return intern(c, CRS_VALUE_GRAYSCALE, color, -1, -1, NULL, 0, -1, 0);
}
int crl_alpha(struct crl_ctx *c, int color) {
	if (color < 0) color = crl_const_color(c, crl_white);
	return intern(c, CRS_VALUE_ALPHA, color, -1, -1, NULL, 0, -1, 0);
}

/* ---- bsdf nodes (defaults: the constructors in reference src/nodes/shaders/) ----------------------------- */
int crl_diffuse(struct crl_ctx *c, int color) {
	if (color < 0) color = crl_const_color(c, crl_black);
	return intern(c, CRS_BSDF_DIFFUSE, color, -1, -1, NULL, 0, -1, 0);
}
int crl_metal(struct crl_ctx *c, int color, int roughness) {
	if (color < 0) color = crl_const_color(c, crl_black);
	if (roughness < 0) roughness = crl_const_value(c, 0.0f);
	return intern(c, CRS_BSDF_METAL, color, roughness, -1, NULL, 0, -1, 0);
}
int crl_glass(struct crl_ctx *c, int color, int roughness, int ior) {
	if (color < 0) color = crl_const_color(c, crl_black);
	if (roughness < 0) roughness = crl_const_value(c, 0.0f);
	if (ior < 0) ior = crl_const_value(c, 1.45f);
	return intern(c, CRS_BSDF_GLASS, color, roughness, ior, NULL, 0, -1, 0);
}
int crl_plastic(struct crl_ctx *c, int color) {                            /* plastic.c:89-100 */
	int col = color < 0 ? crl_const_color(c, crl_black) : color;
	int rough = crl_const_color(c, crl_black);
	int diff = crl_diffuse(c, color);
	return intern(c, CRS_BSDF_PLASTIC, col, rough, diff, NULL, 0, -1, 0);
}
int crl_mix(struct crl_ctx *c, int A, int B, int factor) {                 /* mix.c:52-67 */
	if (A == B) return A;
	if (A < 0) A = crl_diffuse(c, crl_const_color(c, crl_black));
	if (B < 0) B = crl_diffuse(c, crl_const_color(c, crl_black));
	if (factor < 0) factor = crl_const_value(c, 0.5f);
	return intern(c, CRS_BSDF_MIX, A, B, factor, NULL, 0, -1, 0);
}
int crl_add(struct crl_ctx *c, int A, int B) {
	if (A == B) return A;
	if (A < 0) A = crl_diffuse(c, crl_const_color(c, crl_black));
	if (B < 0) B = crl_diffuse(c, crl_const_color(c, crl_black));
	return intern(c, CRS_BSDF_ADD, A, B, -1, NULL, 0, -1, 0);
}
int crl_transparent(struct crl_ctx *c, int color) {
	if (color < 0) color = crl_const_color(c, crl_white);
	return intern(c, CRS_BSDF_TRANSPARENT, color, -1, -1, NULL, 0, -1, 0);
}
int crl_emissive(struct crl_ctx *c, int color, int strength) {
	if (color < 0) color = crl_const_color(c, crl_black);
	if (strength < 0) strength = crl_const_value(c, 1.0f);
	return intern(c, CRS_BSDF_EMISSIVE, color, strength, -1, NULL, 0, -1, 0);
}
int crl_background(struct crl_ctx *c, int color, int strength, int offset) {
	if (color < 0) color = crl_const_color(c, crl_gray);
	if (strength < 0) strength = crl_const_value(c, 1.0f);
	if (offset < 0) offset = crl_const_value(c, 0.0f);
	return intern(c, CRS_BSDF_BACKGROUND, color, strength, offset, NULL, 0, -1, 0);
}

int crl_warning_bsdf(struct crl_ctx *c) {                                  /* bsdfnode.c:16-21 */
	int a = crl_diffuse(c, crl_const_color(c, (struct crl_color){ 1.0f, 0.0f, 0.5f, 1.0f }));
	int b = crl_diffuse(c, crl_const_color(c, (struct crl_color){ 0.2f, 0.2f, 0.2f, 1.0f }));
	int f = crl_grayscale(c, crl_checker(c, -1, -1, crl_const_value(c, 500.0f)));
	return crl_mix(c, a, b, f);
}

/* ---- legacy materials (reference src/datatypes/material.c:52-107) ----------------------------------------- */
static int append_alpha(struct crl_ctx *c, int base, int color) {
	return crl_mix(c, crl_transparent(c, crl_const_color(c, crl_white)), base, crl_alpha(c, color));
}

void crl_assign_bsdf(struct crl_ctx *c, struct crl_material *m) {
	int roughness = m->specular_map >= 0 ? crl_grayscale(c, crl_image(c, m->specular_map, CRS_IMG_NO_BILINEAR))
	                                     : crl_const_value(c, m->roughness);
	int color = m->texture >= 0 ? crl_image(c, m->texture, CRS_IMG_SRGB_TRANSFORM) : crl_const_color(c, m->diffuse);
	int spec = crl_const_color(c, m->specular);
	m->bsdf = -1;
	if (m->illum == 5) m->bsdf = append_alpha(c, crl_metal(c, color, roughness), color);
	else if (m->illum == 7) m->bsdf = append_alpha(c, crl_glass(c, spec, roughness, crl_const_value(c, m->IOR)), spec);
	if (m->bsdf >= 0) return;
	switch (m->type) {
		case CRL_LAMBERTIAN:
		case CRL_EMISSION: m->bsdf = append_alpha(c, crl_diffuse(c, color), color); break;
		case CRL_GLASS:    m->bsdf = append_alpha(c, crl_glass(c, color, roughness, crl_const_value(c, m->IOR)), color); break;
		case CRL_METAL:    m->bsdf = append_alpha(c, crl_metal(c, color, roughness), color); break;
		case CRL_PLASTIC:  m->bsdf = append_alpha(c, crl_plastic(c, color), color); break;
		default:           m->bsdf = crl_warning_bsdf(c); break;
	}
}
