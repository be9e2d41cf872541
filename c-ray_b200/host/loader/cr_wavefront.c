/*
 * cr_wavefront.c — Wavefront OBJ + MTL reader of the scene loader.
 *
 * What it has to reproduce (reference src/utils/loaders/formats/wavefront/wavefront.c, mtlloader.c and the
 * line/token buffer of src/utils/textbuffer.c), quirks included, because polygon indices, vertex-buffer
 * offsets and material order all end up in the flat scene:
 *   - a "line" ends at '\n'; text after the last '\n' is ignored unless the file has no '\n' at all
 *     (textbuffer.c:48-57, nextLine :104-114); lines are cut at 2047 bytes (fillLineBuffer :162-174);
 *   - tokens are split on ONE delimiter character, empty tokens are kept (so "f 1//2" has an empty vt field);
 *   - the per-file vertex count is the number of lines that START with "v" (so vn/vt lines inflate it,
 *     wavefront.c:139), and that inflated number is both the base for negative indices and the amount the
 *     global vertex buffer grows by (:119-125, :249);
 *   - an index of 0 ("unused") becomes (global count - 1), i.e. -1 only for the first mesh (:104-125);
 *   - quads become (a,b,c),(a,c,d) (:78-101); a face needs v/vt/vn or v//vn syntax (anything else crashes
 *     the reference; reported as an error here);
 *   - MTL: Ka/Kd/Ks/Ke colours get alpha 1, everything else in a material starts at zero (calloc).
 */
#include "cr_loader_int.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <libgen.h>

#define LINE_MAX_BYTES 2048

/* ---- file + line iteration ------------------------------------------------------------------------ */
static char *read_file(const char *path, size_t *len) {
	FILE *f = fopen(path, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END);
	long n = ftell(f);
	fseek(f, 0, SEEK_SET);
	if (n <= 0) { fclose(f); return NULL; }            /* empty files count as missing (fileio.c:74-78) */
	char *buf = malloc((size_t)n + 1);
	if (!buf || fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); fclose(f); return NULL; }
	buf[n] = '\0';
	fclose(f);
	if (len) *len = (size_t)n;
	return buf;
}

struct lines { char *buf; size_t count, cur, off; };

static void lines_init(struct lines *l, char *text) {
	l->buf = text; l->count = 0; l->cur = 0; l->off = 0;
	size_t n = strlen(text);
	for (size_t i = 0; i < n; ++i) if (text[i] == '\n') { text[i] = '\0'; l->count++; }
}
static char *lines_first(struct lines *l) { l->cur = 0; l->off = 0; return l->buf; }
static char *lines_next(struct lines *l) {
	if (l->cur + 1 >= l->count) return NULL;
	l->off += strlen(l->buf + l->off) + 1;
	l->cur++;
	return l->buf + l->off;
}

/* ---- tokens ------------------------------------------------------------------------------------------ */
struct toks { char buf[LINE_MAX_BYTES]; size_t count, cur, off; };

static void toks_fill(struct toks *t, const char *s, char delim) {
	size_t n = strlen(s);
	if (n > LINE_MAX_BYTES - 1) n = LINE_MAX_BYTES - 1;
	memcpy(t->buf, s, n);
	t->buf[n] = '\0';
	t->count = 0;
	for (size_t i = 0; i < n + 1; ++i)
		if (t->buf[i] == delim || t->buf[i] == '\0') { t->buf[i] = '\0'; t->count++; }
}
static char *toks_first(struct toks *t) { t->cur = 0; t->off = 0; return t->buf; }
static char *toks_peek(struct toks *t) {
	if (t->cur + 1 >= t->count) return NULL;
	return t->buf + t->off + strlen(t->buf + t->off) + 1;
}
static char *toks_next(struct toks *t) {
	char *n = toks_peek(t);
	if (!n) return NULL;
	t->off = (size_t)(n - t->buf);
	t->cur++;
	return n;
}

static char *dir_of(const char *path) {                 /* fileio.c:180-194: dirname + "/" */
	char *copy = strdup(path);
	const char *d = dirname(copy);
	size_t n = strlen(d);
	char *out = malloc(n + 2);
	memcpy(out, d, n);
	out[n] = '/'; out[n + 1] = '\0';
	free(copy);
	return out;
}

static char *concat(const char *a, const char *b) {
	char *s = malloc(strlen(a) + strlen(b) + 1);
	strcpy(s, a);
	strcat(s, b);
	return s;
}

static int starts_with(const char *prefix, const char *s) { return strncmp(prefix, s, strlen(prefix)) == 0; }

/* ---- textures ---------------------------------------------------------------------------------------- */
int crl_load_texture(struct crl_ctx *c, const char *path_in) {
	char *path = strdup(path_in);
	path[strcspn(path, "\n")] = 0;                      /* textureloader.c:58 */
	struct cr_image img;
	int rc = cr_image_load(path, &img);
	if (rc) {
		fprintf(stderr, "cr_loader: cannot decode texture \"%s\" (%d)\n", path, rc);
		free(path);
		return -1;
	}
	free(path);
	c->textures = realloc(c->textures, (size_t)(c->texture_count + 1) * sizeof(*c->textures));
	c->textures[c->texture_count] = img;
	return c->texture_count++;
}

/* ---- MTL --------------------------------------------------------------------------------------------- */
static int next_float(struct toks *t, float *out) {
	const char *s = toks_next(t);
	if (!s) return -1;
	*out = (float)atof(s);
	return 0;
}

static int mtl_color(struct toks *t, struct crl_color *out) {
	if (next_float(t, &out->r) || next_float(t, &out->g) || next_float(t, &out->b)) return -1;
	out->a = 1.0f;
	return 0;
}

static int load_mtl(struct crl_ctx *c, const char *path, struct crl_material **out, int *out_count) {
	char *text = read_file(path, NULL);
	if (!text) return 1;                                 /* reference: warning, mesh keeps the warning material */
	struct lines file;
	lines_init(&file, text);
	char *dir = dir_of(path);

	int total = 0;
	for (char *h = lines_first(&file); h; h = lines_next(&file)) if (starts_with("newmtl", h)) total++;
	struct crl_material *mats = calloc((size_t)total + 1, sizeof(*mats));
	for (int i = 0; i < total; ++i) mats[i].texture = mats[i].specular_map = mats[i].bsdf = -1;
	struct crl_material *cur = NULL;
	int used = 0, rc = 0;
	struct toks *line = malloc(sizeof(*line));

	for (char *head = lines_first(&file); head && !rc; head = lines_next(&file)) {
		toks_fill(line, head, ' ');
		char *first = toks_first(line);
		if (first[0] == '#' || head[0] == '\0') continue;
		if (!strcmp(first, "newmtl")) {
			const char *name = toks_peek(line);
			if (!name || used >= total) { rc = -1; break; }
			cur = &mats[used++];
			cur->name = strdup(name);
			continue;
		}
		int known = !strcmp(first, "Ka") || !strcmp(first, "Kd") || !strcmp(first, "Ks") || !strcmp(first, "Ke") ||
		            !strcmp(first, "illum") || !strcmp(first, "Ns") || !strcmp(first, "d") || !strcmp(first, "r") ||
		            !strcmp(first, "sharpness") || !strcmp(first, "Ni") || !strcmp(first, "map_Kd") ||
		            !strcmp(first, "norm") || !strcmp(first, "map_Ns");
		if (!known) continue;
		if (!cur) { rc = -1; break; }                     /* the reference dereferences NULL here */
		struct crl_color tmp;
		float f;
		if (!strcmp(first, "Kd")) rc = mtl_color(line, &cur->diffuse);
		else if (!strcmp(first, "Ks")) rc = mtl_color(line, &cur->specular);
		else if (!strcmp(first, "Ke")) rc = mtl_color(line, &cur->emission);
		else if (!strcmp(first, "Ka")) rc = mtl_color(line, &tmp);
		else if (!strcmp(first, "illum")) { const char *s = toks_next(line); if (!s) rc = -1; else cur->illum = atoi(s); }
		else if (!strcmp(first, "Ni")) rc = next_float(line, &cur->IOR);
		else if (!strcmp(first, "map_Kd") || !strcmp(first, "map_Ns")) {
			const char *s = toks_next(line);
			if (!s) { rc = -1; break; }
			char *p = concat(dir, s);
			int handle = crl_load_texture(c, p);
			free(p);
			if (first[4] == 'K') cur->texture = handle; else cur->specular_map = handle;
		} else rc = next_float(line, &f);                 /* Ns, d, r, sharpness: parsed, never read downstream */
		/* "norm" (normal map) is decoded by the reference and never used by any node: not loaded */
	}
	free(line);
	free(dir);
	free(text);
	if (rc) {
		snprintf(c->err, sizeof(c->err), "malformed MTL file %s", path);
		free(mats);
		return -1;
	}
	*out = mats;
	*out_count = total;
	return 0;
}

/* ---- OBJ --------------------------------------------------------------------------------------------- */
static int fix_index(size_t max, int old) {              /* wavefront.c:104-113 */
	if (old == 0) return -1;
	if (old < 0) return (int)max + old;
	return old - 1;
}

static int find_material(const struct crl_material *set, int count, const char *name) {
	for (int i = 0; i < count; ++i) if (name && set[i].name && !strcmp(set[i].name, name)) return i;
	return 0;
}

int crl_load_obj(struct crl_ctx *c, const char *path, struct crl_mesh *out) {
	memset(out, 0, sizeof(*out));
	char *text = read_file(path, NULL);
	if (!text) return 1;
	struct lines file;
	lines_init(&file, text);
	char *dir = dir_of(path);
	struct toks *line = malloc(sizeof(*line)), *batch = malloc(sizeof(*batch));

	size_t fileVertices = 0, fileTexCoords = 0, fileNormals = 0, filePolys = 0;
	for (char *h = lines_first(&file); h; h = lines_next(&file)) {
		if (starts_with("v", h)) fileVertices++;
		if (starts_with("vt", h)) fileTexCoords++;
		if (starts_with("vn", h)) fileNormals++;
		if (h[0] == 'f') { toks_fill(line, h, ' '); filePolys += line->count > 4 ? 2 : 1; }
	}
	float *vertices = calloc(fileVertices * 3 + 1, sizeof(float));
	float *texcoords = calloc(fileTexCoords * 2 + 1, sizeof(float));
	float *normals = calloc(fileNormals * 3 + 1, sizeof(float));
	struct crs_poly *polys = calloc(filePolys + 1, sizeof(*polys));
	size_t nv = 0, nt = 0, nn = 0, np = 0;
	struct crl_material *materials = NULL;
	int material_count = 0, current_material = 0, rc = 0;
	const int vbase = c->vertex_count, tbase = c->texcoord_count, nbase = c->normal_count;

	for (char *head = lines_first(&file); head && !rc; head = lines_next(&file)) {
		toks_fill(line, head, ' ');
		char *first = toks_first(line);
		if (first[0] == '#' || first[0] == '\0' || first[0] == 'o' || first[0] == 'g') continue;
		if (!strcmp(first, "v") || !strcmp(first, "vn")) {
			float *dst = first[1] ? &normals[3 * nn] : &vertices[3 * nv];
			if ((first[1] ? nn >= fileNormals : nv >= fileVertices) ||
			    next_float(line, &dst[0]) || next_float(line, &dst[1]) || next_float(line, &dst[2])) { rc = -1; break; }
			if (first[1]) nn++; else nv++;
		} else if (!strcmp(first, "vt")) {
			float *dst = &texcoords[2 * nt];
			if (nt >= fileTexCoords || next_float(line, &dst[0]) || next_float(line, &dst[1])) { rc = -1; break; }
			nt++;
		} else if (!strcmp(first, "f")) {
			size_t tris = line->count - 3;               /* wavefront.c:80 */
			if (line->count < 4 || tris > 2 || np + tris > filePolys) { rc = -2; break; }
			for (size_t i = 0; i < tris && !rc; ++i) {
				struct crs_poly *p = &polys[np++];
				/* token numbers of the three corners: 1,2,3 then 1,3,4 */
				const size_t corner[2][3] = { { 1, 2, 3 }, { 1, 3, 4 } };
				for (int j = 0; j < 3; ++j) {
					toks_first(line);
					char *spec = NULL;
					for (size_t k = 0; k < corner[i][j]; ++k) spec = toks_next(line);
					if (!spec) { rc = -2; break; }
					toks_fill(batch, spec, '/');
					const char *sv = toks_first(batch), *st = toks_next(batch), *sn = toks_next(batch);
					if (!st || !sn) { rc = -2; break; }
					p->v[j] = vbase + fix_index(fileVertices, atoi(sv));
					p->t[j] = tbase + fix_index(fileTexCoords, atoi(st));
					p->n[j] = nbase + fix_index(fileNormals, atoi(sn));
				}
				p->material = (uint32_t)current_material;
				p->has_normals = p->n[0] != -1;
			}
		} else if (!strcmp(first, "usemtl")) {
			current_material = find_material(materials, material_count, toks_peek(line));
		} else if (!strcmp(first, "mtllib")) {
			const char *name = toks_peek(line);
			if (!name) { rc = -1; break; }
			char *p = concat(dir, name);
			struct crl_material *set = NULL;
			int n = 0, mrc = load_mtl(c, p, &set, &n);
			free(p);
			if (mrc < 0) { rc = -3; break; }
			if (mrc == 0) { materials = set; material_count = n; }
			else { materials = NULL; material_count = 0; }
		}
	}
	free(line); free(batch); free(dir); free(text);
	if (rc) {
		if (rc == -2) snprintf(c->err, sizeof(c->err), "%s: faces must be triangles or quads written v/vt/vn or v//vn", path);
		else if (rc == -1) snprintf(c->err, sizeof(c->err), "%s: malformed OBJ statement", path);
		free(vertices); free(texcoords); free(normals); free(polys);
		return -1;
	}

	if (!materials) {                                    /* wavefront.c:236-241: the pink warning material */
		materials = calloc(1, sizeof(*materials));
		materials[0].diffuse = (struct crl_color){ 1.0f, 0.0f, 0.5f, 1.0f };
		materials[0].type = CRL_LAMBERTIAN;
		materials[0].texture = materials[0].specular_map = materials[0].bsdf = -1;
		material_count = 1;
	}
	out->polys = polys;
	out->poly_count = (int)filePolys;
	out->materials = materials;
	out->material_count = material_count;
	out->texcoord_count = (int)nt;

	c->vertices = realloc(c->vertices, ((size_t)c->vertex_count + fileVertices + 1) * 3 * sizeof(float));
	memcpy(c->vertices + 3 * (size_t)c->vertex_count, vertices, fileVertices * 3 * sizeof(float));
	c->normals = realloc(c->normals, ((size_t)c->normal_count + fileNormals + 1) * 3 * sizeof(float));
	memcpy(c->normals + 3 * (size_t)c->normal_count, normals, fileNormals * 3 * sizeof(float));
	c->texcoords = realloc(c->texcoords, ((size_t)c->texcoord_count + fileTexCoords + 1) * 2 * sizeof(float));
	memcpy(c->texcoords + 2 * (size_t)c->texcoord_count, texcoords, fileTexCoords * 2 * sizeof(float));
	c->vertex_count += (int)fileVertices;
	c->normal_count += (int)fileNormals;
	c->texcoord_count += (int)fileTexCoords;
	free(vertices); free(normals); free(texcoords);
	return 0;
}
