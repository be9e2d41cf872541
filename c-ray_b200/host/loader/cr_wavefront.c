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
#include <pthread.h>

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
struct crl_tex_job { pthread_t thread; unsigned char *buf; size_t len; struct cr_image img; int rc; };

static void *tex_decode_thread(void *arg) {
	struct crl_tex_job *j = arg;
	j->rc = cr_image_decode(j->buf, j->len, &j->img);
	free(j->buf);
	j->buf = NULL;
	return NULL;
}

void crl_textures_join(struct crl_ctx *c) {
	for (int i = 0; i < c->texture_count; ++i) {
		struct crl_tex_job *j = c->tex_jobs ? c->tex_jobs[i] : NULL;
		if (!j) continue;
		pthread_join(j->thread, NULL);
		if (j->rc) { c->async_failed = 1; free(j->img.data); }
		else c->textures[i] = j->img;
		free(j);
		c->tex_jobs[i] = NULL;
	}
}

int crl_load_texture(struct crl_ctx *c, const char *path_in) {
	char *path = strdup(path_in);
	path[strcspn(path, "\n")] = 0;                      /* textureloader.c:58 */
	size_t len = 0;
	unsigned char *buf = (unsigned char *)read_file(path, &len);
	struct cr_image img;
	memset(&img, 0, sizeof img);
	struct crl_tex_job *job = NULL;
	int rc = buf ? 0 : -100;
	if (buf && c->async_textures && cr_image_probe(buf, len)) {
		job = calloc(1, sizeof *job);
		if (job) {
			job->buf = buf; job->len = len;
			if (pthread_create(&job->thread, NULL, tex_decode_thread, job) != 0) { free(job); job = NULL; }
		}
	}
	if (buf && !job) { rc = cr_image_decode(buf, len, &img); free(buf); }
	if (rc) {
		fprintf(stderr, "cr_loader: cannot decode texture \"%s\" (%d)\n", path, rc);
		free(path);
		return -1;
	}
	free(path);
	c->textures = realloc(c->textures, (size_t)(c->texture_count + 1) * sizeof(*c->textures));
	c->tex_jobs = realloc(c->tex_jobs, (size_t)(c->texture_count + 1) * sizeof(*c->tex_jobs));
	c->textures[c->texture_count] = img;
	c->tex_jobs[c->texture_count] = job;
	return c->texture_count++;
}

/* ---- MTL --------------------------------------------------------------------------------------------- */
static void free_materials(struct crl_material *set, int count) {
	if (!set) return;
	for (int i = 0; i < count; ++i) free(set[i].name);
	free(set);
}

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
		free_materials(mats, total);
		return -1;
	}
	*out = mats;
	*out_count = total;
	return 0;
}

/* ---- OBJ --------------------------------------------------------------------------------------------- */
/* The 23 MB Venus mesh of the headline scene is ~1.2 M lines; parsing is split over threads by cutting the text at
 * line boundaries.  Pass 1 (parallel) counts per chunk what the reference's count()/countPolygons() count and what its
 * parse loop will emit; a serial step turns that into output offsets and resolves mtllib/usemtl in file order; pass 2
 * (parallel) parses every chunk into its slice.  The result is independent of the number of chunks. */
static int fix_index(size_t max, int old) {              /* wavefront.c:104-113 */
	if (old == 0) return -1;
	if (old < 0) return (int)max + old;
	return old - 1;
}

static int find_material(const struct crl_material *set, int count, const char *name) {
	for (int i = 0; i < count; ++i) if (name && set[i].name && !strcmp(set[i].name, name)) return i;
	return 0;
}

/* in-place tokens of one line: `len` bytes at `buf`, delimiters already replaced by NULs */
struct ltoks { char *buf; size_t len, count, cur; char *pos; };
static char *lt_first(struct ltoks *t) { t->cur = 0; t->pos = t->buf; return t->buf; }
static char *lt_peek(struct ltoks *t) { return t->cur + 1 >= t->count ? NULL : t->pos + strlen(t->pos) + 1; }
static char *lt_next(struct ltoks *t) { char *n = lt_peek(t); if (n) { t->pos = n; t->cur++; } return n; }
static void lt_split(struct ltoks *t, char *line, size_t len, char delim) {
	t->buf = line; t->len = len; t->count = 1;
	for (size_t i = 0; i < len; ++i) if (line[i] == delim) { line[i] = '\0'; t->count++; }
	line[len] = '\0';
}
static int lt_float(struct ltoks *t, float *out) {
	const char *s = lt_next(t);
	if (!s) return -1;
	*out = (float)atof(s);
	return 0;
}

static inline size_t line_len(const char *p, const char *end) {
	const char *nl = memchr(p, '\n', (size_t)(end - p));
	return nl ? (size_t)(nl - p) : (size_t)(end - p);
}
static inline int first_is(const char *p, size_t len, const char *word, size_t wlen) {
	return len >= wlen && !memcmp(p, word, wlen) && (len == wlen || p[wlen] == ' ');
}

struct mtl_event { char *line; int is_lib; int index; };

struct obj_chunk {
	char *begin, *end;
	size_t starts_v, starts_vt, starts_vn, count_polys;   /* count() / countPolygons() */
	size_t nv, nt, nn, np;                                /* what the parse loop emits */
	struct mtl_event *events; int nevents, cap;
	size_t ov, ot, on, op;                                /* output offsets (prefix sums) */
	int start_material, rc;
};

struct obj_job {
	struct obj_chunk *chunks;
	float *vertices, *texcoords, *normals;
	struct crs_poly *polys;
	size_t fileVertices, fileTexCoords, fileNormals, filePolys;
	int vbase, tbase, nbase;
};

static void obj_pass1(void *arg, int k) {
	struct obj_chunk *c = &((struct obj_job *)arg)->chunks[k];
	for (char *p = c->begin; p < c->end;) {
		size_t full = line_len(p, c->end), len = full > LINE_MAX_BYTES - 1 ? LINE_MAX_BYTES - 1 : full;
		if (len && p[0] == 'v') {
			c->starts_v++;
			if (len >= 2 && p[1] == 't') c->starts_vt++;
			if (len >= 2 && p[1] == 'n') c->starts_vn++;
			if (first_is(p, len, "v", 1)) c->nv++;
			else if (first_is(p, len, "vt", 2)) c->nt++;
			else if (first_is(p, len, "vn", 2)) c->nn++;
		} else if (len && p[0] == 'f') {
			size_t tokens = 1;
			for (size_t i = 0; i < len; ++i) tokens += p[i] == ' ';
			c->count_polys += tokens > 4 ? 2 : 1;
			if (first_is(p, len, "f", 1)) c->np += tokens - 3;      /* validated in pass 2 */
		} else if (first_is(p, len, "usemtl", 6) || first_is(p, len, "mtllib", 6)) {
			if (c->nevents == c->cap) { c->cap = c->cap ? c->cap * 2 : 8; c->events = realloc(c->events, sizeof(*c->events) * (size_t)c->cap); }
			c->events[c->nevents++] = (struct mtl_event){ p, p[0] == 'm', 0 };
		}
		p += full + 1;
	}
}

static void obj_pass2(void *arg, int k) {
	struct obj_job *J = arg;
	struct obj_chunk *c = &J->chunks[k];
	size_t nv = c->ov, nt = c->ot, nn = c->on, np = c->op;
	int material = c->start_material, event = 0;
	struct ltoks line, spec;
	for (char *p = c->begin; p < c->end && !c->rc;) {
		size_t full = line_len(p, c->end), len = full > LINE_MAX_BYTES - 1 ? LINE_MAX_BYTES - 1 : full;
		char *next = p + full + 1;
		const char c0 = len ? p[0] : '\0';
		if (c0 == 'v' || c0 == 'f') {
			lt_split(&line, p, len, ' ');
			char *first = lt_first(&line);
			if (!strcmp(first, "v") || !strcmp(first, "vn")) {
				float *dst = first[1] ? &J->normals[3 * nn] : &J->vertices[3 * nv];
				if (lt_float(&line, &dst[0]) || lt_float(&line, &dst[1]) || lt_float(&line, &dst[2])) { c->rc = -1; break; }
				if (first[1]) nn++; else nv++;
			} else if (!strcmp(first, "vt")) {
				float *dst = &J->texcoords[2 * nt];
				if (lt_float(&line, &dst[0]) || lt_float(&line, &dst[1])) { c->rc = -1; break; }
				nt++;
			} else if (!strcmp(first, "f")) {
				const size_t tris = line.count - 3;            /* wavefront.c:80 */
				if (line.count < 4 || tris > 2 || np + tris > J->filePolys) { c->rc = -2; break; }
				int corner[4][3];
				char *corners[4];
				for (size_t k2 = 0; k2 < line.count - 1; ++k2) corners[k2] = lt_next(&line);
				for (size_t k2 = 0; k2 < line.count - 1; ++k2) {   /* each "v/vt/vn" corner, parsed once */
					char *sp = corners[k2];
					lt_split(&spec, sp, strlen(sp), '/');
					const char *sv = lt_first(&spec), *st = lt_next(&spec), *sn = lt_next(&spec);
					if (!st || !sn) { c->rc = -2; break; }
					corner[k2][0] = J->vbase + fix_index(J->fileVertices, atoi(sv));
					corner[k2][1] = J->tbase + fix_index(J->fileTexCoords, atoi(st));
					corner[k2][2] = J->nbase + fix_index(J->fileNormals, atoi(sn));
				}
				if (c->rc) break;
				static const int pick[2][3] = { { 0, 1, 2 }, { 0, 2, 3 } };   /* quads: (a,b,c),(a,c,d) wavefront.c:84-99 */
				for (size_t i = 0; i < tris; ++i) {
					struct crs_poly *poly = &J->polys[np++];
					for (int j = 0; j < 3; ++j) {
						poly->v[j] = corner[pick[i][j]][0];
						poly->t[j] = corner[pick[i][j]][1];
						poly->n[j] = corner[pick[i][j]][2];
					}
					poly->material = (uint32_t)material;
					poly->has_normals = poly->n[0] != -1;
				}
			}
		} else if (c0 == 'u' || c0 == 'm') {
			if (event < c->nevents && c->events[event].line == p) {
				if (!c->events[event].is_lib) material = c->events[event].index;
				event++;
			}
		}
		p = next;
	}
}

int crl_load_obj(struct crl_ctx *c, const char *path, struct crl_mesh *out) {
	memset(out, 0, sizeof(*out));
	char *text = read_file(path, NULL);
	if (!text) return 1;
	char *dir = dir_of(path);
	/* the lines the reference iterates: everything up to the last '\n' (or the whole text if there is none) */
	size_t len = strlen(text);
	char *end = text + len;
	{
		char *last = NULL;
		for (char *q = end; q > text; --q) if (q[-1] == '\n') { last = q; break; }
		if (last) end = last - 1;                             /* exclusive end of the last counted line */
		else end = text + len;
	}
	const int threads = crl_thread_count();
	int nchunks = (len > (1u << 20) && threads > 1) ? threads * 2 : 1;
	struct obj_job J;
	memset(&J, 0, sizeof(J));
	J.chunks = calloc((size_t)nchunks, sizeof(*J.chunks));
	{
		char *p = text;
		for (int k = 0; k < nchunks; ++k) {
			char *stop = k == nchunks - 1 ? end : text + (size_t)(end - text) * (size_t)(k + 1) / (size_t)nchunks;
			if (stop < p) stop = p;
			if (k != nchunks - 1) {                              /* move to the next line start */
				char *nl = stop < end ? memchr(stop, '\n', (size_t)(end - stop)) : NULL;
				stop = nl ? nl + 1 : end;
			}
			J.chunks[k].begin = p;
			J.chunks[k].end = stop;
			p = stop;
		}
	}
	crl_parallel_for(nchunks, obj_pass1, &J);

	struct crl_material *materials = NULL;
	int material_count = 0, current_material = 0, rc = 0;
	size_t nv = 0, nt = 0, nn = 0, np = 0;
	for (int k = 0; k < nchunks && !rc; ++k) {
		struct obj_chunk *ch = &J.chunks[k];
		J.fileVertices += ch->starts_v; J.fileTexCoords += ch->starts_vt; J.fileNormals += ch->starts_vn; J.filePolys += ch->count_polys;
		ch->ov = nv; ch->ot = nt; ch->on = nn; ch->op = np;
		nv += ch->nv; nt += ch->nt; nn += ch->nn; np += ch->np;
		ch->start_material = current_material;
		for (int e = 0; e < ch->nevents && !rc; ++e) {
			struct mtl_event *ev = &ch->events[e];
			size_t l = line_len(ev->line, ch->end);
			if (l > LINE_MAX_BYTES - 1) l = LINE_MAX_BYTES - 1;
			char *copy = malloc(l + 1);
			memcpy(copy, ev->line, l); copy[l] = '\0';
			char *arg = strchr(copy, ' ');                      /* second token: up to the next space */
			if (arg) { arg++; char *sp = strchr(arg, ' '); if (sp) *sp = '\0'; }
			if (ev->is_lib) {
				if (!arg) rc = -1;
				else {
					char *p = concat(dir, arg);
					struct crl_material *set = NULL;
					int n = 0, mrc = load_mtl(c, p, &set, &n);
					free(p);
					if (mrc < 0) rc = -3;
					else {
						free_materials(materials, material_count);          /* a later mtllib replaces the set (wavefront.c:213) */
						materials = mrc == 0 ? set : NULL;
						material_count = mrc == 0 ? n : 0;
					}
				}
			} else {
				current_material = find_material(materials, material_count, arg);
				ev->index = current_material;
			}
			free(copy);
		}
	}
	if (!rc && (nv > J.fileVertices || nt > J.fileTexCoords || nn > J.fileNormals || np > J.filePolys)) rc = -2;
	if (!rc) {
		J.vertices = calloc(J.fileVertices * 3 + 1, sizeof(float));
		J.texcoords = calloc(J.fileTexCoords * 2 + 1, sizeof(float));
		J.normals = calloc(J.fileNormals * 3 + 1, sizeof(float));
		J.polys = calloc(J.filePolys + 1, sizeof(*J.polys));
		J.vbase = c->vertex_count; J.tbase = c->texcoord_count; J.nbase = c->normal_count;
		crl_parallel_for(nchunks, obj_pass2, &J);
		for (int k = 0; k < nchunks; ++k) if (J.chunks[k].rc && !rc) rc = J.chunks[k].rc;
	}
	for (int k = 0; k < nchunks; ++k) free(J.chunks[k].events);
	free(J.chunks); free(dir); free(text);
	if (rc) {
		if (rc == -2) snprintf(c->err, sizeof(c->err), "%s: faces must be triangles or quads written v/vt/vn or v//vn", path);
		else if (rc == -1) snprintf(c->err, sizeof(c->err), "%s: malformed OBJ statement", path);
		free(J.vertices); free(J.texcoords); free(J.normals); free(J.polys);
		free_materials(materials, material_count);
		return -1;
	}

	if (!materials) {                                    /* wavefront.c:236-241: the pink warning material */
		materials = calloc(1, sizeof(*materials));
		materials[0].diffuse = (struct crl_color){ 1.0f, 0.0f, 0.5f, 1.0f };
		materials[0].type = CRL_LAMBERTIAN;
		materials[0].texture = materials[0].specular_map = materials[0].bsdf = -1;
		material_count = 1;
	}
	out->polys = J.polys;
	out->poly_count = (int)J.filePolys;
	out->materials = materials;
	out->material_count = material_count;
	out->texcoord_count = (int)nt;

	c->vertices = realloc(c->vertices, ((size_t)c->vertex_count + J.fileVertices + 1) * 3 * sizeof(float));
	memcpy(c->vertices + 3 * (size_t)c->vertex_count, J.vertices, J.fileVertices * 3 * sizeof(float));
	c->normals = realloc(c->normals, ((size_t)c->normal_count + J.fileNormals + 1) * 3 * sizeof(float));
	memcpy(c->normals + 3 * (size_t)c->normal_count, J.normals, J.fileNormals * 3 * sizeof(float));
	c->texcoords = realloc(c->texcoords, ((size_t)c->texcoord_count + J.fileTexCoords + 1) * 2 * sizeof(float));
	memcpy(c->texcoords + 2 * (size_t)c->texcoord_count, J.texcoords, J.fileTexCoords * 2 * sizeof(float));
	c->vertex_count += (int)J.fileVertices;
	c->normal_count += (int)J.fileNormals;
	c->texcoord_count += (int)J.fileTexCoords;
	free(J.vertices); free(J.normals); free(J.texcoords);
	return 0;
}
