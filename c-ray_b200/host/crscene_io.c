/*
 * crscene_io.c — save/load the flat scene description of include/crscene.h.
 *
 * File layout: [u32 magic][u32 version][u64 total bytes][struct crs_scene with NULL pointers]
 * followed by the 14 arrays in the order they are declared in struct crs_scene, each starting on a
 * 16-byte boundary.  No reference counterpart: c-ray never serialises its scene (the cluster mode
 * re-sends the JSON, reference src/utils/protocol/server.c:296-323).
 */
#include "../../include/crscene.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <pthread.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

struct section { const void *ptr; size_t bytes; };

static size_t align16(size_t x) { return (x + 15u) & ~(size_t)15u; }

static void sections_of(const struct crs_scene *s, struct section sec[14]) {
	sec[0]  = (struct section){ s->instances,    (size_t)s->instance_count   * sizeof(struct crs_instance) };
	sec[1]  = (struct section){ s->spheres,      (size_t)s->sphere_count     * sizeof(struct crs_sphere) };
	sec[2]  = (struct section){ s->meshes,       (size_t)s->mesh_count       * sizeof(struct crs_mesh) };
	sec[3]  = (struct section){ s->materials,    (size_t)s->material_count   * sizeof(struct crs_material) };
	sec[4]  = (struct section){ s->nodes,        (size_t)s->node_count       * sizeof(struct crs_node) };
	sec[5]  = (struct section){ s->textures,     (size_t)s->texture_count    * sizeof(struct crs_texture) };
	sec[6]  = (struct section){ s->bvhs,         (size_t)s->bvh_count        * sizeof(struct crs_bvh) };
	sec[7]  = (struct section){ s->bvh_nodes,    (size_t)s->bvh_node_count   * sizeof(struct crs_bvh_node) };
	sec[8]  = (struct section){ s->prim_indices, (size_t)s->prim_index_count * sizeof(int32_t) };
	sec[9]  = (struct section){ s->polys,        (size_t)s->poly_count       * sizeof(struct crs_poly) };
	sec[10] = (struct section){ s->vertices,     (size_t)s->vertex_count     * 3 * sizeof(float) };
	sec[11] = (struct section){ s->normals,      (size_t)s->normal_count     * 3 * sizeof(float) };
	sec[12] = (struct section){ s->texcoords,    (size_t)s->texcoord_count   * 2 * sizeof(float) };
	sec[13] = (struct section){ s->texdata,      (size_t)s->texdata_bytes };
}

static size_t header_bytes(void) { return align16(16 + sizeof(struct crs_scene)); }

int crscene_save(const struct crs_scene *s, const char *path) {
	struct section sec[14];
	sections_of(s, sec);
	size_t total = header_bytes();
	for (int i = 0; i < 14; ++i) total += align16(sec[i].bytes);

	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	uint32_t magic = CRS_MAGIC, version = CRS_VERSION;
	uint64_t total64 = total;
	struct crs_scene copy = *s;
	copy.instances = NULL; copy.spheres = NULL; copy.meshes = NULL; copy.materials = NULL;
	copy.nodes = NULL; copy.textures = NULL; copy.bvhs = NULL; copy.bvh_nodes = NULL;
	copy.prim_indices = NULL; copy.polys = NULL; copy.vertices = NULL; copy.normals = NULL;
	copy.texcoords = NULL; copy.texdata = NULL; copy.owner = NULL;
	static const char zeros[16] = {0};
	int ok = 1;
	ok &= fwrite(&magic, 4, 1, f) == 1;
	ok &= fwrite(&version, 4, 1, f) == 1;
	ok &= fwrite(&total64, 8, 1, f) == 1;
	ok &= fwrite(&copy, sizeof(copy), 1, f) == 1;
	size_t pad = header_bytes() - (16 + sizeof(copy));
	if (pad) ok &= fwrite(zeros, 1, pad, f) == pad;
	for (int i = 0; i < 14 && ok; ++i) {
		if (sec[i].bytes) ok &= fwrite(sec[i].ptr, 1, sec[i].bytes, f) == sec[i].bytes;
		pad = align16(sec[i].bytes) - sec[i].bytes;
		if (pad) ok &= fwrite(zeros, 1, pad, f) == pad;
	}
	if (fclose(f) != 0) ok = 0;
	return ok ? 0 : -2;
}

/* Files are mapped, not read: the arrays are only ever read (the uploader copies them to the GPU once), so a private
 * read-only mapping of the page cache saves the 50 MB copy a frame-per-scene host pays (~25 ms for hdr.json).
 * Live mappings are remembered so that crscene_free can tell them from malloc'ed scenes (the loader's, crloader.h). */
#define CRS_MAX_MAPPINGS 256
static struct { void *base; size_t size; } g_maps[CRS_MAX_MAPPINGS];
static pthread_mutex_t g_maps_lock = PTHREAD_MUTEX_INITIALIZER;

static int remember_mapping(void *base, size_t size) {
	int ok = 0;
	pthread_mutex_lock(&g_maps_lock);
	for (int i = 0; i < CRS_MAX_MAPPINGS && !ok; ++i)
		if (!g_maps[i].base) { g_maps[i].base = base; g_maps[i].size = size; ok = 1; }
	pthread_mutex_unlock(&g_maps_lock);
	return ok;
}

static size_t forget_mapping(void *base) {
	size_t size = 0;
	pthread_mutex_lock(&g_maps_lock);
	for (int i = 0; i < CRS_MAX_MAPPINGS && !size; ++i)
		if (g_maps[i].base == base) { size = g_maps[i].size; g_maps[i].base = NULL; g_maps[i].size = 0; }
	pthread_mutex_unlock(&g_maps_lock);
	return size;
}

int crscene_load(struct crs_scene *out, const char *path) {
	memset(out, 0, sizeof(*out));
	const int fd = open(path, O_RDONLY);
	if (fd < 0) return -1;
	struct stat st;
	if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); return -2; }
	const long size = (long)st.st_size;
	if (size < (long)header_bytes()) { close(fd); return -3; }
	uint8_t *buf = mmap(NULL, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
	int mapped = buf != MAP_FAILED && remember_mapping(buf, (size_t)size);
	if (!mapped) {                                            /* no mapping (or the table is full): read the file instead */
		if (buf != MAP_FAILED) munmap(buf, (size_t)size);
		buf = NULL;
		if (posix_memalign((void **)&buf, 64, (size_t)size) != 0) { close(fd); return -4; }
		size_t got = 0;
		while (got < (size_t)size) {
			const ssize_t n = read(fd, buf + got, (size_t)size - got);
			if (n <= 0) { close(fd); free(buf); return -2; }
			got += (size_t)n;
		}
	}
	close(fd);
#define CRS_FAIL(code) do { if (mapped) munmap(buf, forget_mapping(buf)); else free(buf); memset(out, 0, sizeof(*out)); return (code); } while (0)
	uint32_t magic, version;
	uint64_t total;
	memcpy(&magic, buf, 4); memcpy(&version, buf + 4, 4); memcpy(&total, buf + 8, 8);
	if (magic != CRS_MAGIC || version != CRS_VERSION || total != (uint64_t)size) CRS_FAIL(-5);
	memcpy(out, buf + 16, sizeof(*out));
	struct section sec[14];
	sections_of(out, sec); /* sizes only; pointers are NULL in the file */
	size_t off = header_bytes();
	void *ptrs[14];
	for (int i = 0; i < 14; ++i) {
		if (off > (size_t)size || sec[i].bytes > (size_t)size - off) CRS_FAIL(-6);      /* overflow-safe: counts come from the file */
		ptrs[i] = sec[i].bytes ? buf + off : NULL;
		off += align16(sec[i].bytes);
	}
#undef CRS_FAIL
	out->instances = ptrs[0]; out->spheres = ptrs[1]; out->meshes = ptrs[2]; out->materials = ptrs[3];
	out->nodes = ptrs[4]; out->textures = ptrs[5]; out->bvhs = ptrs[6]; out->bvh_nodes = ptrs[7];
	out->prim_indices = ptrs[8]; out->polys = ptrs[9]; out->vertices = ptrs[10]; out->normals = ptrs[11];
	out->texcoords = ptrs[12]; out->texdata = ptrs[13];
	out->owner = buf;
	return 0;
}

void crscene_free(struct crs_scene *s) {
	if (s && s->owner) {
		const size_t mapped = forget_mapping(s->owner);
		if (mapped) munmap(s->owner, mapped); else free(s->owner);
	}
	if (s) memset(s, 0, sizeof(*s));
}

/* Re-target a loaded scene to another image size / sample count / bounce limit, exactly as the
 * reference's CLI overrides do (-d WxH, -s N: src/utils/args.c:108-131 applied in
 * sceneloader.c:425-467) followed by newCamera (src/datatypes/camera.c:22-42): only the aspect-derived
 * sensor height changes; sensor width, aperture and the composite transform do not depend on W/H.
 * Values <= 0 keep the current setting. */
int crscene_set_config(struct crs_scene *s, int width, int height, int samples, int bounces) {
	if (!s) return -1;
	if (width > 0 && height > 0) {
		s->prefs.image_width = (uint32_t)width;
		s->prefs.image_height = (uint32_t)height;
		s->camera.width = width;
		s->camera.height = height;
		const float aspect = (float)s->camera.width / (float)s->camera.height; /* camera.c:29 */
		s->camera.sensor_y = s->camera.sensor_x / aspect;                       /* camera.c:31 */
	}
	if (samples > 0) s->prefs.sample_count = (uint32_t)samples;
	if (bounces > 0) s->prefs.bounces = (uint32_t)bounces;
	return 0;
}
