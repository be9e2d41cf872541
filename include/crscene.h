/*
 * crscene.h — the flat, pointer-free scene description that crosses the drop-in boundary.
 *
 * c-ray builds its scene into `struct world` (reference src/datatypes/scene.h:14-39): pointer-linked
 * instances, meshes, spheres, a camera, a hash-consed node graph and private BVH structs.  The GPU hot
 * path cannot chase those pointers, so the host side flattens the world ONCE after the BVH build
 * (the hook point is right after reference src/datatypes/scene.c:184) into the arrays below.  Every
 * array mirrors a reference structure field-for-field (citations next to each struct) so that the
 * flattening is a pure re-indexing step with no arithmetic: all floats are copied bit-for-bit.
 *
 * The same description can be serialised to a `.crscene` file (crscene_save / crscene_load in
 * c-ray_b200/host/crscene_io.c); the file is just the header + the arrays, 16-byte aligned.
 *
 * Plain C99, no CUDA or torch types.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRS_MAGIC   0x4e435343u /* "CSCN" little endian */
#define CRS_VERSION 2u

/* ---- node graph (reference src/nodes/) ------------------------------------------------------ */
enum crs_node_kind {
	CRS_NODE_NONE = 0,
	/* bsdf nodes: reference src/nodes/shaders/<name>.c */
	CRS_BSDF_DIFFUSE     = 1,  /* in0=color                                   diffuse.c:40-47   */
	CRS_BSDF_METAL       = 2,  /* in0=color in1=roughness(value)              metal.c:40-55     */
	CRS_BSDF_GLASS       = 3,  /* in0=color in1=roughness(value) in2=IOR(val) glass.c:41-87     */
	CRS_BSDF_PLASTIC     = 4,  /* in0=color in1=roughness(color) in2=diffuse  plastic.c:42-87   */
	CRS_BSDF_MIX         = 5,  /* in0=A in1=B in2=factor(value)               mix.c:42-50       */
	CRS_BSDF_ADD         = 6,  /* in0=A in1=B                                 add.c:42-49       */
	CRS_BSDF_TRANSPARENT = 7,  /* in0=color                                   transparent.c:40-44 */
	CRS_BSDF_EMISSIVE    = 8,  /* in0=color in1=strength(value)               emission.c:42-49  */
	CRS_BSDF_BACKGROUND  = 9,  /* in0=color in1=strength(value) in2=offset(v) background.c:39-66 */
	CRS_BSDF_ISOTROPIC   = 10, /* in0=color (volume scattering, no JSON path) isotropic.c:40-47 */
	/* color nodes: reference src/nodes/textures/<name>.c, converter/blackbody.c */
	CRS_COLOR_CONSTANT   = 32, /* f[0..3]=rgba                                constant.c:39-42  */
	CRS_COLOR_IMAGE      = 33, /* tex, options                                image.c:31-48     */
	CRS_COLOR_CHECKER    = 34, /* in0=A in1=B in2=scale(value)                checker.c:31-54   */
	CRS_COLOR_GRADIENT   = 35, /* f[0..3]=down f[4..7]=up                     gradient.c:40-45  */
	CRS_COLOR_BLACKBODY  = 36, /* in0=temperature(value)                      blackbody.c:38-42 */
	CRS_COLOR_VECTOCOLOR = 37, /* in0=vector -> (x, y, z, 0)                  converter/vectocolor.c:38-43 */
	CRS_COLOR_COMBINE_VALUE = 38, /* in0=value -> (v, v, v, 1)                converter/combine.c:38-43 */
	CRS_COLOR_COMBINE_RGB = 39, /* in0..2 = R, G, B values, alpha 1           converter/combinergb.c:44-53 */
	/* value nodes */
	CRS_VALUE_CONSTANT   = 64, /* f[0]                                        valuenode.c       */
	CRS_VALUE_GRAYSCALE  = 65, /* in0=color                                   grayscale.c:40-43 */
	CRS_VALUE_ALPHA      = 66, /* in0=color                                   alpha.c:38-41     */
	/* the node types only the reference's C constructors can build (no JSON path) */
	CRS_VALUE_MATH       = 67, /* in0=A in1=B (values), options = enum mathOp converter/math.c:44-97 */
	CRS_VALUE_FRESNEL    = 68, /* in0=IOR(value) in1=normal(vector, never evaluated)  input/fresnel.c:43-55 */
	CRS_VALUE_RAYLENGTH  = 69, /* hit distance                                input/raylength.c:36-40 */
	/* vector nodes: struct vectorValue { v, c, f } — every consumer reads .v only */
	CRS_VECTOR_CONSTANT  = 96, /* f[0..2]                                     vectornode.c:38-42 */
	CRS_VECTOR_NORMAL    = 97, /* surface normal of the hit                   input/normal.c:36-40 */
	CRS_VECTOR_VECMATH   = 98  /* in0=A in1=B (vectors), options = enum vecOp converter/vecmath.c:43-83 */
};
/* options of CRS_VALUE_MATH: enum mathOp in the reference's order (converter/math.h:11-27) */
enum crs_math_op { CRS_MATH_ADD, CRS_MATH_SUBTRACT, CRS_MATH_MULTIPLY, CRS_MATH_DIVIDE, CRS_MATH_POWER, CRS_MATH_LOG, CRS_MATH_SQRT,
                   CRS_MATH_ABS, CRS_MATH_MIN, CRS_MATH_MAX, CRS_MATH_SINE, CRS_MATH_COSINE, CRS_MATH_TANGENT, CRS_MATH_TO_RADIANS,
                   CRS_MATH_TO_DEGREES, CRS_MATH_OP_COUNT };
/* options of CRS_VECTOR_VECMATH: enum vecOp (converter/vecmath.h:11-22); DOT and LENGTH produce .f and leave .v = (0,0,0) */
enum crs_vec_op { CRS_VEC_ADD, CRS_VEC_SUBTRACT, CRS_VEC_MULTIPLY, CRS_VEC_AVERAGE, CRS_VEC_DOT, CRS_VEC_CROSS, CRS_VEC_NORMALIZE,
                  CRS_VEC_REFLECT, CRS_VEC_LENGTH, CRS_VEC_ABS, CRS_VEC_OP_COUNT };

/* image node options: reference src/nodes/textures/image.h */
#define CRS_IMG_SRGB_TRANSFORM 0x01u
#define CRS_IMG_NO_BILINEAR    0x02u

struct crs_node {           /* 64 bytes */
	int32_t  kind;          /* enum crs_node_kind */
	int32_t  in[3];         /* node indices, -1 = unused */
	float    f[8];          /* constants */
	int32_t  tex;           /* texture index for CRS_COLOR_IMAGE, -1 = NULL texture */
	uint32_t options;       /* CRS_IMG_* */
	uint32_t pad[2];
};

/* ---- textures (reference src/datatypes/image/texture.h:25-36) ---------------------------------- */
struct crs_texture {        /* 32 bytes */
	uint32_t width, height;
	uint32_t channels;      /* 1, 3 or 4 */
	uint32_t is_float;      /* precision == float_p */
	uint32_t has_alpha;
	uint32_t pad;
	uint64_t data_offset;   /* byte offset into texdata[], 16-byte aligned; row 0 = image TOP
	                           (storage row (H-1)-y, texture.c:24-28) */
};

/* ---- materials (the fields of reference `struct material` the hot path reads) ------------------ */
struct crs_material {       /* 32 bytes: material.h:62-83 */
	float   emission[4];    /* added at every hit, pathtrace.c:44 */
	float   IOR;            /* read by plastic.c:58-68 through record->material.IOR */
	int32_t bsdf;           /* root bsdf node index */
	uint32_t pad[2];
};

/* ---- geometry ---------------------------------------------------------------------------------- */
struct crs_poly {           /* 44 bytes: poly.h:11-18; indices are GLOBAL into vertices/normals/texcoords */
	int32_t  v[3];
	int32_t  n[3];
	int32_t  t[3];          /* -1 = no texture coordinate */
	uint32_t material;      /* index into the owning mesh's material range */
	uint32_t has_normals;
};

struct crs_bvh_node {       /* 32 bytes: bvh.c:37-42 (layout kept: minx,maxx,miny,maxy,minz,maxz) */
	float    bounds[6];
	uint32_t first_child_or_prim;
	uint32_t prim_count_leaf; /* bits 0..29 primCount, bit 30 isLeaf */
};
#define CRS_BVH_LEAF_BIT   (1u << 30)
#define CRS_BVH_COUNT_MASK ((1u << 30) - 1u)

struct crs_bvh {            /* one per mesh + one top level */
	uint32_t node_offset;   /* into bvh_nodes[] */
	uint32_t node_count;
	uint32_t prim_offset;   /* into prim_indices[] */
	uint32_t prim_count;
};

struct crs_mesh {           /* mesh.h:20-46 */
	uint32_t poly_offset;   /* into polys[] */
	uint32_t poly_count;
	uint32_t material_offset; /* into materials[] */
	uint32_t material_count;
	uint32_t texcoord_count;  /* mesh->textureCoordCount (0 => uv = (-1,-1), instance.c:151) */
	uint32_t bvh;             /* index into bvhs[] */
	float    ray_offset;      /* instance.c:227 */
	uint32_t pad;
};

struct crs_sphere {         /* sphere.h:15-19 */
	float    radius;
	float    ray_offset;    /* instance.c:106 */
	uint32_t material;      /* index into materials[] */
	uint32_t pad;
};

enum crs_instance_kind { CRS_INST_SPHERE = 0, CRS_INST_MESH = 1 };

struct crs_instance {       /* instance.h:23-28; matrices row-major 4x4 like transforms.h:21-23 */
	float    A[16];
	float    Ainv[16];
	uint32_t kind;          /* enum crs_instance_kind */
	uint32_t object;        /* sphere or mesh index */
	uint32_t pad[2];
};

struct crs_camera {         /* camera.h:15-33 — exactly the fields getCameraRay reads (camera.c:58-87) */
	float sensor_x, sensor_y;
	float aperture, focal_distance;
	float forward[3], right[3], up[3];
	int32_t width, height;
	uint32_t pad;
	float A[16];            /* composite.A */
};

struct crs_prefs {          /* renderer.h:58-87 (the subset the hot path and the dispatcher need) */
	uint32_t image_width, image_height;
	uint32_t sample_count, bounces;
	uint32_t tile_width, tile_height;
	uint32_t tile_order;    /* enum renderOrder, tile.h:15-21 */
	uint32_t thread_count;
};

/* ---- the whole scene ---------------------------------------------------------------------------- */
struct crs_scene {
	struct crs_prefs  prefs;
	struct crs_camera camera;
	int32_t  background;    /* root node of scene->background */
	uint32_t top_bvh;       /* index into bvhs[] of the top-level BVH (prims = instance indices) */

	uint32_t instance_count, sphere_count, mesh_count, material_count;
	uint32_t node_count, texture_count, bvh_count;
	uint32_t bvh_node_count, prim_index_count, poly_count;
	uint32_t vertex_count, normal_count, texcoord_count;
	uint64_t texdata_bytes;

	struct crs_instance *instances;
	struct crs_sphere   *spheres;
	struct crs_mesh     *meshes;
	struct crs_material *materials;
	struct crs_node     *nodes;
	struct crs_texture  *textures;
	struct crs_bvh      *bvhs;
	struct crs_bvh_node *bvh_nodes;
	int32_t             *prim_indices;
	struct crs_poly     *polys;
	float               *vertices;   /* 3 floats each: g_vertices, vertexbuffer.c:23 */
	float               *normals;    /* 3 floats each: g_normals */
	float               *texcoords;  /* 2 floats each: g_textureCoords */
	uint8_t             *texdata;

	void *owner;            /* backing allocation when loaded from a file (crscene_free) */
};

/* Serialisation (host C, c-ray_b200/host/crscene_io.c). Return 0 on success, negative on error. */
int  crscene_save(const struct crs_scene *s, const char *path);
int  crscene_load(struct crs_scene *out, const char *path);
void crscene_free(struct crs_scene *s);
/* -d WxH / -s N / bounces override on a loaded scene (camera.c:22-42 arithmetic); <= 0 keeps the value */
int  crscene_set_config(struct crs_scene *s, int width, int height, int samples, int bounces);


#ifdef __cplusplus
}
#endif
