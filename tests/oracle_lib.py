"""ctypes view of oracle/libcray_oracle.so (the CPU restatement) — test infrastructure only."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Prefs(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("image_width", "image_height", "sample_count", "bounces",
                                          "tile_width", "tile_height", "tile_order", "thread_count")]


class Camera(C.Structure):
    _fields_ = [("sensor_x", C.c_float), ("sensor_y", C.c_float), ("aperture", C.c_float),
                ("focal_distance", C.c_float), ("forward", C.c_float * 3), ("right", C.c_float * 3),
                ("up", C.c_float * 3), ("width", C.c_int32), ("height", C.c_int32), ("pad", C.c_uint32),
                ("A", C.c_float * 16)]


class Scene(C.Structure):
    """struct crs_scene (include/crscene.h)"""
    _fields_ = [("prefs", Prefs), ("camera", Camera), ("background", C.c_int32), ("top_bvh", C.c_uint32),
                ("instance_count", C.c_uint32), ("sphere_count", C.c_uint32), ("mesh_count", C.c_uint32),
                ("material_count", C.c_uint32), ("node_count", C.c_uint32), ("texture_count", C.c_uint32),
                ("bvh_count", C.c_uint32), ("bvh_node_count", C.c_uint32), ("prim_index_count", C.c_uint32),
                ("poly_count", C.c_uint32), ("vertex_count", C.c_uint32), ("normal_count", C.c_uint32),
                ("texcoord_count", C.c_uint32), ("texdata_bytes", C.c_uint64)] + \
               [(n, C.c_void_p) for n in ("instances", "spheres", "meshes", "materials", "nodes", "textures",
                                          "bvhs", "bvh_nodes", "prim_indices", "polys", "vertices", "normals",
                                          "texcoords", "texdata", "owner")]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("paths", "rays", "node_pairs", "tri_tests", "sphere_tests",
                                          "inst_visits", "draws", "max_depth", "max_stack", "max_pairs_ray")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


HIT_KAT_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("pixIdx", "<i4"), ("instIndex", "<i4"), ("polyIndex", "<i4"),
                          ("o", "<f4", 3), ("d", "<f4", 3), ("distance", "<f4"), ("uv", "<f4", 2),
                          ("hitPoint", "<f4", 3), ("normal", "<f4", 3), ("emission", "<f4", 3),
                          ("out", "<f4", 3), ("color", "<f4", 4), ("nextDraw", "<f4"), ("pad", "<f4", 9)])
assert HIT_KAT_DTYPE.itemsize == 160

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "libcray_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libcray_oracle.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.crscene_load.argtypes = [C.POINTER(Scene), C.c_char_p]
        L.crscene_load.restype = C.c_int
        L.crscene_free.argtypes = [C.POINTER(Scene)]
        L.crscene_set_config.argtypes = [C.POINTER(Scene)] + [C.c_int] * 4
        L.cro_render.argtypes = [C.POINTER(Scene)] + [C.c_int] * 6 + [C.c_void_p, C.c_int, C.POINTER(Counters)]
        L.cro_render.restype = C.c_int
        L.cro_sampler_kat.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.cro_trace_kat.argtypes = [C.POINTER(Scene), C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.cro_to_srgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


class OracleScene:
    def __init__(self, path, width=0, height=0, samples=0, bounces=0):
        self.s = Scene()
        rc = lib().crscene_load(C.byref(self.s), path.encode())
        if rc != 0:
            raise RuntimeError(f"crscene_load({path}) failed: {rc}")
        lib().crscene_set_config(C.byref(self.s), width, height, samples, bounces)
        self.W = self.s.prefs.image_width
        self.H = self.s.prefs.image_height

    def render(self, threads=8, tile=None, passes=None, rgb=None, count=False):
        """Full frame (or tile=(x0,y0,x1,y1), passes=(begin,count)); returns fp32 (H,W,3) y-flipped like renderBuffer."""
        if rgb is None:
            rgb = np.zeros((self.H, self.W, 3), dtype=np.float32)
        x0, y0, x1, y1 = tile if tile else (0, 0, self.W, self.H)
        pb, pc = passes if passes else (0, self.s.prefs.sample_count)
        ctr = Counters()
        lib().cro_render(C.byref(self.s), x0, y0, x1, y1, pb, pc, rgb.ctypes.data, threads,
                         C.byref(ctr) if count else None)
        return (rgb, ctr.as_dict()) if count else rgb

    def trace_kat(self, x, y, p):
        out = np.zeros(1, dtype=HIT_KAT_DTYPE)
        lib().cro_trace_kat(C.byref(self.s), x, y, p, out.ctypes.data)
        return out[0]

    def close(self):
        lib().crscene_free(C.byref(self.s))


def sampler_kat(pix, p, maxp, n):
    out = np.zeros(n, dtype=np.float32)
    lib().cro_sampler_kat(pix, p, maxp, n, out.ctypes.data)
    return out
