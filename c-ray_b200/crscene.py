"""numpy view of the flat scene (include/crscene.h) and ctypes binding of the scene loader (include/crloader.h).

`load_json(path)` runs libcrloader.so (c-ray_b200/host/loader/, plain C: JSON + OBJ/MTL + PNG/HDR + BVH build) and
returns a FlatScene that `crgpu.GpuScene.from_flat` can upload; `arrays(scene)` exposes the 14 arrays as structured
numpy arrays (copies), which is what tests/test_loader.py compares against the reference's exported scenes.
"""
import ctypes as C
import os

import numpy as np

from crgpu import FlatScene

HERE = os.path.dirname(os.path.abspath(__file__))
LOADER_PATH = os.path.join(HERE, "libcrloader.so")

INSTANCE = np.dtype([("A", "<f4", 16), ("Ainv", "<f4", 16), ("kind", "<u4"), ("object", "<u4"), ("pad", "<u4", 2)])
SPHERE = np.dtype([("radius", "<f4"), ("ray_offset", "<f4"), ("material", "<u4"), ("pad", "<u4")])
MESH = np.dtype([("poly_offset", "<u4"), ("poly_count", "<u4"), ("material_offset", "<u4"), ("material_count", "<u4"),
                 ("texcoord_count", "<u4"), ("bvh", "<u4"), ("ray_offset", "<f4"), ("pad", "<u4")])
MATERIAL = np.dtype([("emission", "<f4", 4), ("IOR", "<f4"), ("bsdf", "<i4"), ("pad", "<u4", 2)])
NODE = np.dtype([("kind", "<i4"), ("in", "<i4", 3), ("f", "<f4", 8), ("tex", "<i4"), ("options", "<u4"), ("pad", "<u4", 2)])
TEXTURE = np.dtype([("width", "<u4"), ("height", "<u4"), ("channels", "<u4"), ("is_float", "<u4"), ("has_alpha", "<u4"),
                    ("pad", "<u4"), ("data_offset", "<u8")])
BVH = np.dtype([("node_offset", "<u4"), ("node_count", "<u4"), ("prim_offset", "<u4"), ("prim_count", "<u4")])
BVH_NODE = np.dtype([("bounds", "<f4", 6), ("first", "<u4"), ("count_leaf", "<u4")])
POLY = np.dtype([("v", "<i4", 3), ("n", "<i4", 3), ("t", "<i4", 3), ("material", "<u4"), ("has_normals", "<u4")])
assert (INSTANCE.itemsize, SPHERE.itemsize, MESH.itemsize, MATERIAL.itemsize, NODE.itemsize, TEXTURE.itemsize,
        BVH.itemsize, BVH_NODE.itemsize, POLY.itemsize) == (144, 16, 32, 32, 64, 32, 16, 32, 44)

LEAF_BIT = 1 << 30
COUNT_MASK = LEAF_BIT - 1

_SECTIONS = [("instances", INSTANCE, "instance_count", 1), ("spheres", SPHERE, "sphere_count", 1),
             ("meshes", MESH, "mesh_count", 1), ("materials", MATERIAL, "material_count", 1),
             ("nodes", NODE, "node_count", 1), ("textures", TEXTURE, "texture_count", 1), ("bvhs", BVH, "bvh_count", 1),
             ("bvh_nodes", BVH_NODE, "bvh_node_count", 1), ("prim_indices", np.dtype("<i4"), "prim_index_count", 1),
             ("polys", POLY, "poly_count", 1), ("vertices", np.dtype("<f4"), "vertex_count", 3),
             ("normals", np.dtype("<f4"), "normal_count", 3), ("texcoords", np.dtype("<f4"), "texcoord_count", 2),
             ("texdata", np.dtype("u1"), "texdata_bytes", 1)]


def arrays(scene):
    """dict name -> numpy copy of each array of a FlatScene (vertices/normals as (n,3), texcoords as (n,2))."""
    out = {}
    for name, dtype, count_field, width in _SECTIONS:
        n = int(getattr(scene, count_field)) * width
        ptr = getattr(scene, name)
        if n == 0 or not ptr:
            a = np.zeros(0, dtype=dtype)
        else:
            a = np.frombuffer(C.string_at(ptr, n * dtype.itemsize), dtype=dtype).copy()
        out[name] = a.reshape(-1, width) if width > 1 else a
    return out


_loader = None


def loader():
    global _loader
    if _loader is None:
        if not os.path.exists(LOADER_PATH):
            raise RuntimeError(f"{LOADER_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LOADER_PATH)
        L.crloader_load_json.argtypes = [C.POINTER(FlatScene), C.c_char_p]
        L.crloader_load_json.restype = C.c_int
        L.crloader_last_error.restype = C.c_char_p
        _loader = L
    return _loader


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def load_json(path):
    """Parse a c-ray JSON scene (+ its assets) into a FlatScene.  Raises RuntimeError with the loader's message."""
    s = FlatScene()
    rc = loader().crloader_load_json(C.byref(s), os.fsencode(path))
    if rc != 0:
        raise RuntimeError(f"crloader_load_json({path}) failed ({rc}): {loader().crloader_last_error().decode()}")
    return s


def free(scene):
    """Release a FlatScene returned by load_json."""
    if scene.owner:
        _libc.free(scene.owner)
        scene.owner = None
