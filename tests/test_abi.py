"""CPU tests of the C-ABI boundary: libcrgpu.so loads without a GPU, exports every symbol declared in
include/crgpu.h / include/crscene.h, and fails loudly (no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import crgpu
from conftest import ROOT, GOLDEN


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crgpu_\w+|crscene_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = crgpu.lib()
    names = declared_symbols("crgpu.h") + declared_symbols("crscene.h")
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libcrgpu.so does not export {n}"


def test_nccl_library_exports_every_declared_symbol():
    """include/crgpu_nccl.h -> libcrgpu_nccl.so.  Checked in a subprocess: the library links the system NCCL, and a process
    that later imports torch (with its own bundled NCCL) must never have both loaded — the reason it is a separate .so."""
    import subprocess
    import sys
    names = declared_symbols("crgpu_nccl.h")
    assert names == ["crgpu_comm_create", "crgpu_comm_create_rank", "crgpu_comm_destroy", "crgpu_comm_gather_tiles",
                     "crgpu_comm_gather_tiles_rank", "crgpu_comm_set_stream", "crgpu_comm_unique_id"]
    code = ("import ctypes as C, sys\n"
            "C.CDLL(sys.argv[1] + '/libcrgpu.so', mode=C.RTLD_GLOBAL)\n"
            "L = C.CDLL(sys.argv[1] + '/libcrgpu_nccl.so')\n"
            "[getattr(L, n) for n in sys.argv[2:]]\n"
            "assert L.crgpu_comm_create(None, 0, None) != 0\n"      # bad arguments are rejected before any NCCL call
            "assert L.crgpu_comm_create_rank(None, 0, 1, 0, None) != 0 and L.crgpu_comm_unique_id(None) != 0\n"
            "print('NCCL-ABI-OK')\n")
    r = subprocess.run([sys.executable, "-c", code, os.path.join(ROOT, "c-ray_b200")] + names, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert "NCCL-ABI-OK" in r.stdout, r.stdout[-2000:]


def test_struct_layouts_match_header():
    assert C.sizeof(crgpu.Stats) == 7 * 8 + 4 * 4
    assert C.sizeof(crgpu.Prefs) == 32 and C.sizeof(crgpu.Camera) == 128


def test_crscene_roundtrip(tmp_path):
    L = crgpu.lib()
    L.crscene_save.argtypes = [C.POINTER(crgpu.FlatScene), C.c_char_p]
    s = crgpu.FlatScene()
    src = os.path.join(GOLDEN, "g_legacy.crscene")
    assert L.crscene_load(C.byref(s), src.encode()) == 0
    dst = tmp_path / "copy.crscene"
    assert L.crscene_save(C.byref(s), str(dst).encode()) == 0
    assert open(src, "rb").read() == open(dst, "rb").read()
    L.crscene_free(C.byref(s))


def test_crscene_load_rejects_garbage(tmp_path):
    L = crgpu.lib()
    p = tmp_path / "bad.crscene"
    p.write_bytes(b"not a scene" * 100)
    s = crgpu.FlatScene()
    assert L.crscene_load(C.byref(s), str(p).encode()) != 0
    assert L.crscene_load(C.byref(s), b"/nonexistent/file") != 0


def test_set_config_matches_camera_c():
    L = crgpu.lib()
    s = crgpu.FlatScene()
    assert L.crscene_load(C.byref(s), os.path.join(GOLDEN, "g_nodes.crscene").encode()) == 0
    sx = np.float32(s.camera.sensor_x)
    L.crscene_set_config(C.byref(s), 1920, 1080, 1000, 32)
    assert (s.prefs.image_width, s.prefs.image_height, s.prefs.sample_count, s.prefs.bounces) == (1920, 1080, 1000, 32)
    aspect = np.float32(1920) / np.float32(1080)
    assert np.float32(s.camera.sensor_y) == sx / aspect          # camera.c:29-31
    L.crscene_set_config(C.byref(s), 0, 0, 0, 0)                  # <= 0 keeps the values
    assert s.prefs.sample_count == 1000 and s.camera.width == 1920
    L.crscene_free(C.byref(s))


def test_no_cpu_fallback_without_device():
    """Without a GPU the product path must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(crgpu.CrgpuError):
        crgpu.device_count()
    with pytest.raises(crgpu.CrgpuError):
        crgpu.GpuScene(os.path.join(GOLDEN, "g_single.crscene"))


def test_product_does_not_import_oracle():
    """Nothing under c-ray_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "c-ray_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".c", ".h", ".cu", ".cuh", ".py", "Makefile")):
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r'#include\s+".*oracle|import\s+oracle|libcray_oracle|cro_render|from\s+oracle', t):
                    bad.append(f)
    assert not bad, bad
