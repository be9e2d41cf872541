#!/usr/bin/env python3
"""Where does the end-to-end overhead of a fresh scene per frame go?  Times every phase of
crscene_load -> crgpu_scene_create -> first render (allocates the wavefront) -> second render -> read -> destroy
on one GPU, for a share of the hdr.json frame.  usage: e2e_breakdown.py [W H spp]   (default 1920 1080 1000)"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "c-ray_b200")]
import crgpu
W, H, spp = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1920, 1080, 1000)
L = crgpu.lib()
path = os.path.join(ROOT, "scenes", "_built", "hdr.crscene")
for rep in range(2):
    t = [time.perf_counter()]
    flat = crgpu.FlatScene()
    assert L.crscene_load(C.byref(flat), path.encode()) == 0
    L.crscene_set_config(C.byref(flat), W, H, spp, 32)
    t.append(time.perf_counter())
    h = C.c_void_p()
    assert L.crgpu_scene_create(C.byref(flat), 0, C.byref(h)) == 0, L.crgpu_last_error()
    t.append(time.perf_counter())
    L.crscene_free(C.byref(flat))
    st = crgpu.Stats()
    assert L.crgpu_render_tile(h, 0, 0, W, H, 0, spp, 0, C.byref(st)) == 0
    t.append(time.perf_counter())
    assert L.crgpu_render_tile(h, 0, 0, W, H, 0, spp, 0, C.byref(st)) == 0
    t.append(time.perf_counter())
    out = np.empty((H, W, 3), dtype=np.float32)
    assert L.crgpu_framebuffer_read(h, out.ctypes.data, 0, 0, 0, 0) == 0
    t.append(time.perf_counter())
    L.crgpu_scene_destroy(h)
    t.append(time.perf_counter())
    names = ["crscene_load", "crgpu_scene_create", "render #1 (allocates wavefront)", "render #2", "framebuffer_read", "scene_destroy"]
    print(f"rep {rep}: " + "; ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in zip(names, t, t[1:])) + f"; rays {st.rays}")
