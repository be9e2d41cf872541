#!/usr/bin/env python3
"""Quick GPU check of the JSON-loader path and the c-ray API binary (no build step, no pytest): used when GPU time is short.
  1. tests/golden/g_nodes.json, g_meshmat.json, g_single.json -> libcrloader -> GPU, RMSE vs the reference framebuffer
  2. oracle/_ref/cray_main_b200 (reference main.c + libcrhost.so) on input/hdr.json 96x54x4 -> PNG == python path
"""
import os, re, subprocess, sys, time, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "c-ray_b200")]
import crgpu

def decode_png(path):
    raw = open(path, "rb").read()
    pos, idat, W, H = 8, b"", 0, 0
    while pos < len(raw):
        n = int.from_bytes(raw[pos:pos + 4], "big"); tag = raw[pos + 4:pos + 8]; data = raw[pos + 8:pos + 8 + n]
        if tag == b"IHDR": W, H = int.from_bytes(data[:4], "big"), int.from_bytes(data[4:8], "big")
        if tag == b"IDAT": idat += data
        pos += 12 + n
    a = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(H, W * 3 + 1)
    return a[:, 1:].reshape(H, W, 3)

ok = True
gold = os.path.join(ROOT, "tests", "golden")
os.chdir(gold)
for name in ("g_nodes", "g_meshmat", "g_single"):
    t0 = time.time()
    g = crgpu.GpuScene(name + ".json")
    g.render_frame()
    got = g.read()
    ref = np.fromfile(os.path.join(gold, name + ".f32"), dtype=np.float32).reshape(got.shape)
    rmse = float(np.sqrt(np.mean((got.astype(np.float64) - ref) ** 2)))
    g.close()
    print(f"{name}.json -> GPU: RMSE vs reference {rmse:.3e} ({time.time() - t0:.2f}s)")
    ok &= rmse <= 1e-4

refdir = os.path.join(ROOT, "oracle", "_ref")
exe = os.path.join(refdir, "cray_main_b200")
if os.path.exists(exe):
    r = subprocess.run([exe, "input/hdr.json", "-d", "96x54", "-s", "4", "-t", "32x32"], cwd=refdir, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=100)
    print(r.stdout[-600:])
    m = re.search(r'Saving result in "([^"]+)"', r.stdout)
    if r.returncode == 0 and m:
        png = decode_png(os.path.join(refdir, m.group(1)))
        os.chdir(refdir)
        t0 = time.time()
        g = crgpu.GpuScene("input/hdr.json", 96, 54, 4)
        t1 = time.time()
        g.render_frame()
        same = np.array_equal(png, g.srgb8())
        # the same scene through the reference-exported .crscene must give the same fp32 frame
        a = g.read().copy()
        g.close()
        g2 = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", "hdr.crscene"), 96, 54, 4, 30)
        g2.render_frame()
        same2 = np.array_equal(a.view(np.uint32), g2.read().view(np.uint32))
        g2.close()
        print(f"reference main.c on the B200 path: PNG identical to python path: {same}; json vs crscene frame identical: {same2}; scene create {t1 - t0:.2f}s")
        ok &= same and same2
    else:
        ok = False
print("LOADER-GPU-CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
