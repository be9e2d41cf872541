#!/usr/bin/env python3
"""Per-pixel RMSE of the GPU frame against the CPU reference AT THE FULL BENCH CONFIG (BASELINE.json: hdr.json 1920x1080,
1000 spp, 32 bounces, bound 1e-4).  The CPU side is the oracle on all host cores (bit-identical to the strict reference build:
tests/test_oracle.py), ~2e9 samples = a few minutes on the GPU box's 128 threads — too slow for the test suite, so it is a tool.
usage: full_rmse.py [scene W H spp bounces]      prints one JSON line; exit 1 if RMSE > 1e-4"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "c-ray_b200"), os.path.join(ROOT, "tests")]
import crgpu
import oracle_lib as O

scene, W, H, spp, b = (sys.argv[1], *map(int, sys.argv[2:6])) if len(sys.argv) >= 6 else ("hdr", 1920, 1080, 1000, 32)
path = os.path.join(ROOT, "scenes", "_built", scene + ".crscene")
t0 = time.time()
g = crgpu.GpuScene(path, W, H, spp, b)
st = g.render_frame()
gpu = g.read()
g.close()
t1 = time.time()
o = O.OracleScene(path, W, H, spp, b)
cpu = o.render(threads=os.cpu_count())
o.close()
t2 = time.time()
d = gpu.astype(np.float64) - cpu
rmse = float(np.sqrt(np.mean(d * d)))
print(json.dumps({"scene": scene, "config": f"{W}x{H} {spp} spp {b} bounces", "rmse": rmse, "max_abs_diff": float(np.abs(d).max()),
                  "pixels_bit_identical": float((gpu.view(np.uint32) == cpu.view(np.uint32)).all(axis=2).mean()),
                  "mean_radiance": float(cpu.mean()), "gpu_seconds": round(t1 - t0, 2), "oracle_seconds": round(t2 - t1, 1),
                  "oracle_threads": os.cpu_count(), "rays": int(st["rays"])}))
sys.exit(0 if rmse <= 1e-4 else 1)
