"""Env-var sweep of kernel variants: every configuration runs tools/perf_probe.py-style in its own process (the variants are read
once per process) and prints one line: workload, env, frame ms, Mray/s, trace ms, shade ms, frame CRC (results must not move).

  python tools/sweep.py hdr venus -- CRGPU_TRACE_SORT=0,1,2 CRGPU_SHADE_SPLIT=0,1
"""
import itertools
import json
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W_ = {"hdr": ("hdr", 1920, 1080, 128, 32), "venus": ("venus", 2560, 1600, 64, 25), "refraction": ("refraction", 1920, 1080, 64, 512),
      "scene": ("scene", 1920, 1200, 128, 50)}

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, os.path.join(ROOT, "c-ray_b200"))
    import time
    import crgpu
    import shard
    name, W, H, spp, b = W_[sys.argv[2]]
    g = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", name + ".crscene"), W, H, spp, b)
    rects = shard.rank_rects(W, H, 64, 0, 1)
    best = None
    for rep in range(3):
        g.clear()
        t = time.perf_counter()
        st = g.render_tiles(rects, flags=crgpu.FLAG_TIME_KERNELS if rep == 2 else 0)
        dt = time.perf_counter() - t
        if rep == 1:
            best = dt
    crc = zlib.crc32(g.read().tobytes())
    print(json.dumps({"workload": sys.argv[2], "ms": round(best * 1e3, 1), "mray_s": round(st["rays"] / best / 1e6, 1), "trace_ms": round(st["trace_ms"], 1),
                      "shade_ms": round(st["shade_ms"], 1), "total_ms_timed": round(st["total_ms"], 1), "launches": st["kernel_launches"], "crc": f"{crc:08x}"}))
    g.close()
    sys.exit(0)

args = sys.argv[1:]
split = args.index("--") if "--" in args else len(args)
workloads, axes = args[:split] or ["hdr"], args[split + 1:]
names = [a.split("=")[0] for a in axes]
values = [a.split("=")[1].split(",") for a in axes]
for wl in workloads:
    for combo in itertools.product(*values) if axes else [()]:
        env = dict(os.environ)
        env.update(dict(zip(names, combo)))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", wl], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        tag = " ".join(f"{n}={v}" for n, v in zip(names, combo)) or "(defaults)"
        line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else "FAILED rc=%d %s" % (r.returncode, r.stderr[-300:].replace("\n", " | "))
        print(f"{tag:60s} {line}", flush=True)
