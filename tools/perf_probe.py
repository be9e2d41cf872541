"""Wall-clock throughput of one synchronous crgpu_render_tiles call per workload (second of two runs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "c-ray_b200"))
import crgpu, shard
W_ = {"hdr": ("hdr", 1920, 1080, 256, 32), "venus": ("venus", 2560, 1600, 64, 25), "refraction": ("refraction", 1920, 1080, 64, 512),
      "scene": ("scene", 1920, 1200, 128, 50)}
for key in (sys.argv[1:] or list(W_)):
    name, W, H, spp, b = W_[key]
    g = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", name + ".crscene"), W, H, spp, b)
    rects = shard.rank_rects(W, H, 64, 0, 1)
    for rep in range(2):
        g.clear()
        t = time.perf_counter()
        st = g.render_tiles(rects, flags=crgpu.FLAG_TIME_KERNELS if rep else 0)
        dt = time.perf_counter() - t
    print(f"{key:10s} {W}x{H}x{spp} b{b}: {dt*1e3:8.1f} ms  {st['rays']/dt/1e6:8.1f} Mray/s {st['paths']/dt/1e6:8.1f} Msample/s  trace {st['trace_ms']:.1f} shade {st['shade_ms']:.1f} launches {st['kernel_launches']}", flush=True)
    g.close()
