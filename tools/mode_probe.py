import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "c-ray_b200"))
import crgpu, shard
W, H, spp, b = 1920, 1080, 256, 32
g = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", "hdr.crscene"), W, H, spp, b)
rects = shard.rank_rects(W, H, 64, 0, 1)
for name, flags in [("warm", 0), ("sync", 0), ("time", crgpu.FLAG_TIME_KERNELS), ("sync", 0), ("async", crgpu.FLAG_ASYNC), ("async", crgpu.FLAG_ASYNC), ("time", crgpu.FLAG_TIME_KERNELS)]:
    g.clear()
    t = time.perf_counter()
    st = g.render_tiles(rects, flags=flags)
    if st is None:
        st = g.get_stats()
    dt = time.perf_counter() - t
    print(f"{name:6s} wall {dt*1e3:8.1f} ms  events total {st['total_ms']:.1f} trace {st['trace_ms']:.1f} shade {st['shade_ms']:.1f}  Mray/s(wall) {st['rays']/dt/1e6:.0f}", flush=True)
g.close()
