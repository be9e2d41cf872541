"""How do tile size / paths-in-flight affect throughput?  (coherence of many passes of a small tile vs few passes of the whole frame)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "c-ray_b200"))
import crgpu
for name, W, H, spp, b in [("hdr", 1920, 1080, 64, 32), ("venus", 2560, 1600, 32, 25)]:
    for tile, mp in [(0, 8 << 20), (0, 32 << 20), (256, 4 << 20), (128, 1 << 20), (64, 1 << 18), (64, 1 << 20), (32, 1 << 16)]:
        if tile == 0 and W * H > mp: continue
        g = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", name + ".crscene"), W, H, spp, b, max_paths=mp)
        rects = [(x, y, min(x + tile, W), min(y + tile, H)) for y in range(0, H, tile) for x in range(0, W, tile)] if tile else [(0, 0, W, H)]
        import time
        for rep in range(2):
            g.clear(); g.get_stats(); t = time.perf_counter()
            for r in rects: g.render_tile(*r, flags=crgpu.FLAG_ASYNC)
            st = g.get_stats(); dt = time.perf_counter() - t
        print(f"{name} tile={tile or 'frame'} maxpaths={mp>>10}K: {dt*1e3:.1f} ms  Mray/s={st['rays']/dt/1e6:.1f} launches={st['kernel_launches']}", flush=True)
        g.close()
