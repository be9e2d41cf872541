"""Render one frame of a bundled scene once (for ncu captures): python tools/render_once.py hdr 1920 1080 4 32 [tile] [maxpaths]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "c-ray_b200"))
import crgpu
name, W, H, spp, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
tile = int(sys.argv[6]) if len(sys.argv) > 6 else 0
mp = int(sys.argv[7]) if len(sys.argv) > 7 else None
g = crgpu.GpuScene(os.path.join(ROOT, "scenes", "_built", name + ".crscene"), W, H, spp, b, max_paths=mp)
st = g.render_frame(tile=(tile, tile) if tile else None, flags=int(os.environ.get('CRGPU_FLAGS', str(crgpu.FLAG_TIME_KERNELS))))
print(name, W, H, spp, b, "tile", tile, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()},
      "Mray/s %.1f" % (st["rays"] / st["total_ms"] / 1e3))
g.close()
