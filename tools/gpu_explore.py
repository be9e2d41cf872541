"""Exploratory GPU run (gpurun): parity of the CUDA path vs the CPU oracle on the bundled scenes + first timings."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "c-ray_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crgpu, oracle_lib as O

SC = os.path.join(ROOT, "scenes", "_built")
cfgs = [("hdr", 240, 135, 16, 32), ("scene", 160, 100, 16, 50), ("refraction", 240, 135, 8, 512), ("venus", 100, 160, 16, 25)]
if len(sys.argv) > 1 and sys.argv[1] == "perf":
    cfgs = []
for name, W, H, spp, b in cfgs:
    path = os.path.join(SC, name + ".crscene")
    g = crgpu.GpuScene(path, W, H, spp, b)
    o = O.OracleScene(path, W, H, spp, b)
    rng = np.random.default_rng(1)
    xyp = np.stack([rng.integers(0, W, 512), rng.integers(0, H, 512), rng.integers(0, spp, 512)], 1).astype(np.int32)
    kg = np.frombuffer(g.trace_kat(xyp).tobytes(), dtype=O.HIT_KAT_DTYPE)
    ko = np.array([o.trace_kat(int(x), int(y), int(p)) for x, y, p in xyp])
    nbad = 0; fields = {}
    for f in O.HIT_KAT_DTYPE.names:
        if f == "pad": continue
        eq = np.all(np.asarray(kg[f]).reshape(512, -1).view(np.uint32) == np.asarray(ko[f]).reshape(512, -1).view(np.uint32), axis=1)
        fields[f] = int((~eq).sum())
    print(name, "KAT mismatching records per field:", fields, flush=True)
    t = time.time(); st = g.render_frame(flags=crgpu.FLAG_COUNT); dt = time.time() - t
    img = g.read()
    ref, ctr = o.render(threads=os.cpu_count(), count=True)
    d = img.astype(np.float64) - ref
    same = np.all(img.view(np.uint32) == ref.view(np.uint32), axis=2).mean()
    print(name, f"{W}x{H}x{spp} b{b}: rmse={np.sqrt((d**2).mean()):.3e} max={np.abs(d).max():.3e} identical_px={same:.4f} gpu_s={dt:.3f}")
    print("   gpu:", {k: st[k] for k in ("paths", "rays", "node_pairs", "tri_tests", "sphere_tests", "inst_visits", "kernel_launches")})
    print("   cpu:", {k: ctr[k] for k in ("paths", "rays", "node_pairs", "tri_tests", "sphere_tests", "inst_visits")}, flush=True)
    g.close(); o.close()

# perf probes
for name, W, H, spp, b, mp in [("hdr", 1920, 1080, 16, 32, 8 << 20), ("hdr", 1920, 1080, 16, 32, 32 << 20), ("venus", 2560, 1600, 8, 25, 16 << 20), ("refraction", 1920, 1080, 8, 512, 16 << 20)]:
    g = crgpu.GpuScene(os.path.join(SC, name + ".crscene"), W, H, spp, b, max_paths=mp)
    g.render_frame()  # warm
    g.clear()
    st = g.render_frame(flags=crgpu.FLAG_TIME_KERNELS)
    ms = st["total_ms"]
    print(f"PERF {name} {W}x{H}x{spp} b{b} maxpaths={mp>>20}M: total={ms:.1f}ms trace={st['trace_ms']:.1f} shade={st['shade_ms']:.1f} "
          f"Mray/s={st['rays']/ms/1e3:.1f} Msample/s={st['paths']/ms/1e3:.1f} rays/path={st['rays']/st['paths']:.2f} launches={st['kernel_launches']}", flush=True)
    g.close()
