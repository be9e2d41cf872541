"""The headline parity number: per-pixel RMSE of the GPU frame against the CPU reference AT THE BENCH CONFIGURATION
(BASELINE.json: input/hdr.json 1920x1080, 1000 spp, 32 bounces, bound 1e-4, same PCG seeds).  The CPU side is the oracle on all
host cores — bit-identical to the strict reference build at this image size (tests/test_oracle.py) — about 2e9 samples, i.e.
under a minute on the GPU box's 128 threads.  Collected last (file name) so that a slow host never starves the other tests."""
import json
import os
import time

import numpy as np
import pytest

from conftest import BUILT, ROOT


@pytest.mark.gpu
def test_full_config_rmse_hdr_1080p_1000spp():
    import crgpu
    import oracle_lib as O
    scene = os.path.join(BUILT, "hdr.crscene")
    if not os.path.exists(scene):
        pytest.skip("scenes/_built missing")
    if (os.cpu_count() or 1) < 32:
        pytest.skip("the CPU side needs ~2e9 samples: only run where the host has the cores for it")
    W, H, spp, b = 1920, 1080, 1000, 32
    t0 = time.time()
    g = crgpu.GpuScene(scene, W, H, spp, b)
    st = g.render_frame()
    gpu = g.read()
    g.close()
    t1 = time.time()
    o = O.OracleScene(scene, W, H, spp, b)
    cpu = o.render(threads=os.cpu_count())
    o.close()
    t2 = time.time()
    d = gpu.astype(np.float64) - cpu
    rmse = float(np.sqrt(np.mean(d * d)))
    out = {"config": f"hdr.json {W}x{H} {spp} spp {b} bounces", "rmse": rmse, "max_abs_diff": float(np.abs(d).max()),
           "pixels_bit_identical": float((gpu.view(np.uint32) == cpu.view(np.uint32)).all(axis=2).mean()),
           "rays": int(st["rays"]), "gpu_seconds": round(t1 - t0, 2), "oracle_seconds": round(t2 - t1, 1), "oracle_threads": os.cpu_count()}
    print("FULL-CONFIG-RMSE " + json.dumps(out))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "full_config_rmse.json"), "w") as f:
            f.write(json.dumps(out) + "\n")
    except OSError:
        pass
    assert np.isfinite(gpu).all()
    assert rmse <= 1e-4, out
