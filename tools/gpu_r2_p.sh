#!/bin/bash
# round-2 GPU call P (2 GPUs): first frame vs steady state of the in-process 2-worker renderFrame (static spatial shares)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
cd oracle/_ref && timeout 200 python - > ../../gpurun_out/p_frames.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "../../c-ray_b200")
import crhost
for gpus in (2, 1):
    R = crhost.Renderer("input/hdr.json", 1920, 1080, 1000, 32, gpus=gpus, tile=64, quiet=True)
    t0 = time.perf_counter(); R.prepare(); print("gpus", gpus, "prepare %.3f s" % (time.perf_counter() - t0))
    for i in range(3):
        t0 = time.perf_counter(); secs, rays = R.render(); print("gpus", gpus, "frame", i, "renderFrame %.3f s (wall %.3f)" % (secs, time.perf_counter() - t0), rays)
    R.close()
PY
cat ../../gpurun_out/p_frames.txt
