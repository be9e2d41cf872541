#!/bin/bash
# One gpurun call that re-validates on a B200 box everything that changed after the last full GPU run of round 1
# (no GPU time was left): the widened GPU suite, the host-C multi-GPU renderer against one GPU, and a short bench line.
#   gpurun --gpus 2 --timeout 1500 -- 'tools/gpu_recheck.sh 2>&1 | tail -60'        (use --gpus 8 to repeat the -j 8 case)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu (with the five extra bundled scenes)"
CRAY_GPU_EXTRA=1 timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
NGPU=$(nvidia-smi -L | wc -l)
echo "== host-C renderer: -j 1 vs -j N (N=$NGPU), hdr.json 1920x1080 256 spp, per-run fp32 dumps compared bit for bit"
if [ -f oracle/_ref/input/hdr.json ]; then
  ( cd oracle/_ref && timeout 600 ../../c-ray_b200/cray_b200 input/hdr.json -d 1920x1080 -s 256 -b 32 -t 64x64 -j 1 --dump-f32 /tmp/j1.f32 -q; echo "rc(-j 1)=$?" )
  for n in 2 4 8; do
    [ "$n" -le "$NGPU" ] || continue
    ( cd oracle/_ref && /usr/bin/time -f "-j $n wall %es" timeout 900 ../../c-ray_b200/cray_b200 input/hdr.json -d 1920x1080 -s 256 -b 32 -t 64x64 -j $n --dump-f32 /tmp/j$n.f32 -q; echo "rc(-j $n)=$?" )
    python - "$n" <<'PY'
import sys, numpy as np
n = sys.argv[1]
a = np.fromfile("/tmp/j1.f32", dtype=np.uint32).reshape(1080, 1920, 3)
b = np.fromfile(f"/tmp/j{n}.f32", dtype=np.uint32).reshape(1080, 1920, 3)
bad = np.argwhere((a != b).any(axis=2))
print(f"-j {n}: identical to -j 1: {len(bad) == 0}" + ("" if len(bad) == 0 else f"; {len(bad)} pixels differ, rows {bad[:,0].min()}..{bad[:,0].max()} cols {bad[:,1].min()}..{bad[:,1].max()}"))
PY
  done
fi
echo "== full-config RMSE (hdr 1920x1080 1000 spp: GPU vs oracle on all host cores; minutes)"
timeout 900 python tools/full_rmse.py 2>&1 | tail -1
echo "== bench (short)"
timeout 600 python bench.py --steps 2 --warmup 3 --spp 200 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-600
