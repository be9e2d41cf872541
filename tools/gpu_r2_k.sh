#!/bin/bash
# round-2 GPU call K (1 GPU): A/B of keeping the world ray's slab set-up in local memory (CRGPU_TRACE_REUSE_SETUP) on one box,
# and the GPU parity subset on the final binary
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 600 python tools/sweep.py hdr venus refraction -- CRGPU_TRACE_REUSE_SETUP=0,1,0,1 ) > $O/k_sweep_reuse_setup.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_bvh_build.py -m gpu -q -x -k "not multi_gpu and not torchrun" > $O/k_pytest.log 2>&1; echo "pytest rc=$?" >> $O/k_pytest.log
cut -c1-200 $O/k_sweep_reuse_setup.txt; tail -3 $O/k_pytest.log
