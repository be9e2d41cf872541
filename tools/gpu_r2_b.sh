#!/bin/bash
# round-2 GPU call B: full GPU suite (no -x) + kernel-variant sweeps
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_host_c_renderer_multi_gpu_nccl_gather > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
( timeout 900 python tools/sweep.py hdr venus refraction -- CRGPU_TRACE_SORT=0,1,2 CRGPU_TRACE_DEFER=0,1 CRGPU_SHADE_SPLIT=0,1 ) > gpurun_out/b_sweep_main.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_SHADE_MINB=2,3,4 ) > gpurun_out/b_sweep_shade_minb.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_TRACE_STAGE=0,128,512,1024 ) > gpurun_out/b_sweep_stage.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_TRACE_MINB=2,3,4 CRGPU_TRACE_BURST=3,4,6 ) > gpurun_out/b_sweep_trace_minb.txt 2>&1
tail -12 gpurun_out/b_pytest.log; cat gpurun_out/b_sweep_main.txt
