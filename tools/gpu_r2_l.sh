#!/bin/bash
# round-2 GPU call L (1 GPU): compute-sanitizer memcheck over the device BVH build (synthetic inputs incl. 100k boxes) and over the
# wavefront kernels on the golden scenes (known-answer records + framebuffers, incl. the f4 node interpreter), racecheck on one render
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bvh_build.py -m gpu -q -x -k "synthetic" > $O/l_memcheck_bvh.log 2>&1; echo "rc=$?" >> $O/l_memcheck_bvh.log
timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "known_answer or golden or tiling" > $O/l_memcheck_render.log 2>&1; echo "rc=$?" >> $O/l_memcheck_render.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden and (g_nodes or g_f4)" > $O/l_racecheck_render.log 2>&1; echo "rc=$?" >> $O/l_racecheck_render.log
tail -6 $O/l_memcheck_bvh.log; tail -6 $O/l_memcheck_render.log; tail -6 $O/l_racecheck_render.log
