#!/bin/bash
# round-2 GPU call C: instance-step batching / origin-cell sort / batch overlap sweeps + fresh ncu captures of K2 and K3
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
( timeout 600 python tools/sweep.py hdr venus refraction -- CRGPU_TRACE_INSTMIN=1,4,8,12,16 ) > gpurun_out/c_sweep_instmin.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_TRACE_SORT=0,3 CRGPU_OVERLAP=0,1 ) > gpurun_out/c_sweep_sort_overlap.txt 2>&1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c_bench_hdr.json 2> gpurun_out/c_bench_hdr.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/c_prof_trace_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > gpurun_out/c_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -o gpurun_out/c_prof_shade_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > gpurun_out/c_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o gpurun_out/c_prof_trace_venus -f python tools/render_once.py venus 2560 1600 16 25 > gpurun_out/c_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5; cat gpurun_out/c_sweep_instmin.txt gpurun_out/c_sweep_sort_overlap.txt; head -c 600 gpurun_out/c_bench_hdr.json
