#!/bin/bash
# ncu evidence for the current kernels (run under gpurun; outputs to gpurun_out/)
mkdir -p gpurun_out
timeout 60 python tools/render_once.py hdr 1920 1080 32 32 > gpurun_out/render_once.log 2>&1; cat gpurun_out/render_once.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/launches_hdr.csv python tools/render_once.py hdr 1920 1080 32 32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 2 -o gpurun_out/prof_trace_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -o gpurun_out/prof_shade_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bucket -s 1 -c 1 -o gpurun_out/prof_bucket_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > /dev/null 2>&1
ls -la gpurun_out | tail -8
