#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
timeout 60 python tools/render_once.py hdr 1920 1080 64 32; echo "rc=$?"
timeout 60 python tools/render_once.py venus 2560 1600 8 25; echo "rc=$?"
timeout 60 python tools/render_once.py scene 1920 1200 8 50; echo "rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 2 -o gpurun_out/prof_trace_hdr_p -f python tools/render_once.py hdr 1920 1080 4 32 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -o gpurun_out/prof_shade_hdr_p -f python tools/render_once.py hdr 1920 1080 4 32 > /dev/null 2>&1
