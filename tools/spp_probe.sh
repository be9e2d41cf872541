#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/render_once.py venus 2560 1600 8 25
python tools/render_once.py refraction 1920 1080 8 512
python tools/render_once.py scene 1920 1200 8 50
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 2 -o gpurun_out/prof_trace_hdr_flat -f python tools/render_once.py hdr 1920 1080 4 32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 1 -c 2 -o gpurun_out/prof_shade_hdr2 -f python tools/render_once.py hdr 1920 1080 4 32 > /dev/null 2>&1
