#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 60 python tools/render_once.py hdr 1920 1080 4 32; echo "rc=$?"
timeout 60 python tools/render_once.py hdr 1920 1080 64 32; echo "rc=$?"
timeout 60 python tools/render_once.py venus 2560 1600 8 25; echo "rc=$?"
timeout 60 python tools/render_once.py refraction 1920 1080 8 512; echo "rc=$?"
timeout 60 python tools/render_once.py scene 1920 1200 8 50; echo "rc=$?"
