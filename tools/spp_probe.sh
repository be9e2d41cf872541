#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 200 python tools/perf_probe.py hdr venus refraction
echo "--- shade minb 3"; CRGPU_SHADE_MINB=3 timeout 100 python tools/perf_probe.py hdr venus
echo "--- shade minb 4"; CRGPU_SHADE_MINB=4 timeout 100 python tools/perf_probe.py hdr
