#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','msample_per_s','gpu_launches')}); print(d['e2e']['value'], d['roofline']['frac'], d['roofline']['trace_share_of_step'], d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench_n1.err
timeout 200 python tools/perf_probe.py venus refraction scene
