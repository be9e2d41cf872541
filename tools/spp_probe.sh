#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','msample_per_s','gpu_launches','clocks')}); print(d['e2e']); print(d['roofline']); print(d.get('cpu_baseline'))"; tail -3 gpurun_out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --spp 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu rc=$?"
