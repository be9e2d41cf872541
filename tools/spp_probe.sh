#!/bin/bash
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref.json
nproc; lscpu | grep "Model name"
