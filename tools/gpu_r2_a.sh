#!/bin/bash
# round-2 GPU call A: full GPU suite + bench lines of every workload + ncu launch list (1 GPU)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_host.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/a_host.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_host_c_renderer_multi_gpu_nccl_gather > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python bench.py > gpurun_out/a_bench_hdr.json 2> gpurun_out/a_bench_hdr.err
timeout 300 python bench.py --workload venus --steps 2 --warmup 1 > gpurun_out/a_bench_venus.json 2> gpurun_out/a_bench_venus.err
timeout 400 python bench.py --workload refraction --steps 2 --warmup 1 > gpurun_out/a_bench_refraction.json 2> gpurun_out/a_bench_refraction.err
timeout 300 python bench.py --workload hdr8k --spp 250 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_bench_hdr8k_250spp.json 2> gpurun_out/a_bench_hdr8k.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches_hdr128.csv python bench.py --spp 128 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_bench_under_ncu.log 2>&1
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_bench_hdr.json | head -c 1500
