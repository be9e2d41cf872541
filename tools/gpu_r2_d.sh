#!/bin/bash
# round-2 GPU call D (1 GPU): the whole state re-measured after the container was re-created — GPU suite incl. the full-config
# RMSE, bench lines of C2/C3/C4/C5(reduced spp), ncu launch list, ncu --set full of K2/K3, kernel-variant sweeps
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $O/d_smi.txt 2>&1
nproc > $O/d_host.txt; cat /sys/fs/cgroup/cpu.max >> $O/d_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_host_c_renderer_multi_gpu_nccl_gather > $O/d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d_pytest.log
timeout 300 python bench.py > $O/d_bench_hdr.json 2> $O/d_bench_hdr.err
timeout 300 python bench.py --workload venus --steps 2 --warmup 1 > $O/d_bench_venus.json 2> $O/d_bench_venus.err
timeout 400 python bench.py --workload refraction --steps 2 --warmup 1 > $O/d_bench_refraction.json 2> $O/d_bench_refraction.err
timeout 300 python bench.py --workload hdr8k --spp 250 --steps 2 --warmup 1 --no-cpu-baseline > $O/d_bench_hdr8k_250spp.json 2> $O/d_bench_hdr8k.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/d_launches_hdr128.csv python bench.py --spp 128 --steps 1 --warmup 1 --no-cpu-baseline > $O/d_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o $O/d_prof_trace_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > $O/d_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -o $O/d_prof_shade_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > $O/d_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o $O/d_prof_trace_venus -f python tools/render_once.py venus 2560 1600 16 25 > $O/d_ncu3.log 2>&1
( timeout 600 python tools/sweep.py hdr venus refraction -- CRGPU_TRACE_SORT=0,1,2,3 CRGPU_TRACE_DEFER=0,1 ) > $O/d_sweep_sort_defer.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_TRACE_STAGE=0,128,512,1024 ) > $O/d_sweep_stage.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_SHADE_SPLIT=0,1 CRGPU_SHADE_MINB=2,3 ) > $O/d_sweep_shade.txt 2>&1
( timeout 300 python tools/sweep.py hdr venus -- CRGPU_TRACE_INSTMIN=1,8 CRGPU_OVERLAP=0,1 ) > $O/d_sweep_inst_overlap.txt 2>&1
tail -15 $O/d_pytest.log; head -c 1200 $O/d_bench_hdr.json; cat $O/d_sweep_sort_defer.txt
