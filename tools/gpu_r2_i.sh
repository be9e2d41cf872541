#!/bin/bash
# round-2 GPU call I (1 GPU): the final binary re-measured — GPU suite, bench lines C2/C3/C4/C5(250 spp), paths-in-flight sweep,
# ncu launch list + ncu --set full of K2 / K3 (hdr) and K2 (venus)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nproc > $O/i_host.txt; cat /sys/fs/cgroup/cpu.max >> $O/i_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rs > $O/i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/i_pytest.log
timeout 300 python bench.py > $O/i_bench_hdr.json 2> $O/i_bench_hdr.err
timeout 300 python bench.py --workload venus --steps 2 --warmup 1 > $O/i_bench_venus.json 2> $O/i_bench_venus.err
timeout 400 python bench.py --workload refraction --steps 2 --warmup 1 > $O/i_bench_refraction.json 2> $O/i_bench_refraction.err
timeout 300 python bench.py --workload hdr8k --spp 250 --steps 2 --warmup 1 --no-cpu-baseline > $O/i_bench_hdr8k_250spp.json 2> $O/i_bench_hdr8k.err
for mp in 134217728 536870912; do timeout 200 python bench.py --max-paths $mp --steps 2 --warmup 1 --no-cpu-baseline > $O/i_bench_hdr_maxpaths_$mp.json 2>/dev/null; done
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $O/i_bench_reference_arm.json 2> $O/i_bench_reference_arm.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/i_launches_hdr128.csv python bench.py --spp 128 --steps 1 --warmup 1 --no-cpu-baseline > $O/i_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o $O/i_prof_trace_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > $O/i_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -o $O/i_prof_shade_hdr -f python tools/render_once.py hdr 1920 1080 32 32 > $O/i_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o $O/i_prof_trace_venus -f python tools/render_once.py venus 2560 1600 16 25 > $O/i_ncu3.log 2>&1
tail -8 $O/i_pytest.log | cut -c1-200; head -c 700 $O/i_bench_hdr.json; echo; for mp in 134217728 536870912; do head -c 200 $O/i_bench_hdr_maxpaths_$mp.json; echo; done
