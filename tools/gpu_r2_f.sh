#!/bin/bash
# round-2 GPU call F (8 GPUs): -j 2/4/8 + 2/4/8-rank bit-identity, bench N=2/4/8 on C2, C5 (hdr.json 7680x4320, 4000 spp) on 8 GPUs,
# the C dispatcher's own -j 8 timing, and the device BVH build tests
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi -L > $O/f_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or torchrun or bvh" > $O/f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > $O/f_bench_hdr_n8.json 2> $O/f_bench_hdr_n8.err
timeout 200 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 3 --warmup 3 > $O/f_bench_hdr_n4.json 2> $O/f_bench_hdr_n4.err
timeout 200 $TR --nproc-per-node 2 --master-port 29524 bench.py --gpus 2 --steps 3 --warmup 3 > $O/f_bench_hdr_n2.json 2> $O/f_bench_hdr_n2.err
timeout 500 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload hdr8k --steps 2 --warmup 1 > $O/f_bench_hdr8k_n8.json 2> $O/f_bench_hdr8k_n8.err
( cd oracle/_ref && timeout 200 ../../c-ray_b200/cray_b200 input/hdr.json -d 1920x1080 -s 1000 -b 32 -j 8 -o /tmp/f_j8.png ) > $O/f_cli_hdr_j8.txt 2>&1
tail -6 $O/f_pytest.log; for f in n8 n4 n2; do head -c 700 $O/f_bench_hdr_$f.json; echo; done; head -c 900 $O/f_bench_hdr8k_n8.json; tail -4 $O/f_cli_hdr_j8.txt
