#!/bin/bash
# round-2 GPU call J (8 GPUs): strong scaling of C2 with the spatial tile interleave (final binary), N = 8, 4, 2, 1 on the same box
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/j_bench_hdr_n8.json 2> $O/j_bench_hdr_n8.err
timeout 200 $TR --nproc-per-node 4 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $O/j_bench_hdr_n4.json 2> $O/j_bench_hdr_n4.err
timeout 200 $TR --nproc-per-node 2 --master-port 29543 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/j_bench_hdr_n2.json 2> $O/j_bench_hdr_n2.err
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/j_bench_hdr_n1.json 2> $O/j_bench_hdr_n1.err
for f in n1 n2 n4 n8; do head -c 400 $O/j_bench_hdr_$f.json; echo; done
