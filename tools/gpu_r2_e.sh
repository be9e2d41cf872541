#!/bin/bash
# round-2 GPU call E (2 GPUs): the multi-GPU product paths — host-C `-j 2` (threads + NCCL gather in C) and torchrun ranks —
# bit-identity tests, then bench lines at N=2 (C2) and the C dispatcher's own timing
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi -L > $O/e_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or torchrun or two_gpus" > $O/e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/e_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/e_bench_hdr_n2.json 2> $O/e_bench_hdr_n2.err
( cd oracle/_ref && for j in 1 2; do timeout 200 ../../c-ray_b200/cray_b200 input/hdr.json -d 1920x1080 -s 1000 -b 32 -j $j -o /tmp/e_j$j.png --dump-f32 /tmp/e_j$j.f32; done ) > $O/e_cli_hdr.txt 2>&1
cmp /tmp/e_j1.f32 /tmp/e_j2.f32 >> $O/e_cli_hdr.txt 2>&1 && echo "-j1 == -j2 (fp32 frame bytes)" >> $O/e_cli_hdr.txt
tail -6 $O/e_pytest.log; head -c 900 $O/e_bench_hdr_n2.json; tail -8 $O/e_cli_hdr.txt
