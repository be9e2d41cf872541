#!/bin/bash
# round-2 GPU call H (2 GPUs): full GPU suite with the f4 node interpreter + device BVH build, bench N=1 (K3 must not have slowed
# down) and N=2 (spatial tile interleave instead of queue-position dealing)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs > $O/h_pytest.log 2>&1; echo "pytest rc=$?" >> $O/h_pytest.log
timeout 300 python bench.py --no-cpu-baseline > $O/h_bench_hdr_n1.json 2> $O/h_bench_hdr_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > $O/h_bench_hdr_n2.json 2> $O/h_bench_hdr_n2.err
tail -12 $O/h_pytest.log | cut -c1-300; head -c 600 $O/h_bench_hdr_n1.json; echo; head -c 600 $O/h_bench_hdr_n2.json
