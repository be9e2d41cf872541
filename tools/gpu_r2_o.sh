#!/bin/bash
# round-2 GPU call O (2 GPUs): in-process workers (-j N) with the static spatial share: bit-identity -j 1 / -j 2 and CLI timing on C2
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu or cli" > $O/o_pytest.log 2>&1; echo "pytest rc=$?" >> $O/o_pytest.log
( cd oracle/_ref && for j in 1 2; do timeout 100 ../../c-ray_b200/cray_b200 input/hdr.json -d 1920x1080 -s 1000 -b 32 -j $j -o /tmp/o_j$j.png --dump-f32 /tmp/o_j$j.f32; done ) > $O/o_cli_hdr.txt 2>&1
cmp /tmp/o_j1.f32 /tmp/o_j2.f32 >> $O/o_cli_hdr.txt 2>&1 && echo "-j1 == -j2 (fp32 frame bytes)" >> $O/o_cli_hdr.txt
tail -4 $O/o_pytest.log; grep -E "Finished|==" $O/o_cli_hdr.txt
