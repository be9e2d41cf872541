#!/bin/bash
# round-2 GPU call D2 (1 GPU): the GPU suite again after the empty-pass-range fix (call D stopped at test 29 of the suite)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs --deselect tests/test_gpu_parity.py::test_host_c_renderer_multi_gpu_nccl_gather > $O/d2_pytest.log 2>&1; echo "pytest rc=$?" >> $O/d2_pytest.log
tail -25 $O/d2_pytest.log
