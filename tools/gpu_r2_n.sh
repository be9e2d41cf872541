#!/bin/bash
# round-2 GPU call N (1 GPU): the GPU suite on the final tree (after the PNG text labels and the K2 revert), smoke()
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs > $O/n_pytest.log 2>&1; echo "pytest rc=$?" >> $O/n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/n_smoke.log 2>&1; echo "smoke rc=$?" >> $O/n_smoke.log
tail -12 $O/n_pytest.log | cut -c1-200; tail -3 $O/n_smoke.log
