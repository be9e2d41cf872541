#!/bin/bash
# round-2 GPU call G (1 GPU): cooperative leaf phase (CRGPU_TRACE_DEFER=2) and K3 prefetch — parity under the variants, sweeps, ncu
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
CRGPU_TRACE_DEFER=2 CRGPU_SHADE_PREFETCH=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or bundled or invariants or tiling or union" > $O/g_pytest_defer2.log 2>&1; echo "pytest rc=$?" >> $O/g_pytest_defer2.log
( timeout 900 python tools/sweep.py hdr venus -- CRGPU_TRACE_DEFER=0,1,2 CRGPU_TRACE_BURST=2,3,4,6 ) > $O/g_sweep_defer_burst.txt 2>&1
( timeout 400 python tools/sweep.py hdr venus -- CRGPU_SHADE_PREFETCH=0,1 CRGPU_SHADE_MINB=3,4 ) > $O/g_sweep_prefetch.txt 2>&1
( timeout 200 python tools/sweep.py refraction -- CRGPU_TRACE_DEFER=0,2 CRGPU_SHADE_PREFETCH=0,1 ) > $O/g_sweep_refraction.txt 2>&1
CRGPU_TRACE_DEFER=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 1 -c 1 -o $O/g_prof_trace_hdr_defer2 -f python tools/render_once.py hdr 1920 1080 32 32 > $O/g_ncu1.log 2>&1
CRGPU_SHADE_PREFETCH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 3 -c 1 -o $O/g_prof_shade_hdr_prefetch -f python tools/render_once.py hdr 1920 1080 32 32 > $O/g_ncu2.log 2>&1
tail -5 $O/g_pytest_defer2.log; cat $O/g_sweep_defer_burst.txt $O/g_sweep_prefetch.txt $O/g_sweep_refraction.txt | cut -c1-250
