#!/bin/bash
# round-2 GPU call M (1 GPU): last parameter sweep of K2's refill threshold and instance-step batching on the three scenes
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
O=gpurun_out
( timeout 500 python tools/sweep.py hdr venus refraction -- CRGPU_TRACE_REFILL=12,16,20,24 ) > $O/m_sweep_refill.txt 2>&1
( timeout 300 python tools/sweep.py venus refraction -- CRGPU_TRACE_INSTMIN=4,8,12 ) > $O/m_sweep_instmin.txt 2>&1
cut -c1-200 $O/m_sweep_refill.txt $O/m_sweep_instmin.txt
