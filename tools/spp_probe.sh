#!/bin/bash
mkdir -p gpurun_out
run() { # name envs...
  env "$@" timeout 200 python bench.py --workload ${WL:-hdr} --spp 256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bq.json 2>gpurun_out/bq.err; python -c "
import json;d=json.load(open('gpurun_out/bq.json'));print('$* ${WL:-hdr}', d['value'], d['ms_per_step'], d['roofline']['trace_gray_per_s'], d['roofline']['trace_share_of_step'])"; tail -1 gpurun_out/bq.err; }
run CRGPU_TRACE_BURST=1 CRGPU_TRACE_REFILL=16
run CRGPU_TRACE_BURST=2 CRGPU_TRACE_REFILL=16
run CRGPU_TRACE_BURST=2 CRGPU_TRACE_REFILL=8
run CRGPU_TRACE_BURST=1 CRGPU_TRACE_REFILL=8
run CRGPU_TRACE_BURST=2 CRGPU_TRACE_REFILL=12
run CRGPU_TRACE_BURST=3 CRGPU_TRACE_REFILL=16
WL=venus run CRGPU_TRACE_BURST=2 CRGPU_TRACE_REFILL=16
WL=venus run CRGPU_TRACE_BURST=4 CRGPU_TRACE_REFILL=24
WL=venus run CRGPU_TRACE_BURST=2 CRGPU_TRACE_REFILL=8
