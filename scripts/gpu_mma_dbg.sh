#!/bin/bash
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=mma
for w in 0 1 2; do for d in 0 13; do
  export PWMSCAN_MMA_DBG=$d PWMSCAN_MMA_WAIT=$w
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dbg.log 2> gpurun_out/bench_dbg.err || { echo BENCH FAILED; tail -5 gpurun_out/bench_dbg.err; }
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_dbg.log').read().strip().splitlines()[-1])
print('wait=$w dbg=$d','hits', d['config']['hits_per_step_rank0'], d['config']['device_ms_breakdown_last_step'])
"
done; done
