#!/bin/bash
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=mma
for d in 1 5 9 13; do
  export PWMSCAN_MMA_DBG=$d
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_dbg$d.log 2> gpurun_out/bench_dbg$d.err || { echo BENCH FAILED; tail -5 gpurun_out/bench_dbg$d.err; }
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_dbg$d.log').read().strip().splitlines()[-1])
print('dbg=$d','hits', d['config']['hits_per_step_rank0'], d['config']['device_ms_breakdown_last_step'])
"
done
