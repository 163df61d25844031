#!/bin/bash
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=mma
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_mma.log 2>&1; rc=$?
tail -3 gpurun_out/smoke_mma.log | cut -c1-300
[ $rc -eq 0 ] || { echo SMOKE FAILED; exit 1; }
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest_mma.log 2>&1; rc=$?
tail -15 gpurun_out/pytest_mma.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; exit 2; }
for v in mma lds; do
  export PWMSCAN_PREFILTER=$v
  PWMSCAN_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err || { echo BENCH FAILED; tail -20 gpurun_out/bench_$v.err; }
  grep "mma block" gpurun_out/bench_$v.err | sort | uniq | head -4
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1])
print('$v','value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'hits', d['config']['hits_per_step_rank0'], d['config']['device_ms_breakdown_last_step'])
"
done
