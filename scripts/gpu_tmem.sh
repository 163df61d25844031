#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -5 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -40 gpurun_out/pytest.log; exit 2; }
for v in tmem notmem; do
  if [ $v = notmem ]; then export PWMSCAN_NO_TMEM=1; else unset PWMSCAN_NO_TMEM; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err || { echo BENCH FAILED; tail -20 gpurun_out/bench_$v.err; }
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1])
print('$v','value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'prefilter ms',round(d['roofline']['ms_per_launch'],3), 'hits', d['config']['hits_per_step_rank0'])
"
done
