#!/bin/bash
# quick iteration: parity tests + bench (verbose plan)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -5 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -40 gpurun_out/pytest.log; exit 2; }
for v in ${S1_VARIANTS:-""}; do
  echo "== FORCE_S1=$v"
  PWMSCAN_FORCE_S1=$v PWMSCAN_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err || { echo BENCH FAILED; tail -20 gpurun_out/bench_$v.err; }
  grep pwmscan gpurun_out/bench_$v.err | sort | uniq | head -5
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',d['e2e'],'ms/step',round(d['ms_per_step'],3),'prefilter ms',round(d['roofline']['ms_per_launch'],3),'clocks',d['clocks'])
print(d['config'].get('device_ms_breakdown_last_scan')); print('cpu', d.get('cpu_baseline'))
"
done
