#!/bin/bash
# k-mer mask first stage: parity tests, then the C2 bench against the shared-memory kernel
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=kmer
timeout 1200 python -m pytest tests/test_gpu_scan.py tests/test_gpu_apps.py -m gpu -x -q > gpurun_out/pytest_km.log 2>&1; rc=$?
tail -5 gpurun_out/pytest_km.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_km.log; exit 2; }
for v in kmer lds; do
  export PWMSCAN_PREFILTER=$v
  PWMSCAN_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.log 2> gpurun_out/bench_$v.err || { echo BENCH FAILED; tail -20 gpurun_out/bench_$v.err; }
  grep "kmer block\|\] block" gpurun_out/bench_$v.err | sort -u | head -8
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1])
print('$v','value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'prefilter ms',round(d['roofline']['ms_per_launch'],3), 'hits', d['config']['hits_per_step_rank0'])
"
done
