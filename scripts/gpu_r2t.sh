#!/bin/bash
# round 2: event-driven host loop of scan_many, verdict copy behind the ordering; parity + bench (+ trace of the e2e arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py tests/test_gpu_apps.py -m gpu -x -q > gpurun_out/pytest_scan.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_scan.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_scan.log; }
BENCH_TRACE_RANK0=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2t.log 2> gpurun_out/bench_r2t.err; echo rc=$?
tail -1 gpurun_out/bench_r2t.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
grep trace gpurun_out/bench_r2t.err | tail -61 | sed -n '20,32p' | cut -c8-240
