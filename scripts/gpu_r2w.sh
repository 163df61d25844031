#!/bin/bash
# round 2: packed genome store — parity tests, then the bench (packed + ASCII end-to-end arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest_scan.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_scan.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_scan.log; exit 1; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2w.log 2> gpurun_out/bench_r2w.err; echo rc=$?; tail -3 gpurun_out/bench_r2w.err
tail -1 gpurun_out/bench_r2w.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e packed ms',round(d['e2e']['ms_per_step'],3),'e2e ascii ms',round(d['e2e']['from_ascii']['ms_per_step'],3), d['e2e']['h2d_bytes_per_step'], d['e2e']['input'][-120:])"
