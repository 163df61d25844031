#!/bin/bash
# round 2: cdf_kernel without shared-memory tiles (certified fast index test); persistent first stage on the low-priority stream; bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cdf.py tests/test_gpu_apps.py -m gpu -x -q > gpurun_out/pytest_cdf.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_cdf.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_cdf.log; }
timeout 300 python scripts/probe_cdf.py 800 > gpurun_out/probe_cdf800.log 2>&1; tail -1 gpurun_out/probe_cdf800.log
timeout 300 python scripts/probe_cdf.py 100 > gpurun_out/probe_cdf100.log 2>&1; tail -1 gpurun_out/probe_cdf100.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cdf_kernel -c 1 -o gpurun_out/r2l_cdf -f python scripts/probe_cdf.py 800 > gpurun_out/ncu_cdf.log 2>&1; echo "ncu cdf rc=$?"
BENCH_TRACE_RANK0=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2l.log 2> gpurun_out/bench_r2l.err; echo "bench rc=$?"; grep -v trace gpurun_out/bench_r2l.err | tail -3
tail -1 gpurun_out/bench_r2l.log | cut -c1-1200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
