#!/bin/bash
# round 2: cdf_kernel with test-free boundaries; bench with / without the stream split
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cdf.py -m gpu -x -q > gpurun_out/pytest_cdf.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_cdf.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_cdf.log; }
timeout 300 python scripts/probe_cdf.py 800 > gpurun_out/probe_cdf800.log 2>&1; tail -1 gpurun_out/probe_cdf800.log
timeout 300 python scripts/probe_cdf.py 100 > gpurun_out/probe_cdf100.log 2>&1; tail -1 gpurun_out/probe_cdf100.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cdf_kernel -c 1 -o gpurun_out/r2m_cdf -f python scripts/probe_cdf.py 800 > gpurun_out/ncu_cdf.log 2>&1; echo "ncu cdf rc=$?"
for v in split one; do
  if [ $v = one ]; then export PWMSCAN_ONE_STREAM=1; fi
  timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2m_$v.log 2> gpurun_out/bench_r2m_$v.err; echo "bench $v rc=$?"
  tail -1 gpurun_out/bench_r2m_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases'])"
done
unset PWMSCAN_ONE_STREAM
for lanes in 2 3 6; do
  PWMSCAN_LANES=$lanes timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2m_l$lanes.log 2> gpurun_out/bench_r2m_l$lanes.err
  tail -1 gpurun_out/bench_r2m_l$lanes.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes $lanes', d['ms_per_step'])"
done
