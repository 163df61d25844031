#!/bin/bash
# round 2: planner with longest-first blocks + dense-bitmap penalty; unit sizes per block on small chunks; parity
mkdir -p gpurun_out
timeout 600 python scripts/probe_c4.py 4 > gpurun_out/probe_c4_p.log 2>&1; tail -1 gpurun_out/probe_c4_p.log
for bp in 16000000 32000000; do
  for u in 512 1024 2048; do
    echo "bp=$bp unit=$u"; PROBE_BP=$bp PWMSCAN_KM_UNIT=$u timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1
  done
done
timeout 1500 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest_scan.log 2>&1; rc=$?
tail -3 gpurun_out/pytest_scan.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest_scan.log; }
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2p.log 2> gpurun_out/bench_r2p.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_r2p.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases'], d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'])"
