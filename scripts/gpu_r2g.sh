#!/bin/bash
# round 2 (re-entry): smoke, full parity suite, one-chunk probes (first stage v1/v2), launch list, default bench (C4), l2gather roof micro-benchmark
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest.log 2>&1; rc=$?
tail -25 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4.log 2>&1; tail -1 gpurun_out/probe_c4.log
PWMSCAN_KM_KERNEL=2 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_v2.log 2>&1; tail -1 gpurun_out/probe_c4_v2.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2g.log 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2g.err
tail -1 gpurun_out/bench_r2g.log | cut -c1-2500
PWMSCAN_KM_KERNEL=2 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2g_v2.log 2> gpurun_out/bench_r2g_v2.err; echo "bench v2 rc=$?"
tail -1 gpurun_out/bench_r2g_v2.log | cut -c1-1200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2g_launches_c4_chunk.csv python scripts/probe_c4.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
for b in l2gather l2gather_L; do
  if [ -x experiments/masktab/$b ]; then timeout 120 experiments/masktab/$b > gpurun_out/${b}_stdout.txt 2>&1; echo "$b rc=$?"; fi
done
