#!/bin/bash
# round 2: work units drawn in batches (PWMSCAN_KM_BATCH) — config-4 chunk of 64 and 16 Mbp, C2
mkdir -p gpurun_out
for b in 1 2 4 8; do
  echo "batch=$b"
  PWMSCAN_KM_BATCH=$b timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1 | cut -c1-120
  PROBE_BP=16000000 PWMSCAN_KM_BATCH=$b timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1 | cut -c1-120
done
for b in 1 4; do
  PWMSCAN_KM_BATCH=$b timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2H_c2_b$b.log 2>&1
  grep '^{' gpurun_out/bench_r2H_c2_b$b.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 batch=$b', d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
