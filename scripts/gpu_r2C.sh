#!/bin/bash
# round 2: plain path with the second table gathered together with the first (PWMSCAN_KM_EAGER2=1) vs conditionally (0)
mkdir -p gpurun_out
for e in 0 1; do
  echo "eager2=$e"; PWMSCAN_KM_EAGER2=$e timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1 | cut -c1-160
  PWMSCAN_KM_EAGER2=$e timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2C_c2_e$e.log 2>&1
  grep '^{' gpurun_out/bench_r2C_c2_e$e.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['roofline']['ms_per_launch'])"
done
