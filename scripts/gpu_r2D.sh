#!/bin/bash
# round 2: bitmap as the predicate of the plain loop (PWMSCAN_FORCE_BITMAP=2) next to the scan + hit-queue path (1) and no bitmap (0)
mkdir -p gpurun_out
PWMSCAN_FORCE_BITMAP=2 timeout 600 python scripts/probe_c4_blocks.py > gpurun_out/probe_c4_blocks_pred.log 2>&1; echo rc=$?; cat gpurun_out/probe_c4_blocks_pred.log
echo "auto:"; timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1 | cut -c1-160
for e in 0 2; do
  PWMSCAN_FORCE_BITMAP=$e timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2D_c2_b$e.log 2>&1
  grep '^{' gpurun_out/bench_r2D_c2_b$e.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 force_bitmap=$e', d['ms_per_step'], d['roofline']['ms_per_launch'], d['config'].get('parity_checked'))"
done
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -2
