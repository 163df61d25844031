#!/bin/bash
# round 2: iterativeMerge with the warp-per-pair row refresh; all-pairs call with the near-ties re-evaluated by all host cores
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_apps.py tests/test_gpu_align.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/probe_merge.py 2000 2>&1 | tail -2
PWMSCAN_ALN_REFRESH_THREAD=1 timeout 300 python scripts/probe_merge.py 2000 2>&1 | tail -2
timeout 300 python scripts/bench_c5.py 2>&1 | tail -1 | cut -c1-420
