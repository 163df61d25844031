#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apps.py -m gpu -x -q 2>&1 | tail -3
timeout 1500 python scripts/bench_c3.py > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err || { echo C3 FAILED; tail -20 gpurun_out/bench_c3.err; }
cat gpurun_out/bench_c3.log
