#!/bin/bash
# round 2: per-item trace of the end-to-end arm on one GPU
mkdir -p gpurun_out
BENCH_TRACE_RANK0=1 timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2s.log 2> gpurun_out/bench_r2s.err; echo rc=$?
grep trace gpurun_out/bench_r2s.err | tail -30 | cut -c1-240
