#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python scripts/bench_c3.py > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err || { echo C3 FAILED; tail -20 gpurun_out/bench_c3.err; }
cat gpurun_out/bench_c3.log
timeout 900 python scripts/bench_c5.py > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err || { echo C5 FAILED; tail -20 gpurun_out/bench_c5.err; }
cat gpurun_out/bench_c5.log
