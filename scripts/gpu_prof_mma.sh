#!/bin/bash
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=mma
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo BENCH FAILED; tail -20 gpurun_out/plain.err; exit 3; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:mma_prefilter -s 1 -c 1 -o gpurun_out/r1c_mma_prefilter -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
