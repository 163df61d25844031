#!/bin/bash
# ncu evidence for the dominant kernel: launch list (shares) + one --set full capture, after a plain run
mkdir -p gpurun_out
TAG=${TAG:-r1}
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo BENCH FAILED; tail -20 gpurun_out/plain.err; exit 3; }
tail -1 gpurun_out/plain.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prefilter -s 3 -c 2 -o gpurun_out/${TAG}_prefilter -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
