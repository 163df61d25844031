#!/bin/bash
# parity tests -> verbose plan info -> one ncu --set full capture of the prefilter kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -15 gpurun_out/pytest.log
[ $rc -eq 0 ] || echo PYTEST FAILED
PWMSCAN_VERBOSE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2> gpurun_out/plain.err; rc=$?
grep pwmscan gpurun_out/plain.err | sort | uniq -c | head -20
cat gpurun_out/plain.log
[ $rc -eq 0 ] || { echo BENCH FAILED; tail -20 gpurun_out/plain.err; exit 3; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:prefilter -s 1 -c 2 -o gpurun_out/prof_prefilter -f \
    python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
