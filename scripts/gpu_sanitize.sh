#!/bin/bash
# compute-sanitizer memcheck on the smallest end-to-end case (smoke), after a plain run exited 0
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || { echo SMOKE FAILED; tail -20 gpurun_out/smoke.log; exit 1; }
tail -1 gpurun_out/smoke.log | cut -c1-200
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E 'ERROR SUMMARY|Invalid|out of bounds' gpurun_out/memcheck.log | head -10; tail -3 gpurun_out/memcheck.log
