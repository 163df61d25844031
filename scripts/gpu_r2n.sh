#!/bin/bash
# round 2: per-block first-stage cost of a config-4 chunk, cold and warm L2; 16 Mbp chunk for comparison
mkdir -p gpurun_out
timeout 600 python scripts/probe_c4_blocks.py > gpurun_out/probe_c4_blocks.log 2>&1; echo rc=$?; cat gpurun_out/probe_c4_blocks.log
timeout 600 python scripts/probe_c4_blocks.py 16000000 > gpurun_out/probe_c4_blocks16.log 2>&1; echo rc=$?; cat gpurun_out/probe_c4_blocks16.log
