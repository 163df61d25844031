#!/bin/bash
# round 2: per-block first-stage cost without the first-table bitmap (plain path everywhere), 64 Mbp chunk
mkdir -p gpurun_out
PWMSCAN_FORCE_BITMAP=0 timeout 600 python scripts/probe_c4_blocks.py > gpurun_out/probe_c4_blocks_nobm.log 2>&1; echo rc=$?; cat gpurun_out/probe_c4_blocks_nobm.log
PWMSCAN_KM_KERNEL=1 timeout 600 python scripts/probe_c4_blocks.py > gpurun_out/probe_c4_blocks_k1.log 2>&1; echo rc=$?; cat gpurun_out/probe_c4_blocks_k1.log
