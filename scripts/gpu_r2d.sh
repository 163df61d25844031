#!/bin/bash
# round 2: ordering v3 + lane-owned host planes: quick parity subset, probes (v1/v2 first stage), full ncu of the v2 kernel on a C4 chunk, genome runs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4.log 2>&1; tail -1 gpurun_out/probe_c4.log
PWMSCAN_KM_KERNEL=2 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_v2.log 2>&1; tail -1 gpurun_out/probe_c4_v2.log
PWMSCAN_KM_KERNEL=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2d_launches_c4_chunk.csv python scripts/probe_c4.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
PWMSCAN_KM_KERNEL=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel_v2 -s 1 -c 1 -o gpurun_out/r2d_km2_c4 -f python scripts/probe_c4.py 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
PWMSCAN_KM_KERNEL=2 timeout 900 python scripts/probe_c4_genome.py 3 > gpurun_out/probe_c4_genome_v2.log 2>&1; tail -2 gpurun_out/probe_c4_genome_v2.log
