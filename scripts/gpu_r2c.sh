#!/bin/bash
# round 2: parity tests (default kernel), then the warp-organised first stage (PWMSCAN_KM_KERNEL=2): parity subset, probes, lanes, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4.log 2>&1; tail -2 gpurun_out/probe_c4.log
echo "== v2 kernel"
PWMSCAN_KM_KERNEL=2 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_v2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/probe_c4_v2.log
PWMSCAN_KM_KERNEL=2 timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q > gpurun_out/pytest_v2.log 2>&1; echo "pytest v2 rc=$?"; tail -3 gpurun_out/pytest_v2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches_c4_chunk.csv python scripts/probe_c4.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 python scripts/probe_c4_genome.py 3 > gpurun_out/probe_c4_genome.log 2>&1; tail -2 gpurun_out/probe_c4_genome.log
PWMSCAN_KM_KERNEL=2 timeout 900 python scripts/probe_c4_genome.py 3 > gpurun_out/probe_c4_genome_v2.log 2>&1; tail -2 gpurun_out/probe_c4_genome_v2.log
