#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan.py -m gpu -x -q -k "scan_many or host or c4" > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
PWMSCAN_KM_KERNEL=2 timeout 900 python scripts/probe_c4_genome.py 3 4 5 6 > gpurun_out/probe_c4_genome_v2.log 2>&1; tail -8 gpurun_out/probe_c4_genome_v2.log
