#!/bin/bash
# round 2: IUPAC-aware search through the first stage + ambig_kernel: parity, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/probe_aux.py 2>&1 | head -3 | cut -c1-200
