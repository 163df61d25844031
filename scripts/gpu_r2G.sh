#!/bin/bash
# round 2: where does a SPARSE block of the first stage spend its time?  ncu --set full of one 128-motif block (P(first mask != 0) = 0.3 %)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel_v2 -s 1 -c 1 -o gpurun_out/r2G_km2_sparse -f python scripts/probe_c4_oneblock.py 1 > gpurun_out/r2G_ncu.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/r2G_ncu.log
