#!/bin/bash
# round 2: heap check of the aux entry points (glibc MALLOC_CHECK_/MALLOC_PERTURB_), ncu of scores / max / affinity kernels
mkdir -p gpurun_out
MALLOC_CHECK_=3 MALLOC_PERTURB_=165 timeout 600 python scripts/probe_aux.py > gpurun_out/r2L_probe_aux_malloc_check.log 2>&1; echo "malloc-check rc=$?"; tail -3 gpurun_out/r2L_probe_aux_malloc_check.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scores_kernel|max_kernel|affinity_kernel' -c 3 -o gpurun_out/r2L_aux2 -f python scripts/probe_aux.py > gpurun_out/r2L_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2L_ncu.log | cut -c1-200
