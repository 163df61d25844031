#!/bin/bash
# round 2: the entry points next to the main path — timing and ncu --set full of exact_kernel / scores_kernel / max_kernel / affinity_kernel
mkdir -p gpurun_out
timeout 600 python scripts/probe_aux.py > gpurun_out/r2K_probe_aux.log 2>&1; echo rc=$?; cat gpurun_out/r2K_probe_aux.log | tail -8
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'exact_kernel|scores_kernel|max_kernel|affinity_kernel' -c 4 -o gpurun_out/r2K_aux -f python scripts/probe_aux.py > gpurun_out/r2K_ncu.log 2>&1; echo "ncu rc=$?"
