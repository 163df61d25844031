#!/bin/bash
# round 2: ncu --set full of the first stage (v2 and v1) on a C4 chunk, of the replay/ordering kernels, the pack kernel; C2 bench both first stages; 800 CDFs; l2gather with 256-bit loads
mkdir -p gpurun_out
timeout 120 experiments/masktab/l2gather_v8 > gpurun_out/l2gather_v8_stdout.txt 2>&1; echo "l2gather_v8 rc=$?"
for w in 1 2; do
  PWMSCAN_KM_KERNEL=$w timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_k$w.log 2> gpurun_out/bench_c2_k$w.err; echo "bench c2 k$w rc=$?"
  tail -1 gpurun_out/bench_c2_k$w.log | cut -c1-330
done
timeout 300 python scripts/probe_cdf.py 800 > gpurun_out/probe_cdf800.log 2>&1; tail -1 gpurun_out/probe_cdf800.log
PWMSCAN_KM_KERNEL=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel_v2 -s 1 -c 1 -o gpurun_out/r2h_km2_c4 -f python scripts/probe_c4.py 3 > gpurun_out/ncu_full_km2.log 2>&1; echo "ncu km2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel -s 1 -c 1 -o gpurun_out/r2h_km1_c4 -f python scripts/probe_c4.py 3 > gpurun_out/ncu_full_km1.log 2>&1; echo "ncu km1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rescore_kernel|bucket_scatter_kernel|bucket_sort_kernel|bucket_sort_medium_kernel|pack_kernel' -c 9 -o gpurun_out/r2h_post_c4 -f python scripts/probe_c4.py 2 > gpurun_out/ncu_full_post.log 2>&1; echo "ncu post rc=$?"
ls -la gpurun_out/*.ncu-rep
