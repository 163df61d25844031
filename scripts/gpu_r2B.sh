#!/bin/bash
# round 2, final captures: ncu launch list of the default bench command, ncu --set full of the first stage and of the
# replay / ordering / pack kernels on one config-4 chunk, of cdf_kernel (800 CDFs), of the all-pairs alignment kernel (C5 sample)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2B_plain.log 2> gpurun_out/r2B_plain.err || { echo BENCH FAILED; tail -20 gpurun_out/r2B_plain.err; exit 3; }
grep '^{' gpurun_out/r2B_plain.log | tail -1 | cut -c1-300
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2B_launches_c4_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2B_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel_v2 -s 1 -c 1 -o gpurun_out/r2B_km2_c4 -f python scripts/probe_c4.py 3 > gpurun_out/r2B_ncu_km2.log 2>&1; echo "ncu km2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rescore_kernel|bucket_scatter_kernel|bucket_sort_kernel|bucket_sort_medium_kernel|pack_kernel|bucket_count_kernel|run_offsets_kernel' -c 12 -o gpurun_out/r2B_post_c4 -f python scripts/probe_c4.py 2 > gpurun_out/r2B_ncu_post.log 2>&1; echo "ncu post rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'align_all_kernel' -c 1 -o gpurun_out/r2B_align -f python scripts/bench_c5.py > gpurun_out/r2B_ncu_align.log 2>&1; echo "ncu align rc=$?"
timeout 300 python scripts/bench_c5.py > gpurun_out/r2B_c5.log 2>&1; tail -4 gpurun_out/r2B_c5.log
timeout 300 python scripts/probe_merge.py 2000 > gpurun_out/r2B_merge.log 2>&1; tail -3 gpurun_out/r2B_merge.log
ls -la gpurun_out/r2B*.ncu-rep
