#!/bin/bash
# k-mer mask first stage: table-count variants, then one ncu --set full capture of the kernel
mkdir -p gpurun_out
export PWMSCAN_PREFILTER=kmer
for nt in 1 2 3; do
  PWMSCAN_FORCE_NT=$nt PWMSCAN_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nt$nt.log 2> gpurun_out/bench_nt$nt.err || { echo BENCH FAILED; tail -20 gpurun_out/bench_nt$nt.err; }
  grep "kmer block" gpurun_out/bench_nt$nt.err | sort -u | head -3
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_nt$nt.log').read().strip().splitlines()[-1])
print('nt=$nt','value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'prefilter ms',round(d['roofline']['ms_per_launch'],3), 'hits', d['config']['hits_per_step_rank0'])
"
done
if [ -n "$NCU" ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:km_prefilter -s 3 -c 1 -o gpurun_out/${TAG:-r1d}_km_prefilter -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
fi
