#!/bin/bash
# round 2, last run: what the driver runs (full -m gpu suite, default bench, reference arm, smoke) + C3 app bench + C2 bench and its ncu capture
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 900 python bench.py > gpurun_out/bench_r2P.log 2> gpurun_out/bench_r2P.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2P.err
grep '^{' gpurun_out/bench_r2P.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), 'e2e', (round(d['e2e']['ms_per_step'],3), round(d['e2e']['from_ascii']['ms_per_step'],3)), d['phases']['per_step_rank0_ms'], d['cpu_baseline']['value'], d['roofline']['frac'], d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r2P.log 2>&1; tail -1 gpurun_out/ref_r2P.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python scripts/bench_c3.py > gpurun_out/r2P_c3.log 2>&1; tail -2 gpurun_out/r2P_c3.log | cut -c1-600
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2P_c2.log 2>&1
grep '^{' gpurun_out/bench_r2P_c2.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e']['from_ascii']['ms_per_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:km_prefilter_kernel_v2 -s 3 -c 1 -o gpurun_out/r2P_km2_c2 -f python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2P_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
