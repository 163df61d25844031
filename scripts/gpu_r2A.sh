#!/bin/bash
# round 2: full GPU test-suite, default bench (N=1) as the driver runs it, reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 900 python bench.py > gpurun_out/bench_r2A.log 2> gpurun_out/bench_r2A.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2A.err
grep '^{' gpurun_out/bench_r2A.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), 'e2e', (round(d['e2e']['ms_per_step'],3), round(d['e2e']['from_ascii']['ms_per_step'],3)), d['phases']['per_step_rank0_ms'], d['cpu_baseline'], d['roofline']['frac'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/ref_r2A.log 2>&1; tail -1 gpurun_out/ref_r2A.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
