#!/bin/bash
# round 2: exact replay reading the compact A/C/G/T table (32 B per column); parity; config-4 chunk
mkdir -p gpurun_out
timeout 600 python scripts/probe_c4.py 4 2>&1 | tail -1 | cut -c1-140
timeout 1500 python -m pytest tests/test_gpu_scan.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2I.log 2> gpurun_out/bench_r2I.err; echo rc=$?
grep '^{' gpurun_out/bench_r2I.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), 'e2e', (round(d['e2e']['ms_per_step'],3), round(d['e2e']['from_ascii']['ms_per_step'],3)), d['phases']['per_step_rank0_ms'])"
