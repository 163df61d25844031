#!/bin/bash
# round 2: CUDA_DEVICE_MAX_CONNECTIONS (streams aliased onto 8 hardware queues by default) and the end-to-end arm
mkdir -p gpurun_out
for c in 32 8; do
CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2u_c$c.log 2> gpurun_out/bench_r2u_c$c.err; echo "conn=$c rc=$?"
tail -1 gpurun_out/bench_r2u_c$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
done
