#!/bin/bash
# round 2: config 4 on N GPUs with shares weighted by the calibration pass
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_r2F_n$N.log 2> gpurun_out/bench_r2F_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_r2F_n$N.err | cut -c1-300
grep '^{' gpurun_out/bench_r2F_n$N.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), 'e2e', (round(d['e2e']['ms_per_step'],3), round(d['e2e']['from_ascii']['ms_per_step'],3)) if d['e2e'] else None, d['phases']['ms_per_step_by_rank'], d['config']['rebalance'], d['config']['items_rank0'], d['config']['hits_per_step'])"
