#!/bin/bash
# round 2: config 4 on N GPUs with the contiguous partition (<= 64 Mbp items), per-item trace of rank 0
mkdir -p gpurun_out
N=${N:-8}
BENCH_TRACE_RANK0=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r2q_n$N.log 2> gpurun_out/bench_r2q_n$N.err || { echo BENCH FAILED; tail -30 gpurun_out/bench_r2q_n$N.err; exit 1; }
tail -1 gpurun_out/bench_r2q_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
grep trace gpurun_out/bench_r2q_n$N.err | tail -24 | cut -c1-150
