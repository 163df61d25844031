#!/bin/bash
# round 2 DIAGNOSTIC: config 4 on N GPUs with and without the hit-list copies (what bounds the step at 8 GPUs?)
mkdir -p gpurun_out
N=${N:-8}
run() {
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e > gpurun_out/bench_r2y_$1_n$N.log 2> gpurun_out/bench_r2y_$1_n$N.err
  echo "$1 rc=$?"
  tail -1 gpurun_out/bench_r2y_$1_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), d['phases']['ms_per_step_by_rank'], d['phases']['per_step_rank0_ms'])"
}
run copies "X=1"
run nocopies "PWMSCAN_DIAG_NO_D2H=1"
