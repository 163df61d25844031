#!/bin/bash
# round 2: config 4 on N GPUs: last item of every rank in pieces <= 16 Mbp; 6/7 lanes (default) vs 10
mkdir -p gpurun_out
N=${N:-8}
run() {
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 $3 > gpurun_out/bench_r2z_$1_n$N.log 2> gpurun_out/bench_r2z_$1_n$N.err
  echo "$1 rc=$?"
  grep '^{' gpurun_out/bench_r2z_$1_n$N.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3), 'e2e', (round(d['e2e']['ms_per_step'],3), round(d['e2e']['from_ascii']['ms_per_step'],3)) if d['e2e'] else None, d['phases']['ms_per_step_by_rank'], d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
}
run l6 "X=1" ""
run l10 "PWMSCAN_LANES=10" "--no-e2e"
