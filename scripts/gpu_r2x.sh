#!/bin/bash
# round 2: hit-list copies on the shared D2H stream (default) vs on the lanes' own streams (PWMSCAN_D2H_LANE=1)
mkdir -p gpurun_out
N=${N:-1}
run() {
  if [ $N -eq 1 ]; then
    env $2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2x_$1_n$N.log 2> gpurun_out/bench_r2x_$1_n$N.err
  else
    env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/bench_r2x_$1_n$N.log 2> gpurun_out/bench_r2x_$1_n$N.err
  fi
  echo "$1 rc=$?"
  tail -1 gpurun_out/bench_r2x_$1_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e packed ms',round(d['e2e']['ms_per_step'],3),'e2e ascii ms',round(d['e2e']['from_ascii']['ms_per_step'],3), d['phases']['per_step_rank0_ms'])"
}
run shared "X=1"
run lane "PWMSCAN_D2H_LANE=1"
