#!/bin/bash
# round 2: first stage chained behind the previous item's ordering (default) vs behind its first stage (PWMSCAN_CHAIN=pre)
mkdir -p gpurun_out
N=${N:-1}
run() { # tag, extra env
  if [ $N -eq 1 ]; then
    env $2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2r_$1_n$N.log 2> gpurun_out/bench_r2r_$1_n$N.err
  else
    env $2 BENCH_TRACE_RANK0=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r2r_$1_n$N.log 2> gpurun_out/bench_r2r_$1_n$N.err
  fi
  echo "$1 rc=$?"
  tail -1 gpurun_out/bench_r2r_$1_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
}
run ord "X=1"
[ $N -eq 1 ] && run pre "PWMSCAN_CHAIN=pre"
grep trace gpurun_out/bench_r2r_ord_n$N.err | sed -n "36,50p" | cut -c1-150; grep trace gpurun_out/bench_r2r_ord_n$N.err | tail -8 | cut -c1-150
