#!/bin/bash
mkdir -p gpurun_out
N=${N:-1}
if [ "$N" = "1" ]; then
  PWMSCAN_VERBOSE=1 timeout 1700 python bench.py --workload c4 --steps ${STEPS:-3} --warmup 1 > gpurun_out/bench_c4_n1.log 2> gpurun_out/bench_c4_n1.err || { echo BENCH FAILED; tail -30 gpurun_out/bench_c4_n1.err; exit 1; }
else
  timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --workload c4 --gpus $N --steps ${STEPS:-3} --warmup 1 > gpurun_out/bench_c4_n$N.log 2> gpurun_out/bench_c4_n$N.err || { echo BENCH FAILED; tail -30 gpurun_out/bench_c4_n$N.err; exit 1; }
fi
grep pwmscan gpurun_out/bench_c4_n$N.err | sort | uniq | head -20
tail -1 gpurun_out/bench_c4_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',d['e2e'],'ms/step',round(d['ms_per_step'],3))
print(d['roofline']); print(d['plan']); print(d['config']); print(d['cpu_baseline'])"
