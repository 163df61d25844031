#!/bin/bash
# multi-GPU bench (weak scaling C2 per rank) through torchrun, as the driver launches it
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err || { echo BENCH FAILED; tail -30 gpurun_out/bench_n$N.err; exit 1; }
tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'scaling',d['scaling'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/ref_n$N.log 2> gpurun_out/ref_n$N.err; tail -1 gpurun_out/ref_n$N.log | cut -c1-300
