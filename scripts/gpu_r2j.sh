#!/bin/bash
# round 2: timeline of rank 0's items (PWMSCAN_TRACE) on N GPUs
mkdir -p gpurun_out
N=${N:-8}
if [ $N -eq 1 ]; then
BENCH_TRACE_RANK0=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/trace_c4_n$N.log 2> gpurun_out/trace_c4_n$N.err; echo "rc=$?"
else
BENCH_TRACE_RANK0=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/trace_c4_n$N.log 2> gpurun_out/trace_c4_n$N.err; echo "rc=$?"
fi
grep -c trace gpurun_out/trace_c4_n$N.err
tail -1 gpurun_out/trace_c4_n$N.log | cut -c1-400
