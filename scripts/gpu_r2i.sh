#!/bin/bash
# round 2: default bench (C4, strong scaling) on N GPUs, as the driver launches it
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c4_n$N.log 2> gpurun_out/bench_c4_n$N.err; echo "bench n=$N rc=$?"
tail -1 gpurun_out/bench_c4_n$N.log | cut -c1-3000
tail -5 gpurun_out/bench_c4_n$N.err
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; numactl -H >> gpurun_out/topo_n$N.txt 2>&1; nproc >> gpurun_out/topo_n$N.txt
