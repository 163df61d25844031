#!/bin/bash
# round 2: priority of the two copy streams and the end-to-end arm
mkdir -p gpurun_out
for c in 1 0; do
PWMSCAN_COPY_PRIO=$c BENCH_TRACE_RANK0=1 timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r2v_p$c.log 2> gpurun_out/bench_r2v_p$c.err; echo "copy prio=$c rc=$?"
tail -1 gpurun_out/bench_r2v_p$c.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N',d['n_gpus'],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],3),'e2e ms',round(d['e2e']['ms_per_step'],3), d['phases']['per_step_rank0_ms'], d['config']['items_rank0'])"
done
grep trace gpurun_out/bench_r2v_p1.err | tail -61 | sed -n '20,32p' | cut -c8-240
