#!/bin/bash
# round 2: non-persistent first stage (v3) + priority streams + gather-form cdf_kernel: parity suite, chunk probes per kernel generation, CDFs, bench + trace
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
for k in 2 3; do PWMSCAN_KM_KERNEL=$k timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_k$k.log 2>&1; echo "k$k: $(tail -1 gpurun_out/probe_c4_k$k.log)"; done
PWMSCAN_KM_RANGE=262144 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_k3_r256k.log 2>&1; echo "k3 range 256k: $(tail -1 gpurun_out/probe_c4_k3_r256k.log)"
PWMSCAN_KM_RANGE=524288 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_k3_r512k.log 2>&1; echo "k3 range 512k: $(tail -1 gpurun_out/probe_c4_k3_r512k.log)"
timeout 300 python scripts/probe_cdf.py 800 > gpurun_out/probe_cdf800.log 2>&1; tail -1 gpurun_out/probe_cdf800.log
timeout 300 python scripts/probe_cdf.py 100 > gpurun_out/probe_cdf100.log 2>&1; tail -1 gpurun_out/probe_cdf100.log
BENCH_TRACE_RANK0=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2k.log 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"; grep -v trace gpurun_out/bench_r2k.err | tail -3
tail -1 gpurun_out/bench_r2k.log | cut -c1-1500
PWMSCAN_KM_RANGE=262144 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2k_r256k.log 2> gpurun_out/bench_r2k_r256k.err; echo "bench 256k rc=$?"
tail -1 gpurun_out/bench_r2k_r256k.log | cut -c1-400
timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_k3.log 2> gpurun_out/bench_c2_k3.err; echo "bench c2 rc=$?"
tail -1 gpurun_out/bench_c2_k3.log | cut -c1-400
