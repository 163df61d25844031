#!/bin/bash
# round 2, first GPU pass: smoke, parity tests, one-chunk probe, lanes sweep, default bench
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -5 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4.log 2>&1; cat gpurun_out/probe_c4.log | tail -5
timeout 900 python scripts/probe_c4_genome.py > gpurun_out/probe_c4_genome.log 2>&1; tail -12 gpurun_out/probe_c4_genome.log
timeout 900 python bench.py --steps 5 --warmup 2 > gpurun_out/bench_r2a.log 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2a.err
tail -1 gpurun_out/bench_r2a.log | cut -c1-1500
