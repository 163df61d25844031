#!/bin/bash
# round 2: parity tests, one-chunk probe + its ncu launch list, lanes sweep, default bench
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > gpurun_out/pytest.log 2>&1; rc=$?
tail -5 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4.log 2>&1; tail -3 gpurun_out/probe_c4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG:-r2b}_launches_c4_chunk.csv python scripts/probe_c4.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 python scripts/probe_c4_genome.py ${LANES:-1 2 3 4} > gpurun_out/probe_c4_genome.log 2>&1; tail -9 gpurun_out/probe_c4_genome.log
timeout 900 python bench.py --steps 5 --warmup 2 > gpurun_out/bench_${TAG:-r2b}.log 2> gpurun_out/bench_${TAG:-r2b}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG:-r2b}.err
tail -1 gpurun_out/bench_${TAG:-r2b}.log | cut -c1-1200
