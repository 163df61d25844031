#!/bin/bash
# round 2: full parity suite (default first stage), one-chunk probes + launch list, traced host-mode genome run (first-stage v2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?
tail -3 gpurun_out/pytest.log
[ $rc -eq 0 ] || { echo PYTEST FAILED; tail -60 gpurun_out/pytest.log; }
PWMSCAN_KM_KERNEL=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2e_launches_c4_chunk.csv python scripts/probe_c4.py 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
PWMSCAN_KM_KERNEL=2 timeout 300 python scripts/probe_c4.py > gpurun_out/probe_c4_v2.log 2>&1; tail -1 gpurun_out/probe_c4_v2.log
PWMSCAN_TRACE=1 PWMSCAN_KM_KERNEL=2 timeout 900 python scripts/probe_c4_genome.py 3 > gpurun_out/probe_c4_genome_v2.log 2> gpurun_out/trace_c4_genome.err; grep -v trace gpurun_out/probe_c4_genome_v2.log | tail -2
python scripts/probe_merge.py 2000 2>&1 | tail -3
