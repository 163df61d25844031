"""Key metrics of every kernel launch in an ncu report, as one JSON list (what profiles/*_metrics.json hold).
usage: python scripts/ncu_summary.py report.ncu-rep > profiles/<tag>_metrics.json"""
import csv, io, json, subprocess, sys
WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio', 'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio',
        'smsp__average_warp_latency_issue_stalled_lg_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_wait.ratio',
        'smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio', 'smsp__average_warp_latency_issue_stalled_not_selected.ratio']
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "")}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            try: v = float(r[i].replace(",", ""))
            except ValueError: v = r[i]
            d[w] = {"value": v, "unit": units[i]}
    out.append(d)
json.dump(out, sys.stdout, indent=1)
