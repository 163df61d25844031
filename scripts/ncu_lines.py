"""Per-source-line summary of an ncu report's source page (instructions, divergence, stall samples).
usage: python scripts/ncu_lines.py report.ncu-rep [min_fraction]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread']
for w in want:
    for i, h in enumerate(hdr):
        if h == w: print(f"{w:70s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
hd = rows[h]
iL, iS, iI, iT = hd.index('Line No'), hd.index('# Samples'), hd.index('Instructions Executed'), hd.index('Thread Instructions Executed')
iW = hd.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hd else None
agg = collections.defaultdict(lambda: [0, 0, 0, 0]); text = {}; tot = [0, 0, 0]
def f(x):
    try: return float(x)
    except ValueError: return 0.0
for r in rows[h + 1:]:
    if len(r) < len(hd): continue
    try: ln = int(r[iL])
    except ValueError: continue
    a = agg[ln]; a[0] += f(r[iS]); a[1] += f(r[iI]); a[2] += f(r[iT]); a[3] += (f(r[iW]) if iW is not None else 0.0); text[ln] = r[1]
    tot[0] += f(r[iS]); tot[1] += f(r[iI]); tot[2] += f(r[iT])
print(f"total: samples {tot[0]:.0f}, warp instructions {tot[1]/1e6:.1f} M, thread instructions {tot[2]/1e9:.2f} G")
for ln, a in sorted(agg.items()):
    if a[1] > tot[1] * thr or a[0] > tot[0] * thr:
        print(f"{ln:4d} samples {100*a[0]/tot[0]:5.1f}%  instr {a[1]/1e6:7.1f} M  thr/inst {a[2]/max(a[1],1):5.1f}  smem wavefronts {a[3]/1e6:6.1f} M | {text[ln][:100]}")
