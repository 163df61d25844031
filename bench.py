#!/usr/bin/env python
"""bench.py — PWM scan throughput (Gbp*motifs/s) on B200, with hit-set parity as the gate.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c4] [--impl ours|reference]

Metric (BASELINE.json): PWM scan Gbp*motifs/sec; one unit = one genome position x one motif, both
strands included.  A step = one pass of the scan over the rank's synthetic sequence.

Default workload at every N: BASELINE config 2 per GPU — 100 synthetic PWMs (8-20 bp) over a
synthetic 248 956 422-bp chr1-shaped sequence, p < 1e-5, both strands ("weak" scaling: each rank
scans its own chromosome-sized sequence, no data-path collective; hit lists stay on the host of the
rank that produced them).  --workload c4 is the strong-scaling job of config 4 (800 PWMs x 3.1 Gbp
hg38-shaped genome, chromosome chunks sharded over the ranks).

  value : whole-job throughput with the packed genome, the motif tables and the plan already
          resident in HBM (timed: prefilter kernel + fp64 re-score + sort + hit list D2H).
  e2e   : the same through the C ABI from HOST buffers: per step H2D of the ASCII sequence from
          pinned memory + 2-bit packing + scan + hit list D2H.
  roofline / cpu_baseline : see DESIGN.md §6.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BG = (0.25, 0.25, 0.25, 0.25)
METRIC = "PWM scan throughput (genome positions x motifs per second, both strands)"
UNIT = "Gbp*motifs/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def thresholds(key, motifs_dev, count):
    """p = 1e-5 cutoffs from the device scoreCDF; until that kernel reports ready they come from the
    committed known-answer fixture (tests/golden/config_thresholds.json, oracle-generated)."""
    from bioinformatics_toolkit_b200 import _lib
    try:
        return motifs_dev.pvalue_to_score(1e-5), "device scoreCDF"
    except _lib.PwmScanError:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "config_thresholds.json")))
        return np.array([float.fromhex(v) for v in d[key][:count]]), "fixture tests/golden/config_thresholds.json"


def workload(name, rank, world):
    """Returns (list of (label, ascii uint8 array) this rank scans, mats, fixture key, description)."""
    from bioinformatics_toolkit_b200 import shard, synth
    if name == "c1":
        dna, mats = synth.config_c1()
        return [("c1", np.frombuffer(bytes(dna).upper(), dtype=np.uint8).copy())], mats, "c1", "C1: 1 CTCF-like 19-bp PWM x 10 Mbp"
    if name == "c2":
        L = 248956422
        dna = synth.synth_dna(L, 2 + 1000 * rank, gc=0.41, n_runs=synth.chr1_n_runs(L))
        return [("chr1-shaped/rank%d" % rank, dna)], synth.synth_pwms(100, 12), "c2", \
            "C2: 100 synthetic PWMs (8-20 bp) x 248,956,422-bp chr1-shaped sequence per GPU, p<1e-5, both strands"
    if name == "c4":
        mats = synth.config_c4_motifs(800)
        nmax = max(m.shape[0] for m in mats)
        lengths = [l for _, l in synth.HG38_CHROM_SIZES]
        items = shard.make_items(lengths, len(mats), chunk_len=64_000_000)
        mine = shard.assign(items, world)[rank]
        seqs = []
        for it in mine:
            s, st, en, _, _ = it
            lo, hi = shard.chunk_bytes_range(it, lengths[s], nmax)
            runs = [(0, min(10_000, lengths[s])), (max(0, lengths[s] - 10_000), 10_000), (int(lengths[s] * 0.45), int(lengths[s] * 0.03))]
            seqs.append(("%s:%d-%d" % (synth.HG38_CHROM_SIZES[s][0], st, en), synth.synth_dna(hi - lo, 4 + 100 * s, gc=0.41, start=lo, n_runs=runs)))
        return seqs, mats, "c4", "C4: 800 synthetic PWMs x 3,088,286,401-bp hg38-shaped genome, 64 Mbp chromosome chunks sharded over ranks"
    raise SystemExit("unknown workload " + name)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port: look-ahead scan with libm
    log per visited column exactly like Search.hs:185-188), all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import oracle as O
    from bioinformatics_toolkit_b200 import synth
    cores = os.cpu_count() or 1
    mats = synth.synth_pwms(100, 12) if args.workload != "c4" else synth.config_c4_motifs(800)
    key = "c2" if args.workload != "c4" else "c4"
    th = np.array([float.fromhex(v) for v in json.load(open(os.path.join(ROOT, "tests", "golden", "config_thresholds.json")))[key]])
    sample = int(args.ref_sample_bp)
    dna = synth.synth_dna(sample, 2, gc=0.41)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cnt, _ = O.scan_mt(BG, mats, th, dna, mode=0, nthreads=cores, cap=0)
        t1 = time.perf_counter()
        if i >= args.warmup:
            times.append(t1 - t0)
    ms = 1e3 * float(np.mean(times))
    val = sample * len(mats) / (ms / 1e3) / 1e9
    desc = "first %d bp of the workload's sequence shape x %d motifs, both strands" % (sample, len(mats))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "oracle port of the reference CPU scan (reference is Haskell, no GHC here): " + desc},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c4"])
    ap.add_argument("--ref-sample-bp", type=float, default=4_000_000)
    ap.add_argument("--cpu-sample-bp", type=float, default=2_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bioinformatics_toolkit_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = _lib.default_context(local)
    seqs_host, mats, key, desc = workload(args.workload, rank, world)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    th, th_src = thresholds(key, mo, len(mats))
    plan = _lib.Plan(mo, th)
    M = len(mats)
    units = sum(int(a.size) for _, a in seqs_host) * M          # position x motif units of this rank
    pinned = [torch.from_numpy(a).pin_memory() for _, a in seqs_host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    last_timing = {}
    # A rank with many chunks (c4) keeps two scans in flight (shard.ScanWorkers: two library contexts driven by
    # two host threads), so one chunk's exact replay / sort / hit D2H overlaps the next chunk's prefilter kernel.
    workers = None
    if len(pinned) > 1 and not os.environ.get("PWMSCAN_BENCH_SEQUENTIAL"):
        from bioinformatics_toolkit_b200.shard import ScanWorkers
        workers = ScanWorkers(local, mats, BG, th, workers=2)
        dev_seqs = workers.upload([p.numpy() for p in pinned])
    else:
        dev_seqs = [_lib.DeviceSeq(ctx, p.numpy()) for p in pinned]

    def step_resident():
        nh, pre_ms = 0, 0.0
        last_timing.clear()
        if workers is not None:
            counts, t = workers.scan(dev_seqs, consume=len)        # like the sequential loop: a hit list is dropped once counted
            last_timing.update(t)
            return sum(counts), t["prefilter_ms"]
        for s in dev_seqs:
            h = plan.scan(s)
            nh += len(h)
            t = ctx.timing()
            pre_ms += t["prefilter_ms"]
            for k, v in t.items():
                last_timing[k] = last_timing.get(k, 0.0) + float(v)
        return nh, pre_ms

    def step_e2e():
        nh, h2d, d2h = 0, 0, 0
        if workers is not None:
            counts, t = workers.scan_host([p.numpy() for p in pinned], consume=len)
            return sum(counts), sum(p.numel() for p in pinned), int(t["d2h_bytes"])
        for p in pinned:
            h = plan.scan_host(p.numpy())            # pwmscan_scan_host: host bytes in, host hit list out
            h2d += p.numel()
            d2h += int(ctx.timing()["d2h_bytes"])
            nh += len(h)
        return nh, h2d, d2h

    # ---- device-resident arm -------------------------------------------------------------------
    sampler = ClockSampler(local)          # started before the warm-up: nvidia-smi needs ~100 ms to produce its first row
    sampler.start()
    for _ in range(args.warmup):
        flush.fill_(1)
        nh, _ = step_resident()
    pre_ms_tot = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                                         # L2 flush (256 MiB write) between steps
        nh, pre_ms = step_resident()
        pre_ms_tot += pre_ms
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    el = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    un = torch.tensor([float(units)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(un, op=dist.ReduceOp.SUM)
    ms_step = 1e3 * el.item() / args.steps
    value = un.item() / (ms_step / 1e3) / 1e9

    # ---- end-to-end arm (host buffers -> hits on the host) -----------------------------------------
    for _ in range(max(1, args.warmup - 1)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, h2d, d2h = step_e2e()
    barrier()
    t1 = time.perf_counter()
    el2 = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(el2, op=dist.ReduceOp.MAX)
    e2e_val = un.item() / (el2.item() / args.steps) / 1e9

    # ---- plain pinned->device copy of the same bytes, to read the e2e number against the PCIe link ----
    dst = torch.empty(pinned[0].numel(), dtype=torch.uint8, device="cuda")
    dst.copy_(pinned[0], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(pinned[0], non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * pinned[0].numel() / (time.perf_counter() - t0) / 1e9
    del dst

    # ---- roofline of the dominant kernel (prefilter) ------------------------------------------------
    peaks = load_peaks()
    info = plan.info()
    pre_ms_launch = pre_ms_tot / args.steps / max(1, len(dev_seqs))
    bp_launch = sum(int(a.size) for _, a in seqs_host) / max(1, len(dev_seqs))
    clk = peaks["sm_max_mhz"] * 1e6
    hbm_bytes = 0.25 * bp_launch * info["genome_passes"]          # 2-bit genome read once per block pass
    hbm_ach = hbm_bytes / (pre_ms_launch / 1e3) / 1e9
    ncol_mean = float(np.mean([m.shape[0] for m in mats]))
    lane_ops = 2.0 * ncol_mean * bp_launch * M                     # brute-force look-ups + adds, both strands
    issue_peak = 148 * 128 * clk / 1e12
    traffic = None
    if info["mode"] == "kmer":
        # governing roof: the L1 gather rate, one 32-byte sector per clock per SM (experiments/masktab/l2gather.cu
        # reaches 96 % of it: 8.9 TB/s).  Algorithmic bytes: the plan's expected mask sectors per position x 32 B.
        alg_bytes = info["table_sectors_per_position"] * 32.0 * bp_launch
        peak = 148 * 32 * clk / 1e9
        ach = alg_bytes / (pre_ms_launch / 1e3) / 1e9
        # DRAM traffic of one launch on this exact workload, from the committed ncu --set full capture
        traffic = 761.9e6 if (args.workload == "c2" and M == 100) else None   # profiles/r1j: 752.8 MB read (64 MiB of tables + L2 misses) + 9.1 MB written
        roofline = {"bound": "l1-gather", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "kernel": "km_prefilter_kernel", "ms_per_launch": pre_ms_launch,
                    "peak_source": "148 SMs x one 32-B sector per clk x sm_max_mhz (%s); measured 8.9 TB/s by experiments/masktab/l2gather.cu" % peaks["source"],
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "note": "L2-resident k-mer mask tables gathered with 256-bit loads; the L1 sector rate, not HBM or L2 bytes, is the roof"}
    else:
        # governing roof: shared-memory crossbar, one 128-B wavefront per clock per SM (B300_MICROARCH.md; not a
        # MEASURED_PEAKS.json entry).  Algorithmic bytes: first-stage look-up rows, S1_b x 128 B per position per block.
        smem_bytes = info.get("smem_wavefronts_per_position", 0.0) * 128.0 * bp_launch
        smem_peak = 148 * 128 * clk / 1e9
        smem_ach = smem_bytes / (pre_ms_launch / 1e3) / 1e9
        # (profiles/r1b_prefilter_ncu_full_metrics.json: dram__bytes_read.sum 62.9 MB + dram__bytes_write.sum 3.1 MB)
        traffic = 66.0e6 if (args.workload == "c2" and M == 100 and info["mode"] == "lds") else None
        roofline = {"bound": "smem", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s", "frac": smem_ach / smem_peak,
                    "traffic": traffic, "kernel": "prefilter_kernel", "ms_per_launch": pre_ms_launch,
                    "peak_source": "148 SMs x 128 B/clk (B300_MICROARCH.md) x sm_max_mhz (%s)" % peaks["source"],
                    "algorithmic_bytes_per_launch": smem_bytes,
                    "note": "shared-memory-bandwidth bound look-up kernel; ncu (profiles/) reports the LSU/shared pipe at the same fraction"}
    roofline_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"], "traffic": traffic,
                    "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peaks["source"],
                    "note": "HBM is not the governing roof (0.25 B/base per block pass)"}
    issue = {"bound": "int/fp32 issue of the brute-force algorithm", "achieved": lane_ops / (pre_ms_launch / 1e3) / 1e12,
             "peak": issue_peak, "unit": "T lane-ops/s (2*n per bp*motif)",
             "frac": lane_ops / (pre_ms_launch / 1e3) / 1e12 / issue_peak,
             "note": "> 1 means the table-driven prefilter does fewer operations than the brute-force count"}

    # ---- CPU baseline (oracle port, bounded sample), rank 0 at N=1 only ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sample = int(args.cpu_sample_bp)
        sub = seqs_host[0][1][5_000_000:5_000_000 + sample] if seqs_host[0][1].size > 5_000_000 + sample else seqs_host[0][1][:sample]
        t0 = time.perf_counter()
        cnt, _ = O.scan_mt(BG, mats, th, sub, mode=0, nthreads=cores, cap=0)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.scan_mt(BG, mats, th, sub[: max(1, sub.size // 4)], mode=0, nthreads=1, cap=0)
        dt1 = time.perf_counter() - t0
        cpu = {"value": sub.size * M / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d bp x %d motifs, both strands, look-ahead scan with libm log per visited column (Search.hs:185-188)" % (sub.size, M),
               "one_core_value": (sub.size // 4) * M / dt1 / 1e9}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak",
            "vs_baseline": None, "dtype": "bit-mask / u16 prefilter + f64 re-score", "data": "synthetic",
            "config": {"workload": desc, "l2": "256 MiB flush write between steps", "thresholds": th_src,
                       "scans_in_flight": 2 if workers is not None else 1,
                       "hits_per_step_rank0": int(nh),
                       "device_ms_breakdown_last_step": {k: round(float(v), 4) for k, v in last_timing.items()}},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * el2.item() / args.steps, "pinned_h2d_gbs_this_box": h2d_gbs,
                    # the end-to-end path is bound by the host->device copy of the ASCII bytes (1 B per base): this rank's share
                    "pcie_roofline": {"bound": "pcie-h2d", "achieved": h2d / (el2.item() / args.steps) / 1e9, "peak": h2d_gbs, "unit": "GB/s",
                                      "frac": h2d / (el2.item() / args.steps) / 1e9 / h2d_gbs,
                                      "peak_source": "plain cudaMemcpy of the same pinned buffer, measured in this run"}},
            "gpu_launches": int(args.steps * len(dev_seqs) * 3), "clocks": clocks,   # km_prefilter + rescore + decode per scan (the radix sort is CUB)
            "roofline": roofline, "roofline_hbm": roofline_hbm, "issue_roofline": issue, "plan": info, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
