#!/usr/bin/env python
"""bench.py — PWM scan throughput (Gbp*motifs/s) on B200, gated by a hit-set parity check against the oracle.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c2|c1] [--impl ours|reference]

Metric (BASELINE.json): PWM scan Gbp*motifs/sec at 1/2/4/8 B200; one unit = one genome position x one motif,
both strands included.  A step = one pass of the scan over the whole workload.

Default workload at every N: BASELINE config 4, the configuration the metric is quoted on — 800 synthetic
JASPAR-sized PWMs over a 3 088 286 401-bp hg38-shaped synthetic genome (25 chromosomes), p < 1e-5, both strands,
cut into (all motif blocks x chromosome chunk of 64 / 64 / 32 / 16 Mbp at 1 / 2 / 4 / 8 GPUs) items that are sharded over the ranks ("strong" scaling:
the job is fixed, N GPUs share it; no data-path collective, hit lists are gathered in the host memory of the
rank that produced them — the decomposition of the reference's findTFBS', Motif/Search.hs:60-84).
--workload c2 (100 PWMs x chr1-shaped 249 Mbp per GPU, weak scaling) and c1 are the smaller configs.

  value : whole-job throughput with the packed genome, the motif tables and the plan already resident in HBM
          (timed per step: first-stage kernels + exact fp64 replay + ordering + hit lists to host memory).
  e2e   : the same through the C ABI from HOST buffers (pwmscan_scan_host_many): per step H2D of the ASCII
          genome from pinned memory + 2-bit packing + scan + hit lists D2H.
  roofline / phases / cpu_baseline : see DESIGN.md §6.
Before anything is timed, one slice of the workload is scanned on the device and by the oracle and the two hit
lists must be bit-identical ("parity_checked"); the run aborts otherwise.
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
DESC = {
    "c1": "C1: 1 CTCF-like 19-bp PWM x 10 Mbp",
    "c2": "C2: 100 synthetic PWMs (8-20 bp) x 248,956,422-bp chr1-shaped sequence per GPU, p<1e-5, both strands",
    "c4": "C4: 800 synthetic PWMs (6-22 bp) x 3,088,286,401-bp hg38-shaped genome, p<1e-5, both strands, "
          "(all motif blocks x chromosome chunk) items sharded over the ranks",
}


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


def fixture_thresholds(key, count):
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "config_thresholds.json")))
    return np.array([float.fromhex(v) for v in d[key][:count]])


def motifs_of(name):
    from bioinformatics_toolkit_b200 import synth
    if name == "c1":
        return [synth.ctcf_like_pwm()]
    if name == "c2":
        return synth.synth_pwms(100, 12)
    return synth.config_c4_motifs(800)


C4_WEIGHTS = None        # set by main() after the calibration pass (multi-GPU runs)


def c4_items(world):
    """The (all motif blocks x chromosome chunk of <= 64 Mbp) items of config 4 and their assignment to the ranks: equal shares of
    the genome's concatenated coordinates (shard.partition_contiguous)."""
    from bioinformatics_toolkit_b200 import shard, synth
    lengths = [l for _, l in synth.HG38_CHROM_SIZES]
    return lengths, shard.partition_contiguous(lengths, 800, world, weights=C4_WEIGHTS if world > 1 else None)


def c4_chunk(it, lengths, nmax):
    from bioinformatics_toolkit_b200 import shard, synth
    s, st, en, _, _ = it
    lo, hi = shard.chunk_bytes_range(it, lengths[s], nmax)
    runs = [(0, min(10_000, lengths[s])), (max(0, lengths[s] - 10_000), 10_000), (int(lengths[s] * 0.45), int(lengths[s] * 0.03))]
    return "%s:%d-%d" % (synth.HG38_CHROM_SIZES[s][0], st, en), synth.synth_dna(hi - lo, 4 + 100 * s, gc=0.41, start=lo, n_runs=runs)


def workload(name, rank, world):
    """Returns (list of (label, ascii uint8 array, window starts owned) this rank scans, mats)."""
    from bioinformatics_toolkit_b200 import synth
    mats = motifs_of(name)
    if name == "c1":
        dna, _ = synth.config_c1()
        a = np.frombuffer(bytes(dna).upper(), dtype=np.uint8).copy()
        return [("c1", a, a.size)], mats
    if name == "c2":
        L = 248956422
        dna = synth.synth_dna(L, 2 + 1000 * rank, gc=0.41, n_runs=synth.chr1_n_runs(L))
        return [("chr1-shaped/rank%d" % rank, dna, L)], mats
    nmax = max(m.shape[0] for m in mats)
    lengths, per_rank = c4_items(world)
    out = []
    for it in per_rank[rank]:
        label, a = c4_chunk(it, lengths, nmax)
        out.append((label, a, it[2] - it[1]))
    return out, mats


def reference_sample(name, sample):
    """The bounded sample of the workload the CPU arms scan: the first bases of the workload's first sequence."""
    from bioinformatics_toolkit_b200 import synth
    if name == "c4":
        lengths, _ = c4_items(1)
        L = lengths[0]
        return synth.synth_dna(sample, 4, gc=0.41, start=20_000_000, n_runs=[(int(L * 0.45), int(L * 0.03))]), "chr1:20,000,000-"
    if name == "c2":
        return synth.synth_dna(sample, 2, gc=0.41, start=20_000_000), "chr1-shaped:20,000,000-"
    dna, _ = synth.config_c1()
    return np.frombuffer(bytes(dna).upper(), dtype=np.uint8)[:sample].copy(), "c1:0-"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port: look-ahead scan with libm log per visited
    column exactly like Search.hs:185-188; the reference is Haskell and there is no GHC here), all host threads, on a
    bounded sample of the same workload; the throughput is extrapolated from that sample."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    mats = motifs_of(args.workload)
    th = fixture_thresholds(args.workload, len(mats))
    sample = int(args.ref_sample_bp)
    dna, where = reference_sample(args.workload, sample)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        cnt, _ = O.scan_mt(BG, mats, th, dna, mode=0, nthreads=cores, cap=0)
        t1 = time.perf_counter()
        if i >= args.warmup:
            times.append(t1 - t0)
    ms = 1e3 * float(np.mean(times))
    val = dna.size * len(mats) / (ms / 1e3) / 1e9
    desc = "%s%d (%d bp) x %d motifs, both strands; throughput extrapolated from this sample" % (where, 20_000_000 + dna.size, dna.size, len(mats))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.workload == "c4" else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": DESC[args.workload]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                         "note": "oracle port of the reference CPU scan; the genuine reference is single-threaded Haskell"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def parity_gate(ctx, plan, mats, th, seqs_host):
    """One slice of this rank's first sequence through the device and through the oracle: hit lists must be bit-identical."""
    from oracle import oracle as O
    from bioinformatics_toolkit_b200 import _lib
    a = seqs_host[0][1]
    n = min(a.size, 1_000_000 if len(mats) > 200 else 2_000_000)
    lo = max(0, min(a.size - n, a.size // 3))
    sl = np.ascontiguousarray(a[lo:lo + n])
    h = plan.scan(_lib.DeviceSeq(ctx, sl))
    cnt, o = O.scan_mt(BG, mats, th, sl, mode=1, nthreads=os.cpu_count() or 1, cap=max(1 << 22, 8 * len(h)))
    ok = (cnt == len(h) and np.array_equal(h.motif, o[0]) and np.array_equal(h.strand, o[1]) and np.array_equal(h.pos, o[2])
          and np.array_equal(h.score.view(np.uint64), o[3].view(np.uint64)))
    if not ok:
        raise SystemExit("bench.py: PARITY FAILURE — device hit list differs from the oracle on %d bp (device %d hits, oracle %d)" % (n, len(h), cnt))
    return {"parity_checked": True, "parity_slice_bp": int(n), "parity_hits": int(cnt)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c4"])
    ap.add_argument("--ref-sample-bp", type=float, default=None)
    ap.add_argument("--cpu-sample-bp", type=float, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rebalance", action="store_true", help="multi-GPU c4: keep equal shares (no calibration pass)")
    args = ap.parse_args()
    big = args.workload == "c4"
    if args.ref_sample_bp is None:
        args.ref_sample_bp = 1_000_000 if big else 4_000_000
    if args.cpu_sample_bp is None:
        args.cpu_sample_bp = 4_000_000 if big else 2_000_000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if rank == 0 and os.environ.get("BENCH_TRACE_RANK0"):
        os.environ["PWMSCAN_TRACE"] = "1"              # per-item timeline of rank 0 on stderr (pwmscan_scan_many)
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
    t_setup = time.perf_counter()
    seqs_host, mats = workload(args.workload, rank, world)
    mo = _lib.DeviceMotifs(ctx, mats, BG)
    th = mo.pvalue_to_score(1e-5)                    # device scoreCDF + cdf' (Motif.hs:213-214, 252-326)
    plan = _lib.Plan(mo, th)
    M = len(mats)

    def stage(seqs):
        pinned_ = [torch.from_numpy(a).pin_memory() for _, a, _ in seqs]
        arrays_ = [p.numpy() for p in pinned_]
        return pinned_, arrays_, [_lib.DeviceSeq(ctx, a) for a in arrays_]

    pinned, host_arrays, dev_seqs = stage(seqs_host)
    rebalance = None
    if big and world > 1 and not args.no_rebalance:
        # Calibration pass (outside the timed region): the ranks of a box are not equally fast end to end (hit lists into host
        # memory: 17.4 vs 12.4 ms per step on the two halves of the 8 x B200 box), so every rank times a few steps on equal
        # shares and the genome is re-cut in proportion to the measured rates (shard.rank_weights / partition_contiguous).
        global C4_WEIGHTS
        from bioinformatics_toolkit_b200 import shard
        history = []
        shares = [1.0 / world] * world
        for _round in range(2):                                   # two proportional corrections (a rank's time is not quite linear in its share)
            for _ in range(2):
                plan.scan_many(dev_seqs, consume=len)
            dist.barrier()
            t0c = time.perf_counter()
            for _ in range(3):
                plan.scan_many(dev_seqs, consume=len)
            torch.cuda.synchronize()
            mine = torch.tensor([(time.perf_counter() - t0c) / 3], dtype=torch.float64, device="cuda")
            g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(g, mine)
            t_r = [float(x.item()) for x in g]
            history.append([round(1e3 * x, 3) for x in t_r])
            if max(t_r) <= 1.05 * min(t_r):
                break
            shares = [float(x) for x in shard.rank_weights(t_r, shares=shares)]
            C4_WEIGHTS = shares
            del dev_seqs, host_arrays, pinned
            seqs_host, _ = workload(args.workload, rank, world)
            pinned, host_arrays, dev_seqs = stage(seqs_host)
        rebalance = {"calibration_ms_by_rank": history, "weights": C4_WEIGHTS and [round(x, 4) for x in C4_WEIGHTS]}
    units = sum(int(own) for _, _, own in seqs_host) * M          # position x motif units of this rank
    gate = parity_gate(ctx, plan, mats, th, seqs_host) if rank == 0 else {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_setup = time.perf_counter() - t_setup
    acc = {}

    def add_timing():
        for k, v in ctx.timing().items():
            acc[k] = acc.get(k, 0.0) + float(v)

    def step_resident():
        counts = plan.scan_many(dev_seqs, consume=len)        # a hit list is dropped once counted (a streaming consumer)
        add_timing()
        return sum(counts)

    def step_e2e():
        counts = plan.scan_host_many(host_arrays, consume=len)    # pwmscan_scan_host_many: host bytes in, host hit lists out
        add_timing()
        return sum(counts)

    def step_e2e_packed():
        counts = plan.scan_packed_many(packed_items, consume=len)  # pwmscan_scan_packed_many: packed planes of the genome store in, host hit lists out
        add_timing()
        return sum(counts)

    def hit_bytes_rank0():
        return res["d2h_bytes"] / args.steps

    def l2_flush():
        flush.fill_(1)                                         # 256 MiB write on torch's stream ...
        torch.cuda.synchronize()                               # ... finished before the library's own streams start the step

    # ---- device-resident arm -------------------------------------------------------------------
    sampler = ClockSampler(local)          # started before the warm-up: nvidia-smi needs ~100 ms to produce its first row
    sampler.start()
    for _ in range(args.warmup):
        l2_flush()
        nh = step_resident()
    acc.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        l2_flush()
        nh = step_resident()
    torch.cuda.synchronize()
    t_mine = time.perf_counter()                     # this rank's own finish, before the barrier evens the ranks out
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop()
    res = dict(acc)
    el = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
    un = torch.tensor([float(units)], dtype=torch.float64, device="cuda")
    hits_all = torch.tensor([float(nh)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        dist.all_reduce(un, op=dist.ReduceOp.SUM)
        dist.all_reduce(hits_all, op=dist.ReduceOp.SUM)
    per_rank_ms = [1e3 * (t_mine - t0) / args.steps]
    if world > 1:
        g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([per_rank_ms[0]], dtype=torch.float64, device="cuda"))
        per_rank_ms = [float(x.item()) for x in g]
    ms_step = 1e3 * el.item() / args.steps
    value = un.item() / (ms_step / 1e3) / 1e9

    # ---- end-to-end arms (host buffers -> hits on the host) -----------------------------------------
    def run_e2e(step_fn):
        for _ in range(max(1, args.warmup - 1)):
            nh2 = step_fn()
        acc.clear()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nh2 = step_fn()
        barrier()
        t1 = time.perf_counter()
        if nh2 != nh:
            raise SystemExit("bench.py: the end-to-end arm found %d hits, the resident arm %d" % (nh2, nh))
        el2 = torch.tensor([t1 - t0], dtype=torch.float64, device="cuda")
        io = torch.tensor([acc["h2d_bytes"] / args.steps, acc["d2h_bytes"] / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(el2, op=dist.ReduceOp.MAX)
            dist.all_reduce(io, op=dist.ReduceOp.SUM)
        return 1e3 * el2.item() / args.steps, io, acc["h2d_bytes"] / args.steps

    e2e = None
    if not args.no_e2e:
        # (1) the packed genome store: the 2-bit + mask planes of every item, made ONCE on the host (pwmscan_pack_host — what
        #     `mkindex --packed` writes next to the bytes), live in pinned host memory; every step uploads them (0.375 B/base),
        #     scans and brings the hit lists back.  This is the whole-genome path of the API (pwmscan_scan_packed_many).
        t_pack = time.perf_counter()
        packed = [_lib.PackedSeq(a, pinned=True) for a in host_arrays]
        t_pack = time.perf_counter() - t_pack
        packed_items = [(p, 0, None) for p in packed]
        pk_ms, pk_io, pk_h2d = run_e2e(step_e2e_packed)
        # (2) the same from the ASCII bytes (1 B/base over PCIe, packed on the device inside the call: pwmscan_scan_host_many)
        e2e_ms, io, _ = run_e2e(step_e2e)
        # plain pinned->device copy of the same bytes, to read the e2e number against the PCIe link of this box
        big_i = int(np.argmax([p.numel() for p in pinned]))
        dst = torch.empty(pinned[big_i].numel(), dtype=torch.uint8, device="cuda")
        dst.copy_(pinned[big_i], non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dst.copy_(pinned[big_i], non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 3 * pinned[big_i].numel() / (time.perf_counter() - t0) / 1e9
        del dst
        my_h2d = acc["h2d_bytes"] / args.steps
        e2e = {"value": un.item() / (pk_ms / 1e3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(pk_io[0].item()), "d2h_bytes_per_step": int(pk_io[1].item()),
               "ms_per_step": pk_ms,
               "input": "packed genome store in pinned host memory (2-bit + invalid-base planes, 0.375 B/base, made once by pwmscan_pack_host "
                        "in %.2f s on this host, outside the timed region — like mkIndex); call: pwmscan_scan_packed_many" % t_pack,
               "pinned_h2d_gbs_this_box": h2d_gbs,
               "pcie_roofline": {"bound": "pcie (both directions)", "achieved": (pk_h2d + hit_bytes_rank0()) / (pk_ms / 1e3) / 1e9, "peak": h2d_gbs, "unit": "GB/s",
                                 "frac": (pk_h2d + hit_bytes_rank0()) / (pk_ms / 1e3) / 1e9 / h2d_gbs,
                                 "peak_source": "plain cudaMemcpy of a pinned buffer, one direction, measured in this run (rank 0)"},
               # the same from ASCII bytes (1 B/base over PCIe, packed on the device inside the call: pwmscan_scan_host_many)
               "from_ascii": {"value": un.item() / (e2e_ms / 1e3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(io[0].item()),
                              "d2h_bytes_per_step": int(io[1].item()), "h2d_gbs_rank0": my_h2d / (e2e_ms / 1e3) / 1e9}}

    # ---- roofline of the dominant kernel (first stage) and the phases of a step ---------------------------
    peaks = load_peaks()
    info = plan.info()
    n_launch = max(1.0, res["prefilter_launches"])
    pre_ms_launch = res["prefilter_ms"] / n_launch
    bp_rank = sum(int(a.size) for _, a, _ in seqs_host)
    bp_launch = bp_rank * args.steps / n_launch
    clk = peaks["sm_max_mhz"] * 1e6
    hbm_bytes = 0.25 * bp_launch * info["genome_passes"]          # 2-bit genome read once per block pass
    ncol_mean = float(np.mean([m.shape[0] for m in mats]))
    lane_ops = 2.0 * ncol_mean * bp_launch * M                     # brute-force look-ups + adds, both strands
    issue_peak = 148 * 128 * clk / 1e12
    sec = pre_ms_launch / 1e3
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")          # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full captures
    ncu_facts = None
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get("km_prefilter_kernel", {}).get(args.workload)
        ncu_facts = tj.get("km_prefilter_kernel_ncu", {}).get(args.workload)
    if info["mode"] == "kmer":
        # governing roof: the L1 gather rate, one 32-byte sector per clock per SM (experiments/masktab/l2gather.cu; profiles/l2gather_*.txt).
        # Algorithmic bytes: the plan's expected mask sectors per position x 32 B (a model; ncu's l1tex sector count agrees within 1 %).
        alg_bytes = info["table_sectors_per_position"] * 32.0 * bp_launch
        peak = 148 * 32 * clk / 1e9
        roofline = {"bound": "l1-gather", "achieved": alg_bytes / sec / 1e9, "peak": peak, "unit": "GB/s", "frac": alg_bytes / sec / 1e9 / peak,
                    "traffic": traffic, "kernel": "km_prefilter_kernel", "ms_per_launch": pre_ms_launch, "launches_per_step": n_launch / args.steps,
                    "peak_source": "148 SMs x one 32-B sector per clk x sm_max_mhz (%s); this repo's micro-benchmark reaches 8.9 TB/s (profiles/l2gather_*.txt); not a MEASURED_PEAKS.json entry" % peaks["source"],
                    "algorithmic_bytes_per_launch": alg_bytes, "ncu": ncu_facts,
                    "note": "L2-resident k-mer mask tables gathered with 256-bit loads; 'achieved' uses the plan's modelled sectors per position "
                            "(ncu's l1tex sector count agrees within a few per cent); with many motifs (c4: six of seven blocks are a shared-memory bitmap "
                            "scan with few gathers) the kernel is instruction- and latency-bound rather than gather-bound: see 'ncu' (issue slots, L1TEX) "
                            "and profiles/README.md; --workload c2 (one dense block) reaches 0.82 of this roof"}
    else:
        smem_bytes = info.get("smem_wavefronts_per_position", 0.0) * 128.0 * bp_launch
        smem_peak = 148 * 128 * clk / 1e9
        roofline = {"bound": "smem", "achieved": smem_bytes / sec / 1e9, "peak": smem_peak, "unit": "GB/s", "frac": smem_bytes / sec / 1e9 / smem_peak,
                    "traffic": None, "kernel": "prefilter_kernel", "ms_per_launch": pre_ms_launch,
                    "peak_source": "148 SMs x 128 B/clk (B300_MICROARCH.md) x sm_max_mhz (%s)" % peaks["source"],
                    "algorithmic_bytes_per_launch": smem_bytes}
    roofline_hbm = {"bound": "hbm", "achieved": hbm_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_bytes / sec / 1e9 / peaks["hbm_gbs"],
                    "traffic": traffic, "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peaks["source"],
                    "note": "HBM is not the governing roof (0.25 B/base per block pass)"}
    issue = {"bound": "int/fp32 issue of the brute-force algorithm", "achieved": lane_ops / sec / 1e12, "peak": issue_peak,
             "unit": "T lane-ops/s (2*n per bp*motif)", "frac": lane_ops / sec / 1e12 / issue_peak,
             "note": "> 1 means the table-driven first stage does fewer operations than the brute-force count"}
    hit_bytes = res["d2h_bytes"] / args.steps
    phases = {"per_step_rank0_ms": {"first_stage": res["prefilter_ms"] / args.steps, "rescore": res["rescore_ms"] / args.steps,
                                    "order": res["order_ms"] / args.steps},
              "ms_per_step_by_rank": [round(x, 3) for x in per_rank_ms],
              "hit_bytes_to_host_per_step_rank0": hit_bytes, "hit_d2h_gbs_if_serial": hit_bytes / (ms_step / 1e3) / 1e9,
              "note": "CUDA-event time on each item's stream, summed over the items of a step; items overlap (3 in flight), so the sum may exceed ms_per_step"}

    # ---- CPU baseline (oracle port, bounded sample), rank 0 at N=1 only ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        sub, where = reference_sample(args.workload, int(args.cpu_sample_bp))
        t0 = time.perf_counter()
        O.scan_mt(BG, mats, th, sub, mode=0, nthreads=cores, cap=0)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        O.scan_mt(BG, mats, th, sub[: max(1, sub.size // 8)], mode=0, nthreads=1, cap=0)
        dt1 = time.perf_counter() - t0
        cpu = {"value": sub.size * M / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%s (%d bp) x %d motifs, both strands, look-ahead scan with libm log per visited column (Search.hs:185-188); extrapolated" % (where, sub.size, M),
               "one_core_value": (sub.size // 8) * M / dt1 / 1e9}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if big else "weak",
            "vs_baseline": None, "dtype": "u8 bit-mask / u16 first stage + f64 re-score", "data": "synthetic",
            "config": {"workload": DESC[args.workload], "items_rank0": len(dev_seqs), "chunk_bp": int(max(a.size for a in host_arrays)), "scans_in_flight": os.environ.get("PWMSCAN_LANES", "6 resident / 7 from host bytes"),
                       "l2": "256 MiB flush write + sync between steps; per step the rank reads %.0f MB of packed genome and %.0f MB of mask tables (> 126 MB L2)" % (0.375 * bp_rank / 1e6, info["table_bytes"] / 1e6),
                       "thresholds": "device scoreCDF", "hits_per_step": int(hits_all.item()), "setup_s": round(t_setup, 1),
                       "shares": ("equal" if not rebalance or not rebalance["weights"] else "weighted by each rank's measured rate (calibration pass outside the timed region)"),
                       "rebalance": rebalance, **gate},
            "e2e": e2e, "gpu_launches": int(res["kernel_launches"]), "clocks": clocks,
            "roofline": roofline, "roofline_hbm": roofline_hbm, "issue_roofline": issue, "phases": phases, "plan": info, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
