"""Multi-GPU sharding of the scan: (motif block x chromosome chunk) work items, no data-path
collective.  Each item is independent: a chunk of window START positions
[start, end) of one sequence (the chunk's bytes carry a halo of max_motif_len - 1 so windows that
start in the chunk and end in the next one are scored by this chunk alone), times a block of motifs.
Ranks take items by longest-processing-time-first; hit lists are concatenated on the host.
"""
import heapq

import numpy as np


def make_items(seq_lengths, n_motifs, chunk_len, motif_block=None):
    """[(seq index, start, end, motif_lo, motif_hi)] covering every (window start, motif) once."""
    motif_block = motif_block or n_motifs
    items = []
    for s, L in enumerate(seq_lengths):
        for st in range(0, max(L, 1), chunk_len):
            en = min(L, st + chunk_len)
            for m0 in range(0, n_motifs, motif_block):
                items.append((s, st, en, m0, min(n_motifs, m0 + motif_block)))
    return items


def pick_chunk_len(total_bp, world_size, items_per_rank=24, lo=8_000_000, hi=64_000_000, step=4_000_000):
    """Chunk length for a whole-genome scan on `world_size` devices: at least ~items_per_rank items per rank, so that the
    fill and drain of a rank's pipeline (first stage of item i+1 over replay / ordering / hit copy of item i) stay a small
    part of the job; never above 64 Mbp (buffer sizes) nor below 8 Mbp (per-launch overheads)."""
    target = total_bp // max(1, items_per_rank * world_size)
    return int(min(hi, max(lo, target // step * step)))


def partition_contiguous(seq_lengths, n_motifs, world_size, max_chunk=64_000_000, align=32, tail_chunk=16_000_000, weights=None):
    """Whole-genome scan on `world_size` devices with as few, as large items as possible: the genome's concatenated coordinate
    space is cut into `world_size` contiguous stretches of equal length (a chunk may start anywhere: every item carries its own
    halo), each stretch into its pieces of single sequences, each piece into equal chunks of at most `max_chunk` window starts
    (cut points on multiples of `align`).  Per-launch costs of the first stage (block changes, the tail of every block) are paid
    per item, so 64 Mbp items cost ~25 % less per base than 16 Mbp ones.  Returns a list (per rank) of items
    (seq index, start, end, 0, n_motifs); a rank's items run largest first and its last one is cut into pieces of at most
    `tail_chunk`, so that what is left when its last first stage has ended — replay, ordering and the copy of ONE item's hits —
    is as short as it can be."""
    total = int(sum(seq_lengths))
    if weights is None:
        bounds = [total * r // world_size for r in range(world_size + 1)]
    else:
        # unequal shares (rank_weights): the devices of a box need not be equally fast end to end — e.g. half of the GPUs
        # of the 8 x B200 box this was measured on move their hit lists into host memory at 60 % of the others' rate
        w = np.asarray(weights, dtype=np.float64)
        if w.size != world_size or not np.all(w > 0):
            raise ValueError("one positive weight per rank expected")
        cum = np.concatenate([[0.0], np.cumsum(w)]) / float(np.sum(w))
        bounds = [int(round(total * c)) for c in cum]
        bounds[0], bounds[-1] = 0, total
    out = [[] for _ in range(world_size)]
    base = 0
    for s, L in enumerate(seq_lengths):
        for r in range(world_size):
            lo, hi = max(bounds[r], base), min(bounds[r + 1], base + L)
            if hi <= lo:
                continue
            lo, hi = lo - base, hi - base
            # cut points inside a sequence go to multiples of `align` (the next rank starts exactly where this one stops)
            if lo > 0:
                lo = min(L, (lo + align - 1) // align * align)
            if hi < L:
                hi = min(L, (hi + align - 1) // align * align)
            if hi <= lo:
                continue
            k = max(1, -(-(hi - lo) // max_chunk))
            cuts = [lo + ((hi - lo) * i // k) // align * align for i in range(k)] + [hi]
            for a, b in zip(cuts[:-1], cuts[1:]):
                if b > a:
                    out[r].append((s, a, b, 0, n_motifs))
        base += L
    for r in range(world_size):
        out[r].sort(key=lambda it: (-(it[2] - it[1]), it[0], it[1]))
        # the rank's last item goes out in pieces of at most `tail_chunk`: what cannot overlap anything is the replay, ordering
        # and hit copy of the LAST piece only
        if tail_chunk and out[r] and out[r][-1][2] - out[r][-1][1] > tail_chunk:
            s, a, b, m0, m1 = out[r].pop()
            k = -(-(b - a) // tail_chunk)
            cuts = [a + ((b - a) * i // k) // align * align for i in range(k)] + [b]
            out[r] += [(s, x, y, m0, m1) for x, y in zip(cuts[:-1], cuts[1:]) if y > x]
    return out


def rank_weights(step_seconds, shares=None, damp=1.0):
    """Shares for `partition_contiguous(weights=...)` from a measured pass: rank r needed step_seconds[r] for its share
    shares[r] (equal shares if None); the new share is proportional to its rate share / time (damp < 1 moves only part of
    the way).  Ranks whose times are within a few per cent of each other keep equal shares."""
    t = np.asarray(step_seconds, dtype=np.float64)
    s = np.ones_like(t) if shares is None else np.asarray(shares, dtype=np.float64)
    rate = s / t
    w = rate / rate.sum()
    w0 = s / s.sum()
    w = w0 + damp * (w - w0)
    return w / w.sum()


def item_cost(item, motif_cols=None):
    s, st, en, m0, m1 = item
    w = (m1 - m0) if motif_cols is None else float(np.sum(motif_cols[m0:m1]))
    return (en - st) * w


def assign(items, world_size, motif_cols=None):
    """LPT assignment; returns a list (per rank) of item lists, each in (seq, start, motif) order."""
    heap = [(0.0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out = [[] for _ in range(world_size)]
    order = sorted(range(len(items)), key=lambda i: -item_cost(items[i], motif_cols))
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(items[i])
        heapq.heappush(heap, (load + item_cost(items[i], motif_cols), r))
    for r in range(world_size):
        out[r].sort()
    return out


def chunk_bytes_range(item, seq_length, max_motif_len):
    """Byte range [lo, hi) of the sequence a rank must hold for `item` (chunk + halo)."""
    _, st, en, _, _ = item
    return st, min(seq_length, en + max_motif_len - 1)


def keep_hit_mask(pos, motif_len, item, seq_length):
    """Hits of a chunk scan that belong to the item: window start (chunk coordinates) inside
    [0, end-start) — windows starting in the halo belong to the next chunk."""
    _, st, en, _, _ = item
    return pos < (en - st)


def merge_hits(per_item):
    """Concatenate per-item hit columns [(seq, motif, strand, pos, score)] and sort them into the
    reference's order (sequence, motif, strand, position)."""
    if not per_item:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z, z, np.zeros(0)
    cols = [np.concatenate([np.asarray(h[k]) for h in per_item]) for k in range(5)]
    order = np.lexsort((cols[3], cols[2], cols[1], cols[0]))
    return tuple(c[order] for c in cols)


class ScanWorkers:
    """Several independent scans in flight on ONE device.

    A rank that owns many (motif block x chromosome chunk) items does not have to run them one
    after the other: every worker is a library context of its own (its own stream, scratch
    buffers and copy of the plan) driven by a host thread (the C calls release the GIL), so the
    tail of one item — exact replay, sort, decode, hit-list D2H — overlaps the prefilter kernel of
    the next item.  Same results as scanning the items sequentially, in the order given.
    (Haskell side: `mapConcurrently` over the items with one `withDevice` context per worker.)
    """

    def __init__(self, device, mats, bg, thres, strands=2, skip_ambiguous=True, workers=2):
        from concurrent.futures import ThreadPoolExecutor
        from . import _lib
        self._lib = _lib
        self.ctxs = [_lib.Context(device) for _ in range(workers)]
        self.motifs = [_lib.DeviceMotifs(c, mats, bg) for c in self.ctxs]
        self.plans = [_lib.Plan(m, thres, strands=strands, skip_ambiguous=skip_ambiguous) for m in self.motifs]
        self.pool = ThreadPoolExecutor(workers)

    def __len__(self):
        return len(self.ctxs)

    def upload(self, arrays):
        """Packs every array into HBM; item i lives with worker i mod k."""
        return [self._lib.DeviceSeq(self.ctxs[i % len(self.ctxs)], a) for i, a in enumerate(arrays)]

    def _run(self, items, fn):
        k = len(self.ctxs)
        out = [None] * len(items)
        timing = [dict() for _ in range(k)]

        def work(w):
            for i in range(w, len(items), k):
                out[i] = fn(w, items[i])
                for key, v in self.ctxs[w].timing().items():
                    timing[w][key] = timing[w].get(key, 0.0) + float(v)
        for f in [self.pool.submit(work, w) for w in range(k)]:
            f.result()
        total = {}
        for t in timing:
            for key, v in t.items():
                total[key] = total.get(key, 0.0) + v
        return out, total

    def scan(self, dev_seqs, consume=None):
        """dev_seqs from `upload`.  Returns ([Hits per item], summed device timing).  With `consume`, every
        hit list is handed to consume(hits) in its worker thread and only the return value is kept — the
        pinned hit buffer goes back to the pool at once instead of piling up until the end (a streaming
        consumer, like the BED writer of scan_motif)."""
        c = consume or (lambda h: h)
        return self._run(dev_seqs, lambda w, s: c(self.plans[w].scan(s)))

    def scan_host(self, arrays, consume=None):
        """Host byte arrays (pinned for full PCIe speed) -> ([Hits per item], summed device timing)."""
        c = consume or (lambda h: h)
        return self._run(arrays, lambda w, a: c(self.plans[w].scan_host(a)))

    def close(self):
        self.pool.shutdown()
        for p in self.plans:
            p.free()
        for m in self.motifs:
            m.free()
