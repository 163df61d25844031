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
