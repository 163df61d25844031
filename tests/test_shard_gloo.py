"""Multi-rank host logic on CPU: world_size-2 gloo job in which every rank scans its shard of
(chromosome chunk x motif block) items with the ORACLE standing in for the device and rank 0
checks that the gathered, merged hit list equals the single-process scan (no hit lost or doubled
at chunk boundaries)."""
import os
import sys

import numpy as np
import pytest

from bioinformatics_toolkit_b200 import shard, synth

BG = (0.25, 0.25, 0.25, 0.25)


def test_items_cover_everything_once():
    items = shard.make_items([1000, 10, 0, 2500], n_motifs=7, chunk_len=400, motif_block=3)
    cover = {}
    for (s, st, en, m0, m1) in items:
        for m in range(m0, m1):
            cover.setdefault((s, m), []).append((st, en))
    for s, L in enumerate([1000, 10, 0, 2500]):
        for m in range(7):
            iv = sorted(cover[(s, m)])
            assert iv[0][0] == 0 and iv[-1][1] == L
            assert all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
    parts = shard.assign(items, 3)
    assert sorted(sum(parts, [])) == sorted(items)
    loads = [sum(shard.item_cost(i) for i in p) for p in parts]
    assert max(loads) <= 1.5 * (sum(loads) / 3) + 400 * 3


def test_contiguous_partition_covers_everything_once_in_equal_shares():
    lengths = [l for _, l in synth.HG38_CHROM_SIZES]
    for world in (1, 2, 3, 8):
        parts = shard.partition_contiguous(lengths, 800, world, max_chunk=64_000_000)
        cover = {}
        for r in parts:
            sizes = [it[2] - it[1] for it in r]
            big = [x for x in sizes if x > 16_000_000]
            assert big == sorted(big, reverse=True) and sizes[:len(big)] == big                           # largest first, the last item in pieces <= 16 Mbp
            for (s, st, en, m0, m1) in r:
                assert (m0, m1) == (0, 800) and 0 < en - st <= 64_000_000 and (st % 32 == 0)
                cover.setdefault(s, []).append((st, en))
        for s, L in enumerate(lengths):
            iv = sorted(cover[s])
            assert iv[0][0] == 0 and iv[-1][1] == L and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
        loads = [sum(it[2] - it[1] for it in r) for r in parts]
        assert max(loads) - min(loads) <= 64 * world                                                       # equal up to the alignment of the cuts
    tiny = shard.partition_contiguous([100, 0, 7], 5, 4, max_chunk=40, tail_chunk=0)
    assert sorted(sum(tiny, [])) == sorted([(0, 0, 32, 0, 5), (0, 32, 64, 0, 5), (0, 64, 96, 0, 5), (0, 96, 100, 0, 5), (2, 0, 7, 0, 5)])


def test_weighted_shares_follow_the_measured_rates():
    lengths = [l for _, l in synth.HG38_CHROM_SIZES]
    w = shard.rank_weights([17.4] * 4 + [12.4] * 4)                 # what the two halves of the 8 x B200 box needed on equal shares
    assert abs(w.sum() - 1) < 1e-12 and abs(w[4] / w[0] - 17.4 / 12.4) < 1e-9
    half = shard.rank_weights([2.0, 1.0], damp=0.5)                  # only half of the way
    assert abs(half[0] - (0.5 + 0.5 * (1 / 3 - 0.5))) < 1e-12
    parts = shard.partition_contiguous(lengths, 800, 8, weights=w)
    loads = np.array([sum(it[2] - it[1] for it in r) for r in parts], dtype=float)
    assert loads.sum() == sum(lengths) and np.allclose(loads / loads.sum(), w, atol=1e-6)
    cover = {}
    for r in parts:
        for (s, st, en, _, _) in r:
            cover.setdefault(s, []).append((st, en))
    for s, L in enumerate(lengths):
        iv = sorted(cover[s])
        assert iv[0][0] == 0 and iv[-1][1] == L and all(a[1] == b[0] for a, b in zip(iv, iv[1:]))
    with pytest.raises(ValueError):
        shard.partition_contiguous(lengths, 800, 8, weights=[1.0] * 7)


def _worker(rank, world, port, q, contiguous=False):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [30000, 12000, 777]
    seqs = [synth.synth_dna(L, 100 + i, n_runs=[(L // 2, 50)]) for i, L in enumerate(lengths)]
    mats = synth.synth_pwms(10, 77, nmin=6, nmax=18)
    thres = [0.6 * O.optimal_score(BG, m) for m in mats]
    nmax = max(m.shape[0] for m in mats)
    items = shard.make_items(lengths, len(mats), chunk_len=7000, motif_block=4)
    mine = shard.assign(items, world, np.array([m.shape[0] for m in mats]))[rank]
    if contiguous:
        mine = shard.partition_contiguous(lengths, len(mats), world, max_chunk=9000)[rank]
    per_item = []
    for it in mine:
        s, st, en, m0, m1 = it
        lo, hi = shard.chunk_bytes_range(it, lengths[s], nmax)
        cnt, h = O.scan_mt(BG, mats[m0:m1], thres[m0:m1], seqs[s][lo:hi], mode=1, cap=1 << 18)
        keep = shard.keep_hit_mask(h[2], None, it, lengths[s])
        per_item.append((np.full(keep.sum(), s), h[0][keep] + m0, h[1][keep], h[2][keep] + st, h[3][keep]))
    gathered = [None] * world
    dist.all_gather_object(gathered, per_item)
    if rank == 0:
        merged = shard.merge_hits(sum(gathered, []))
        ref = []
        for s, L in enumerate(lengths):
            cnt, h = O.scan_mt(BG, mats, thres, seqs[s], mode=1, cap=1 << 18)
            ref.append((np.full(cnt, s), h[0], h[1], h[2], h[3]))
        ref = shard.merge_hits(ref)
        ok = all(np.array_equal(a, b) for a, b in zip(merged, ref)) and len(ref[0]) > 50
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("contiguous", [False, True])
def test_world_size_2_gloo(contiguous):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (2000 if contiguous else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, contiguous)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert ok
