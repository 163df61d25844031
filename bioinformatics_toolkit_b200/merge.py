"""Bio.Motif.Merge mirror (bioinformatics-toolkit/src/Bio/Motif/Merge.hs).

mergePWM (32-46), mergePWMWeighted (48-71), dilute (74-81), trim (83-90), mergeTree(Weighted)
(92-111), iterativeMerge (113-170), buildTree (173-176), cutTreeBy (178-186).  The all-pairs
alignment matrix and every row refresh of iterativeMerge run on the device (one launch each); the
greedy argmin / bookkeeping stays on the host, in the reference's scan order.
"""
import math

import numpy as np

from .alignment import pair_index
from .motif import PWM, rcPWM, size


def mergePWM(m1, m2, shift):
    a, b, s = (m1._mat, m2._mat, shift) if shift >= 0 else (m2._mat, m1._mat, -shift)
    n1, n2 = a.shape[0], b.shape[0]
    rows, i = [], 0
    while True:
        i2 = i - s
        if i2 < 0 or (i < n1 and i2 >= n2):
            rows.append(a[i].copy())
        elif i < n1 and i2 < n2:
            rows.append((a[i] + b[i2]) / 2)
        elif i >= n1 and i2 < n2:
            rows.append(b[i2].copy())
        else:
            break
        i += 1
    return PWM(None, np.array(rows))


def mergePWMWeighted(m1, m2, shift):
    """m1, m2 = (PWM, weights); Merge.hs:48-71."""
    (p1, w1), (p2, w2), s = (m1, m2, shift) if shift >= 0 else (m2, m1, -shift)
    a, b = p1._mat, p2._mat
    n1, n2 = a.shape[0], b.shape[0]
    rows, ws, i = [], [], 0
    while True:
        i2 = i - s
        if i2 < 0 or (i < n1 and i2 >= n2):
            rows.append(a[i].copy()); ws.append(w1[i])
        elif i < n1 and i2 < n2:
            wx, wy = w1[i], w2[i2]
            rows.append((float(wx) * a[i] + float(wy) * b[i2]) / float(wx + wy)); ws.append(wx + wy)
        elif i >= n1 and i2 < n2:
            rows.append(b[i2].copy()); ws.append(w2[i2])
        else:
            break
        i += 1
    return PWM(None, np.array(rows)), ws


def dilute(pwm_ws):
    pwm, ws = pwm_ws
    n = max(ws)
    rows = []
    for w, r in zip(ws, pwm._mat):
        if w < n:
            d = float(n - w)
            rows.append((float(w) * r + 0.25 * d) / float(n))
        else:
            rows.append(r.copy())
    return PWM(None, np.array(rows))


def _kld(xs, ys):
    acc = 0.0
    for x, y in zip(xs, ys):
        if x == 0:
            continue
        acc = acc + x * (math.log(x) / math.log(2) - math.log(y) / math.log(2))
    return acc


def trim(bg, cutoff, pwm):
    rs = list(pwm._mat)
    f = [_kld(r, bg.freqs) < cutoff for r in rs]
    lo = 0
    while lo < len(rs) and f[lo]:
        lo += 1
    hi = len(rs)
    while hi > lo and f[hi - 1]:
        hi -= 1
    return PWM(None, np.array(rs[lo:hi]).reshape(-1, 4))


class Leaf:
    def __init__(self, a):
        self.a = a


class Branch:
    def __init__(self, size_, dist, left, right):
        self.size, self.dist, self.left, self.right = size_, dist, left, right


def flatten(t):
    return [t.a] if isinstance(t, Leaf) else flatten(t.left) + flatten(t.right)


def mergeTree(align, t):
    if isinstance(t, Leaf):
        return t.a._pwm
    a, b = mergeTree(align, t.left), mergeTree(align, t.right)
    _, (same, i) = align(a, b)
    return mergePWM(a, b, i) if same else mergePWM(a, rcPWM(b), i)


def mergeTreeWeighted(align, t):
    if isinstance(t, Leaf):
        return t.a._pwm, [1] * size(t.a._pwm)
    (a, w1), (b, w2) = mergeTreeWeighted(align, t.left), mergeTreeWeighted(align, t.right)
    _, (same, i) = align(a, b)
    return mergePWMWeighted((a, w1), (b, w2), i) if same else mergePWMWeighted((a, w1), (rcPWM(b), w2[::-1]), i)


def iterativeMerge(align, th, motifs):
    """Merge.hs:113-170.  Returns [([names], PWM, weights)].

    With a device AlignFn the PWMs AND the distance matrix stay resident on the device (`align.resident`): a merge replaces
    one slot, refreshes one row/column and finds the next minimum there; only that minimum comes back.  The host-matrix path
    below (one all-pairs call, then per-row minimum tracking) serves AlignFns without a resident matrix; an arbitrary
    Python `align` callable falls back to per-call alignment."""
    n = len(motifs)
    items = [([m._name], m._pwm, [1] * size(m._pwm)) for m in motifs]
    if n < 2:
        return items
    resident = align.resident([m._pwm for m in motifs]) if hasattr(align, "resident") else None
    if resident is not None and hasattr(resident, "matrix_init"):
        # the matrix, the row refresh and the search for the minimum all stay on the device (pwmscan_alnset_matrix_init /
        # _merge_step); per merge only mergePWMWeighted runs here
        try:
            nxt = resident.matrix_init()                           # initialisation, Merge.hs:161-165 + the first `loop` (149-158)
            while nxt is not None and nxt[2] < th:
                i, j, _, same, p = nxt
                nm1, pwm1, w1 = items[i]
                nm2, pwm2, w2 = items[j]
                pwm_, w_ = mergePWMWeighted((pwm1, w1), (pwm2, w2), p) if same else mergePWMWeighted((pwm1, w1), (rcPWM(pwm2), w2[::-1]), p)
                items[i] = (nm1 + nm2, pwm_, w_)
                items[j] = None
                nxt = resident.merge_step(i, j, pwm_._mat)
        finally:
            resident.free()
        return [it for it in items if it is not None]
    if resident is not None:
        d, sd, pos = resident.pairs()
    else:
        d, sd, pos = align.all_pairs([m._pwm for m in motifs])
    D = np.full((n, n), np.inf)
    S = np.zeros((n, n), dtype=bool)
    P = np.zeros((n, n), dtype=np.int64)
    iu = np.triu_indices(n, 1)                                     # row-major order == packed order
    D[iu], S[iu], P[iu] = d, sd.astype(bool), pos
    alive = np.ones(n, dtype=bool)
    # per-row minimum (first one in the row) of the upper triangle: the global "first minimum in the scan order i
    # ascending, j ascending" (strict <, Merge.hs:149-158) is the first row whose minimum is smallest
    rmin_col = np.argmin(D, axis=1)
    rmin_val = D[np.arange(n), rmin_col]

    def recompute(rows):
        for r in rows:
            c = int(np.argmin(D[r]))
            rmin_col[r], rmin_val[r] = c, D[r, c]

    try:
        while True:
            i = int(np.argmin(rmin_val))
            j = int(rmin_col[i])
            if not (rmin_val[i] < th):
                break
            same, p = bool(S[i, j]), int(P[i, j])
            nm1, pwm1, w1 = items[i]
            nm2, pwm2, w2 = items[j]
            pwm_, w_ = mergePWMWeighted((pwm1, w1), (pwm2, w2), p) if same else mergePWMWeighted((pwm1, w1), (rcPWM(pwm2), w2[::-1]), p)
            D[:, j] = np.inf                                       # Merge.hs:135 (column j; row j becomes unreachable too)
            D[j, :] = np.inf
            items[i] = (nm1 + nm2, pwm_, w_)
            items[j] = None
            alive[j] = False
            rmin_val[j], rmin_col[j] = np.inf, 0
            dirty = set(int(r) for r in np.flatnonzero((rmin_col == j) & np.isfinite(rmin_val)))   # their minimum sat in column j
            dirty.add(i)
            others = np.flatnonzero(alive)
            others = others[others != i]
            if others.size:                                        # row refresh, Merge.hs:138-145, argument order kept
                if resident is not None:
                    resident.update(i, pwm_._mat)
                    a = np.where(i < others, i, others).astype(np.int32)
                    b = np.where(i < others, others, i).astype(np.int32)
                    dd, ss, pp = resident.pairs(a, b)
                else:
                    pw = [pwm_] + [items[q][1] for q in others]
                    k2 = np.arange(others.size)
                    a = np.where(i < others, 0, k2 + 1)
                    b = np.where(i < others, k2 + 1, 0)
                    dd, ss, pp = align.pairs(pw, a, b)
                r = np.minimum(i, others)
                c = np.maximum(i, others)
                D[r, c], S[r, c], P[r, c] = dd, ss.astype(bool), pp
                low = others[others < i]                           # rows q < i: their entry in column i changed
                if low.size:
                    v = D[low, i]
                    was = rmin_col[low] == i
                    better = ~was & ((v < rmin_val[low]) | ((v == rmin_val[low]) & (i < rmin_col[low])))
                    rmin_val[low[better]] = v[better]
                    rmin_col[low[better]] = i
                    dirty.update(int(q) for q in low[was])
            recompute(dirty)
    finally:
        if resident is not None:
            resident.free()
    return [it for it in items if it is not None]


def _upgma(D, leaves):
    """Average linkage on a symmetric distance matrix (inf on the diagonal), closest pair first, ties to the lowest
    indices.  The minimum is tracked per row, so a merge costs O(n) plus the few rows whose nearest neighbour was
    one of the merged clusters."""
    n = D.shape[0]
    nodes = dict(enumerate(leaves))
    sizes = np.ones(n)
    alive = np.ones(n, dtype=bool)
    rmin_col = np.argmin(D, axis=1)
    rmin_val = D[np.arange(n), rmin_col]
    for _ in range(n - 1):
        a = int(np.argmin(rmin_val))
        b = int(rmin_col[a])
        if a > b:
            a, b = b, a
        dist = float(D[a, b])
        na, nb = sizes[a], sizes[b]
        alive[b] = False
        cs = np.flatnonzero(alive)
        cs = cs[cs != a]
        v = (na * D[a, cs] + nb * D[b, cs]) / (na + nb)
        D[a, cs] = v
        D[cs, a] = v
        D[b, :] = np.inf
        D[:, b] = np.inf
        rmin_val[b] = np.inf
        nodes[a] = Branch(int(na + nb), dist, nodes[a], nodes[b])
        sizes[a] = na + nb
        if cs.size:
            gone = (rmin_col[cs] == a) | (rmin_col[cs] == b)                     # their nearest neighbour was merged
            better = ~gone & ((v < rmin_val[cs]) | ((v == rmin_val[cs]) & (a < rmin_col[cs])))
            rmin_val[cs[better]] = v[better]
            rmin_col[cs[better]] = a
            for r in list(cs[gone]) + [a]:
                c = int(np.argmin(D[r]))
                rmin_col[r], rmin_val[r] = c, D[r, c]
        else:
            rmin_val[a] = np.inf
    return nodes[int(np.flatnonzero(alive)[0])]


def buildTree(align, motifs):
    """hclust Average motifs (fst . align) (Merge.hs:173-176).  `clustering`'s hclust is a
    third-party package the reference does not pin; this is plain average linkage (UPGMA) merging
    the closest pair first (ties: lowest indices)."""
    n = len(motifs)
    if n == 1:
        return Leaf(motifs[0])
    d, _, _ = align.all_pairs([m._pwm for m in motifs])
    D = np.full((n, n), np.inf)
    iu = np.triu_indices(n, 1)
    D[iu] = d
    D = np.minimum(D, D.T)
    return _upgma(D, [Leaf(m) for m in motifs])


def cutAt(tree, h):
    if isinstance(tree, Leaf) or tree.dist <= h:
        return [tree]
    return cutAt(tree.left, h) + cutAt(tree.right, h)


def cutTreeBy(start, step, fn, tree):
    x = start
    while True:
        clusters = cutAt(tree, x)
        if fn(clusters) or not (x - step > 0):
            return clusters
        x = x - step
