"""Host logic of buildTree (Merge.hs:173-176): the row-minimum-tracking UPGMA must build exactly the tree of the
plain O(n^3) formulation (closest pair first, ties to the lowest indices, average linkage) — random matrices,
matrices with many exact ties, and infinities."""
import numpy as np

from bioinformatics_toolkit_b200.merge import Branch, Leaf, _upgma


def upgma_plain(D, leaves):
    n = D.shape[0]
    nodes = {i: leaves[i] for i in range(n)}
    sizes = {i: 1 for i in range(n)}
    active = list(range(n))
    while len(active) > 1:
        sub = D[np.ix_(active, active)]
        k = int(np.argmin(sub))
        ai, bi = divmod(k, len(active))
        a, b = active[ai], active[bi]
        if a > b:
            a, b = b, a
        dist = float(D[a, b])
        na, nb = sizes[a], sizes[b]
        for c in active:
            if c != a and c != b:
                v = (na * D[a, c] + nb * D[b, c]) / (na + nb)
                D[a, c] = D[c, a] = v
        D[b, :] = np.inf
        D[:, b] = np.inf
        nodes[a] = Branch(na + nb, dist, nodes[a], nodes[b])
        sizes[a] = na + nb
        active.remove(b)
    return nodes[active[0]]


def ser(t):
    return ("L", t.a) if isinstance(t, Leaf) else ("B", int(t.size), float(t.dist).hex(), ser(t.left), ser(t.right))


def sym(n, rng, levels=None):
    d = rng.random((n, n)) if levels is None else rng.integers(0, levels, (n, n)).astype(float) / levels
    D = np.triu(d, 1)
    D = D + D.T
    np.fill_diagonal(D, np.inf)
    return D


def test_upgma_equals_plain_formulation():
    rng = np.random.default_rng(7)
    for n, levels in ((2, None), (3, None), (17, None), (60, None), (40, 3), (55, 8), (90, None)):
        D = sym(n, rng, levels)
        leaves = [Leaf(i) for i in range(n)]
        assert ser(_upgma(D.copy(), leaves)) == ser(upgma_plain(D.copy(), leaves)), (n, levels)
