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


def test_to_shortest_formatting():
    """toShortest (double-conversion / ECMAScript rules) as used by toMEME, writeFasta and `merge_motifs dist`."""
    from bioinformatics_toolkit_b200.motif import _shortest as f
    cases = {1.0: "1", 0.5: "0.5", 123.5: "123.5", 0.1: "0.1", 1e-5: "0.00001", 1e-6: "0.000001", 1e-7: "1e-7", 1.5e-7: "1.5e-7",
             1.2345e-10: "1.2345e-10", 1e21: "1e+21", 1e20: "100000000000000000000", 1.5e21: "1.5e+21", 0.0001: "0.0001",
             0.12889355547153183: "0.12889355547153183", 5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e+308",
             100.0: "100", 1e16: "10000000000000000", 0.000123: "0.000123", -2.5: "-2.5", 3.0e-5: "0.00003", 0.0: "0"}
    for x, s in cases.items():
        assert f(x) == s, (x, f(x), s)
    import random
    random.seed(1)
    for _ in range(5000):
        x = random.uniform(-1, 1) * 10 ** random.randint(-12, 25)
        assert float(f(x)) == x
