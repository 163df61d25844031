"""Synthetic inputs for the BASELINE.json configs.

DNA comes from a counter-based SplitMix64 stream (position i -> mix(seed + (i+1)*GOLDEN)), so any
slice can be generated independently (host numpy here, and chunk by chunk in bench.py).  PWMs are
small and come from numpy's PCG64 Generator with fixed seeds.  The same bytes / doubles are handed
to the oracle and to the CUDA library, so no cross-language RNG parity is needed.
"""
import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)

HG38_CHROM_SIZES = [
    ("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
    ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
    ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
    ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
    ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415), ("chrM", 16569),
]


def splitmix64(seed, start, count):
    """uint64 values z_i for i in [start, start+count) of the counter-based SplitMix64 stream."""
    with np.errstate(over="ignore"):
        idx = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def synth_dna(length, seed, gc=0.5, start=0, n_runs=(), lower_frac=0.0, chunk=1 << 24):
    """ASCII DNA of `length` bases (positions start..start+length of the stream for `seed`).

    gc: GC fraction; n_runs: iterable of (pos, len) blocks (absolute stream positions) set to 'N';
    lower_frac: fraction of positions soft-masked (lower case), taken from other bits of the stream.
    """
    out = np.empty(length, dtype=np.uint8)
    # thresholds on the top 32 bits: A | C | G | T with P(C)=P(G)=gc/2
    t_a = np.uint64(int((1 - gc) / 2 * 2 ** 32))
    t_c = np.uint64(int(((1 - gc) / 2 + gc / 2) * 2 ** 32))
    t_g = np.uint64(int(((1 - gc) / 2 + gc) * 2 ** 32))
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    lt = np.uint64(int(lower_frac * 2 ** 16))
    for s in range(0, length, chunk):
        c = min(chunk, length - s)
        z = splitmix64(seed, start + s, c)
        hi = z >> np.uint64(32)
        code = (hi >= t_a).astype(np.uint8) + (hi >= t_c).astype(np.uint8) + (hi >= t_g).astype(np.uint8)
        seg = lut[code]
        if lower_frac > 0:
            low = (z & np.uint64(0xFFFF)) < lt
            seg = np.where(low, seg | np.uint8(0x20), seg)
        out[s:s + c] = seg
    for (p, ln) in n_runs:
        a, b = max(p - start, 0), min(p + ln - start, length)
        if b > a:
            out[a:b] = ord("N")
    return out


def synth_pwm(n, rng, alpha=0.3, zero_frac=0.125):
    """n x 4 probability matrix with Dirichlet(alpha) columns; ~zero_frac of the columns get their
    smallest cell forced to an exact 0 (rows renormalised) to exercise the pseudocount."""
    m = rng.dirichlet([alpha] * 4, size=n)
    m = np.maximum(m, 1e-6)
    m = m / m.sum(axis=1, keepdims=True)
    for i in range(n):
        if rng.random() < zero_frac:
            j = int(np.argmin(m[i]))
            m[i, j] = 0.0
            m[i] = m[i] / m[i].sum()
    return np.ascontiguousarray(m, dtype=np.float64)


def synth_pwms(count, seed, nmin=8, nmax=20, alpha=0.3, zero_frac=0.125, lengths=None):
    rng = np.random.Generator(np.random.PCG64(seed))
    if lengths is None:
        lengths = rng.integers(nmin, nmax + 1, size=count)
    return [synth_pwm(int(n), rng, alpha, zero_frac) for n in lengths]


def jaspar_like_lengths(count, seed):
    """Lengths 6..22 with a mode at 10-12."""
    rng = np.random.Generator(np.random.PCG64(seed))
    vals = np.arange(6, 23)
    w = np.exp(-0.5 * ((vals - 11.0) / 3.2) ** 2)
    w = w / w.sum()
    return rng.choice(vals, size=count, p=w)


def ctcf_like_pwm(seed=11, n=19):
    """19-column PWM with a strong core flanked by weak columns and a few exact-zero cells (C1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    conc = np.array([2.0, 1.2, 0.6, 0.15, 0.08, 0.3, 0.08, 0.05, 0.05, 0.4, 0.06, 0.05, 0.3, 0.08, 0.1, 0.5, 1.0, 1.5, 2.0])
    conc = np.resize(conc, n)
    m = np.vstack([rng.dirichlet([c] * 4) for c in conc])
    m = np.maximum(m, 1e-6)
    m = m / m.sum(axis=1, keepdims=True)
    for i in (4, 7, 11):
        if i < n:
            j = int(np.argmin(m[i]))
            m[i, j] = 0.0
            m[i] = m[i] / m[i].sum()
    return np.ascontiguousarray(m, dtype=np.float64)


def config_c1():
    dna = synth_dna(10_000_000, 1, gc=0.5, n_runs=[(4_000_000, 10_000)], lower_frac=0.05)
    return dna, [ctcf_like_pwm()]


def chr1_n_runs(length=248956422):
    return [(0, 10_000), (length - 10_000, 10_000), (121_700_000, 18_000_000)]


def config_c2(length=248956422):
    scale = length / 248956422.0
    runs = [(int(p * scale), max(1, int(l * scale))) for (p, l) in chr1_n_runs()]
    dna = synth_dna(length, 2, gc=0.41, n_runs=runs)
    return dna, synth_pwms(100, 12)


def config_c4_motifs(count=800):
    return synth_pwms(count, 13, lengths=jaspar_like_lengths(count, 13))


def config_c5_motifs(count=2000, seed=15):
    """C5: PWMs with n ~ U{6..20}; 5 % exact palindromes and 5 % duplicates / reverse-complement
    duplicates of earlier motifs, to exercise the tie rules of alignmentBy."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = []
    for i in range(count):
        u = rng.random()
        if i > 10 and u < 0.05:
            half = synth_pwm(int(rng.integers(3, 11)), rng, 0.3, 0.1)
            mats.append(np.ascontiguousarray(np.vstack([half, half.reshape(-1)[::-1].reshape(-1, 4)])))
        elif i > 10 and u < 0.10:
            src = mats[int(rng.integers(0, i))]
            mats.append(src.copy() if rng.random() < 0.5 else np.ascontiguousarray(src.reshape(-1)[::-1].reshape(-1, 4)))
        else:
            mats.append(synth_pwm(int(rng.integers(6, 21)), rng, 0.3, 0.1))
    return mats
