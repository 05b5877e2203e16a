"""Seeded synthetic ancient-DNA workload.

No network, no datasets: the reference genome stand-in is R-rand (i.i.d. uniform
ACGT, seed 1) and reads are sampled from a diverged copy of it with 5' C->T /
3' G->A deamination, p = 0.3 * 0.5**distance, plus 0.2 % uniform sequencing
error.  Everything is numpy-vectorised so that the 1 M-read BASELINE config is
generated in a couple of seconds.
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def random_reference(length=16569, seed=1):
    """R-rand: uniform ACGT, deterministic."""
    rng = np.random.default_rng(seed)
    return _ACGT[rng.integers(0, 4, length)].tobytes().decode()


def diverge(ref, divergence=0.005, seed=2, indel_rate=0.0):
    """Sample genome: substitutions at `divergence`, optional 1-3 bp indels."""
    rng = np.random.default_rng(seed)
    g = np.frombuffer(ref.encode(), np.uint8).copy()
    hit = rng.random(len(g)) < divergence
    g[hit] = _ACGT[(np.searchsorted(_ACGT, g[hit]) + rng.integers(1, 4, hit.sum())) % 4]
    if indel_rate > 0:
        out, i = [], 0
        pos = np.flatnonzero(rng.random(len(g)) < indel_rate)
        for p in pos:
            out.append(g[i:p])
            n = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                out.append(_ACGT[rng.integers(0, 4, n)])
                i = p
            else:
                i = p + n
        out.append(g[i:])
        g = np.concatenate(out)
    return g.tobytes().decode()


def make_reads(genome, n_reads, min_len=35, max_len=75, seed=2, circular=True,
               damage=0.3, error=0.002, n_rate=0.0):
    """Returns (bases uint8[sum L], offsets int64[n+1], truth dict).

    bases are ASCII upper-case, reads concatenated in order; truth carries the
    sampled start / strand / length for debugging (never used by the product).
    """
    rng = np.random.default_rng(seed)
    g = np.frombuffer(genome.encode(), np.uint8)
    G = len(g)
    L = rng.integers(min_len, max_len + 1, n_reads)
    if circular:
        start = rng.integers(0, G, n_reads)
    else:
        start = (rng.random(n_reads) * (G - L + 1)).astype(np.int64)
    strand = rng.integers(0, 2, n_reads).astype(np.uint8)
    off = np.zeros(n_reads + 1, np.int64)
    np.cumsum(L, out=off[1:])
    total = int(off[-1])
    rid = np.repeat(np.arange(n_reads), L)               # read index of every base
    pos = np.arange(total) - off[rid]                    # 0-based distance from the read's 5' end
    rl = L[rid]
    rc = strand[rid].astype(bool)
    # forward reads take genome[start+pos]; reverse reads take comp(genome[start+L-1-pos])
    gpos = np.where(rc, start[rid] + rl - 1 - pos, start[rid] + pos) % G
    b = g[gpos]
    b = np.where(rc, _COMP[b], b)
    # deamination in read orientation
    u = rng.random(total)
    ct = (b == ord("C")) & (u < damage * 0.5 ** np.minimum(pos, 60))
    ga = (b == ord("G")) & (u < damage * 0.5 ** np.minimum(rl - 1 - pos, 60))
    b = np.where(ct, ord("T"), b)
    b = np.where(ga & ~ct, ord("A"), b)
    # uniform sequencing error
    e = rng.random(total) < error
    b = np.where(e, _ACGT[rng.integers(0, 4, total)], b)
    if n_rate > 0:
        b = np.where(rng.random(total) < n_rate, ord("N"), b)
    return np.ascontiguousarray(b, np.uint8), off, dict(start=start, strand=strand, length=L)


def read_str(bases, off, i):
    return bases[off[i]:off[i + 1]].tobytes().decode()


def revcomp_bytes(a):
    return _COMP[a[::-1]]


def matrix_text(sm):
    """the layout of the reference's matrix files (matrices/*.txt; read_pssm io.c:408-503) for an int[31][5][5]"""
    sm = np.asarray(sm).reshape(31, 5, 5)
    out = []
    for d in range(31):
        name = "MIDDLE" if d == 15 else (str(d + 1) if d < 15 else str(d - 31))
        out.append(f"# Matrix for position: {name}\n")
        for r in range(4):
            out.append("\t".join(str(int(x)) for x in sm[d, r, :4]) + "\t\n")
        out.append("\n")
    return "".join(out)


def fastq_text(bases, off, lo=0, hi=None, qual="I"):
    """reads lo..hi as FASTQ text (ids r0000000...), constant quality"""
    hi = len(off) - 1 if hi is None else hi
    b = bases.tobytes()
    parts = []
    for i in range(lo, hi):
        s = b[off[i]:off[i + 1]]
        parts.append(b"@r%07d\n%s\n+\n%s\n" % (i, s, qual.encode() * len(s)))
    return b"".join(parts)
