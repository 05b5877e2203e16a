"""Homopolymer-rich synthetic data for the mia -h tests (GPU parity tests and the golden generator share it)."""
import random


def hp_reference(n, seed, long_runs_at=()):
    rng = random.Random(seed)
    out = []
    while sum(len(x) for x in out) < n:
        out.append(rng.choice("ACGT") * min(9, max(1, int(rng.expovariate(0.5)))))
    s = list("".join(out)[:n])
    for pos, length, base in long_runs_at:            # homopolymers laid across the chunked kernel's 256-column boundaries
        s[pos:pos + length] = base * length
    return "".join(s[:n])


def hp_reads(ref, n, seed, lo=30, hi=120, circular=False):
    """reads copied from the reference with homopolymer lengths changed here and there, a few substitutions, both strands"""
    rng = random.Random(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    reads, starts = [], []
    for _ in range(n):
        L = rng.randint(lo, hi)
        p = rng.randint(0, len(ref) - 1 if circular else len(ref) - L)
        src = (ref + ref)[p:p + L]
        runs, i = [], 0
        while i < len(src):
            j = i
            while j < len(src) and src[j] == src[i]:
                j += 1
            runs.append((src[i], j - i))
            i = j
        rd = []
        for b, k in runs:
            x = rng.random()
            if x < 0.09:
                k = max(1, k + rng.choice((-2, -1, 1, 1, 2)))
            elif x < 0.11:
                b = rng.choice("ACGT")
            rd.append(b * k)
        rd = "".join(rd)[:250]
        if rng.random() < 0.5:
            rd = "".join(comp[c] for c in reversed(rd))
        reads.append(rd)
        starts.append(p)
    return reads, starts
