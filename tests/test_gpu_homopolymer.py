"""GPU: mia -h -- the homopolymer-discounted gap candidates (mia.c:882-905) in pass 1, in the rounds and in whole assemblies,
against the oracle (pinned to the unmodified reference with -h in tests/test_oracle_vs_ref.py) and against the reference's own
main loop where oracle/_ref is present."""
import random

import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("circular,k,matrix", [(0, 0, "ancient"), (1, 10, "onepass"), (0, 12, "pe")])
def test_pass1_homopolymer_discount(gpu, oracle, circular, k, matrix):
    # 1,400 columns: five chunks; long runs across the chunk boundaries at 256 / 512 / 768 (one of them longer than... a chunk edge + 40)
    ref = hp_reference(1400, 5, long_runs_at=((250, 14, "A"), (505, 9, "C"), (760, 30, "T"), (1020, 6, "G")))
    reads, _ = hp_reads(ref, 260, 17, circular=bool(circular))
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm(matrix), circular, k, hp=1)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    # the discount matters on this data: without it a good part of the reads align differently
    bad0, out0 = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm(matrix), circular, k, hp=0)
    assert not bad0
    assert int((out["score"] != out0["score"]).sum()) > 40


def test_rounds_homopolymer_discount(gpu, oracle):
    ref = hp_reference(2600, 9, long_runs_at=((300, 12, "G"), (1290, 20, "A")))
    reads, starts = hp_reads(ref, 500, 23, lo=25, hi=200)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    # stored orientation + rough coordinates, as pass 1 would leave them: find the strand with the oracle's pass 1
    ctx = oracle.ctx_new(ref, 0, gpu_checks.load_pssm("ancient"), with_rc=1, k=0, hp=1)
    stored, rc, as_, ae = [], [], [], []
    for rd in reads[:300]:
        o = oracle.pass1(ctx, rd)
        stored.append("".join(comp[c] for c in reversed(rd)) if o["rc"] else rd)
        rc.append(o["rc"]); as_.append(o["as_"]); ae.append(o["ae"])
    oracle.ctx_free(ctx)
    off = np.zeros(len(stored) + 1, np.int64)
    np.cumsum([len(r) for r in stored], out=off[1:])
    bases = np.frombuffer("".join(stored).encode(), np.uint8)
    rng = np.random.default_rng(3)
    as_ = (np.array(as_) + rng.integers(-6, 7, len(stored))).clip(0).astype(np.int32)        # a changed consensus moves the windows a little
    ae = (np.array(ae) + rng.integers(-6, 7, len(stored))).astype(np.int32)
    bad, out = gpu_checks.check_realign(gpu, oracle, ref, bases, off, np.array(rc, np.uint8), as_, ae, gpu_checks.load_pssm("ancient"), circular=0, hp=1)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    bad0, out0 = gpu_checks.check_realign(gpu, oracle, ref, bases, off, np.array(rc, np.uint8), as_, ae, gpu_checks.load_pssm("ancient"), circular=0, hp=0)
    assert not bad0
    assert int((out["score"] != out0["score"]).sum()) > 40


@pytest.mark.parametrize("circular,k,distant", [(1, 10, 0), (0, 12, 1)])
def test_assembly_homopolymer_discount(gpu, circular, k, distant):
    # a whole assembly with -h (and -h -D) against the CPU checker: the unmodified reference's main loop where oracle/_ref travelled
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = hp_reference(2200, 31)
    sample = synth.diverge(ref, 0.02 if not distant else 0.04, seed=3, indel_rate=0.01 if not distant else 0.003)
    reads, _ = hp_reads(sample, 400, 41, lo=35, hi=110 if not distant else 80, circular=bool(circular))
    par = gpu_checks.assembly_parity(gpu, ref, reads, gpu_checks.load_pssm("onepass" if not distant else "ancient"), circular, k, distant, hp=1)
    assert par["pass1_equal"] and par["rounds_equal"] and par["consensus_equal"] and par["converged_equal"], par


def test_homopolymer_mode_refusals(gpu):
    import _pkg
    _pkg.load()
    from mia_b200 import api
    gpu.set_pssm(gpu_checks.load_pssm("ancient"))
    gpu.set_reference("ACGTRYACGTACGTACGTACGGGTTTACGATCGATCGATCGATTTTAGCGCGATATATCG" * 3, circular=0, with_rc=1)
    gpu.build_kmers(0)
    reads = ["ACGTACGTACGGGTTTACGATCGATCGATCGATTT"]
    off = np.array([0, len(reads[0])], np.int64)
    gpu.upload_reads(np.frombuffer(reads[0].encode(), np.uint8), off)
    gpu.set_homopolymer(1)
    try:
        with pytest.raises(api.MiaGpuError, match="only A C G T N"):
            gpu.pass1()
        with pytest.raises(api.MiaGpuError, match="not built"):
            gpu.trim(np.frombuffer(reads[0].encode(), np.uint8), off, "ACGTTTACG")
    finally:
        gpu.set_homopolymer(0)
