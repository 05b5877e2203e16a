"""-m gpu: pass 1 (k-mer filter + whole-reference both-strand DP, chunked kernel) vs the oracle."""
import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


def _reads(n, ref_len, seed, **kw):
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(ref_len, seed=seed)
    g = synth.diverge(ref, kw.pop("divergence", 0.02), seed=seed + 1, indel_rate=kw.pop("indel_rate", 0.004))
    b, off, _ = synth.make_reads(g, n, kw.pop("min_len", 35), kw.pop("max_len", 75), seed=seed + 2, n_rate=0.002)
    reads = [synth.read_str(b, off, i) for i in range(n)]
    rng = np.random.default_rng(seed)
    for i in range(0, n, 17):       # unrelated reads: must be skipped by the filter / score below the cutoff
        reads[i] = "".join("ACGT"[x] for x in rng.integers(0, 4, 50))
    return ref, reads


@pytest.mark.parametrize("k", [0, 8, 12])
def test_pass1_circular(gpu, oracle, k):
    ref, reads = _reads(600, 3000, seed=101)
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), 1, k)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    if k == 12:
        assert (out["status"] & 2).any()       # unrelated reads share no 12-mer with a 3 kb reference


def test_pass1_linear_low_complexity_softmask(gpu, oracle):
    # poly-AC stretch (k-mer position cap + saturation) and a lower-case region with -M
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(1200, seed=5) + "AC" * 300 + synth.random_reference(300, seed=6).lower() + synth.random_reference(600, seed=7)
    rng = np.random.default_rng(9)
    reads = []
    for _ in range(300):
        p = int(rng.integers(0, len(ref) - 80))
        reads.append(ref[p:p + int(rng.integers(25, 75))].upper())
    bad, _ = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("ancient"), 0, 9, soft_mask=1)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_pass1_long_reads(gpu, oracle):
    ref, reads = _reads(200, 2500, seed=111, min_len=100, max_len=256, divergence=0.04, indel_rate=0.01)
    bad, _ = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("pe"), 1, 10)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
