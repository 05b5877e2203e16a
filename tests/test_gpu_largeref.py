"""-m gpu: references too long for shared-memory staging (> 160 KB: BASELINE configs[4], the 1 Mb region) --
the kernels read the reference through L2 instead and the consensus uses global REDs.  200 kb cases are checked
read by read against the oracle; the 1 Mb case through size-independent properties."""
import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu


def test_realign_and_consensus_200kb_vs_oracle(gpu, oracle):
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(2500, 200_000, seed=301, divergence=0.02, indel_rate=0.004)
    sm = gpu_checks.load_pssm("onepass")
    bad, _ = gpu_checks.check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, sm)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    problems, info = gpu_checks.check_consensus(gpu, oracle, ref, bases, off, rc, as_, ae, sm)
    assert not problems, problems
    assert info["n_ins_cols"] > 0


def test_pass1_200kb_kmer_vs_oracle(gpu, oracle):
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(200_000, seed=311)
    g = synth.diverge(ref, 0.02, seed=312, indel_rate=0.004)
    b, off, _ = synth.make_reads(g, 40, 35, 75, seed=313)
    reads = [synth.read_str(b, off, i) for i in range(40)]
    reads[3] = "ACGT" * 12                                   # unrelated: skipped or below the cutoff
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), 1, 12)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"


def test_one_megabase_properties(gpu, monkeypatch):
    import _pkg
    _pkg.load()
    from mia_b200 import api, synth
    L = 1_000_000
    ref = synth.random_reference(L, seed=321)
    genome = synth.diverge(ref, 0.005, seed=322)
    n = 300_000
    b, off, truth = synth.make_reads(genome, n, 35, 75, seed=323, circular=False)
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=0, with_rc=1)
    gpu.build_kmers(14)                                      # 4^14 k-mers vs 2 x 1 M positions: ~0.3 chance hits per read
    gpu.upload_reads(b, off)
    a = gpu.pass1()
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast + general + skipped == n and fast > 0.5 * n, (fast, general, skipped)
    # (1) the windowed fast path and the general chunked kernel agree read by read (sample: the general kernel is slow)
    m = 20_000
    gpu.upload_reads(b[: off[m]], off[: m + 1])
    monkeypatch.setenv("MIAGPU_PASS1_FAST", "0")
    z = gpu.pass1()
    monkeypatch.delenv("MIAGPU_PASS1_FAST")
    for k in ("hits", "score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end", "abr", "n_runs", "status"):
        assert (a[k][:m] == z[k]).all(), (k, int((a[k][:m] != z[k]).sum()))
    # (2) reads land where they were sampled from (substitutions only: no coordinate shift)
    ok = (a["hits"] > 0) & (a["score"] >= 2000)
    assert ok.mean() > 0.995
    assert (a["rc"][ok] == truth["strand"][ok]).mean() > 0.999
    full = ok & (a["abr"] == 0)
    assert (a["as_"][full] == truth["start"][full]).mean() > 0.99
    # (3) one whole round on the accepted reads: consensus = sample genome at covered positions (up to low-coverage calls)
    keep = ok.astype(np.uint8)
    gpu.upload_reads(b, off)
    gpu.pass1()
    gpu.compact_reads(keep, (a["rc"] == 1).astype(np.uint8))
    idx = np.flatnonzero(ok)
    rc, as_, ae = a["rc"][idx].copy(), a["as_"][idx].copy(), a["ae"][idx].copy()
    seq_len = np.diff(off).astype(np.int32)[idx]
    gpu.set_reference(ref, circular=0, with_rc=0)
    gpu.set_alignment_inputs(rc, as_, ae)
    gpu.set_cut_inputs(seq_len)
    d = np.zeros(len(idx), np.uint8)
    cons, fit, gaps = gpu.iterate_resident(dropped=d, want_gaps=True)
    assert len(cons) == L + int(gaps.sum()) - cons.count("-") or len(cons) >= L - 50
    if len(cons) == L:
        c = np.frombuffer(cons.encode(), np.uint8)
        gref = np.frombuffer(genome.encode(), np.uint8)
        called = c != ord("N")
        assert called.mean() > 0.99                          # ~16x coverage
        assert (c[called] == gref[called]).mean() > 0.999
    # the same round through the separate calls (host score cut): identical flags and consensus
    out = gpu.realign(rc, as_, ae)
    s2 = api.score_cut(seq_len, out["score"])
    below = api.cull_flags(seq_len, out["score"], None, 0, 1, s2[0], s2[1])
    assert (below == d).all() and s2 == fit
    cons2, gaps2, _ = gpu.consensus_natural(below, below, 1)
    assert cons2 == cons and (gaps2 == gaps).all()


def test_unmasked_pass1_200kb_vs_oracle(gpu, oracle):
    # no k-mer filter on a reference of 782 chunks: the 16-bit whole-strand sweep (and the general kernel for the gapped
    # winners) against the oracle's full both-strand DP
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(200_000, seed=331)
    g = synth.diverge(ref, 0.03, seed=332, indel_rate=0.01)
    b, off, _ = synth.make_reads(g, 10, 35, 75, seed=333)
    reads = [synth.read_str(b, off, i) for i in range(10)]
    bad, out = gpu_checks.check_pass1(gpu, oracle, ref, reads, gpu_checks.load_pssm("onepass"), 1, 0)
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    fast, general, skipped = gpu.last_pass1_stats()
    assert fast + general == 10 and fast >= 3, (fast, general, skipped)
