"""GPU: mia -h -- the homopolymer-discounted gap candidates (mia.c:882-905) in pass 1, in the rounds and in whole assemblies,
against the oracle (pinned to the unmodified reference with -h in tests/test_oracle_vs_ref.py) and against the reference's own
main loop where oracle/_ref is present."""
import numpy as np
import pytest

import gpu_checks
from hp_data import hp_reference, hp_reads

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("name", ["hp2k_c_k10_h", "hp2k_lin_k12_hD"])
def test_reference_sessions_with_homopolymer_discount(gpu, golden, name):
    # the committed sessions of the unmodified reference with -h (tests/golden/make_golden_hp.py) through driver.ResidentAssembler:
    # per round every read's score / as / ae / rc / strand_known and the consensus
    import gzip
    import json
    import os
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    s = json.load(gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hp.json.gz"), "rt"))["sessions"][name]
    reads = s["reads"]
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    try:
        A = driver.ResidentAssembler(gpu, s["ref"], golden[s["matrix"]], s["circular"], s["k"], 0, distant_ref=s["distant_ref"], hp=1)
        p = A.pass1(bases, off)
        for i, e in enumerate(s["pass1"]):
            assert int(p["hits"][i]) == e["hits"]
            if e["hits"]:
                assert [int(p[k2][i]) for k2 in ("score", "rc", "as_", "ae")] == [e[k2] for k2 in ("score", "rc", "as_", "ae")], i
        for it, e in enumerate(s["iters"]):
            cons, conv = A.iterate()
            got = np.stack([A.score, A.as_, A.ae, A.rc.astype(np.int32), A.strand_known.astype(np.int32)], 1).tolist()
            assert got == e["reads"], f"{name}: iteration {it + 1} per-read results"
            assert cons == e["cons"], f"{name}: iteration {it + 1} consensus"
            assert conv == e["converged"]
    finally:
        gpu.set_homopolymer(0)


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
