"""-m gpu: whole assemblies (pass 1 -> iterate to convergence) through the C ABI against what the
UNMODIFIED reference produced for the same inputs (tests/golden/sessions.json): per-read pass-1
results, per-iteration score/as/ae of every read, gaps, and the consensus of every iteration."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(gpu, golden, name, matrix):
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    s = json.load(open(os.path.join(G, "sessions.json")))[name]
    reads = [r for r in s["reads"] if r]
    exp_p1 = [p for p in s["pass1"] if p is not None]
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    A = driver.Assembler(gpu, s["ref"], golden[matrix], s["circular"], s["k"], s["soft_mask"])
    p = A.pass1(bases, off)
    for i, e in enumerate(exp_p1):
        assert int(p["hits"][i]) == e["hits"], (name, i)
        if e["hits"]:
            got = [int(p[k][i]) for k in ("score", "fw_score", "rc_score", "rc", "as_", "ae", "start")]
            assert got == [e[k] for k in ("score", "fw_score", "rc_score", "rc", "as_", "ae", "start")], (name, i)
    for it, e in enumerate(s["iters"]):
        cons, conv = A.iterate()
        got = np.stack([A.score, A.as_, A.ae, A.rc.astype(np.int32)], 1).tolist()
        assert got == e["reads"], f"{name}: iteration {it + 1} per-read score/as/ae"
        assert np.flatnonzero(A.gaps).tolist() == [g for g in e["gaps"] if g < len(A.last)], f"{name}: iteration {it + 1} gaps"
        assert cons == e["cons"], f"{name}: iteration {it + 1} consensus"
        assert conv == e["converged"]
    return A


def test_reference_fixtures_circular(gpu, golden):
    A = _run(gpu, golden, "tr1_tf_c", "ancient")
    assert A.iter == 3 and A.ghosts == 0


def test_reference_fixtures_linear(gpu, golden):
    _run(gpu, golden, "tr1_tf_lin", "ancient")


def test_reference_fixtures_kmer_softmask_stale_back_pointers(gpu, golden):
    _run(gpu, golden, "tr1_tf_c_k8_M", "ancient")


def test_synthetic_circular_kmer(gpu, golden):
    _run(gpu, golden, "synth2k_c_k10", "onepass")


def test_divergent_seed_kmer(gpu, golden):
    # BASELINE configs[3] in small: starting reference 10 % + indels away from the sample, k = 12
    _run(gpu, golden, "synth3k_div10_c_k12", "ancient")


def test_merged_pe_long_reads(gpu, golden):
    # BASELINE configs[2] in small: 30-140 bp reads (16-bit pair kernels up to their frame limit, 32-bit beyond), pe matrix
    _run(gpu, golden, "synth2k5_pe_long_c_k12", "pe")
