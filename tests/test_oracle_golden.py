"""CPU: the oracle restatement against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  These run everywhere, including the GPU box where
/root/reference does not exist."""
import json
import os

import numpy as np
import pytest

from oracle_driver import OracleRun

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pssm_parser_flat_and_revcom(oracle, golden):
    text = json.load(open(os.path.join(G, "onepass_matrix_fixture.json")))["text"]
    assert (oracle.parse_pssm(text) == golden["onepass"]).all()
    assert (oracle.flat_pssm() == golden["flat"]).all()
    for k in ("ancient", "onepass", "pe", "flat"):
        assert (oracle.revcom_pssm(golden[k]) == golden[k + "_rc"]).all()
        assert (oracle.revcom_pssm(golden[k + "_rc"]) == golden[k]).all()       # involution


def test_sm_depth_rule(oracle):
    # pssm.c:36-46: 5' depth wins over 3' depth for short reads
    assert [oracle.sm_depth(r, 40) for r in (0, 14, 15, 24, 25, 26, 39)] == [0, 14, 15, 15, 16, 17, 30]
    assert [oracle.sm_depth(r, 10) for r in range(10)] == list(range(10))
    assert oracle.sm_depth(15, 16) == 30 and oracle.sm_depth(15, 30) == 16


def test_align_matches_reference_golden(oracle, golden):
    cases = json.load(open(os.path.join(G, "align_cases.json")))
    assert len(cases) >= 400
    for i, c in enumerate(cases):
        mask = None if c["mask"] is None else np.array(c["mask"], np.uint8)
        a = oracle.align(c["ref"], c["read"], golden[c["mat"]], c["sg5"], mask)
        got = [a["score"], a["abr"], a["abc"], a["aer"], a["aec"], a["ref_gapped"], a["read_gapped"]]
        assert got == c["out"], f"case {i}"


def _run_session(oracle, golden, name, matrix):
    s = json.load(open(os.path.join(G, "sessions.json")))[name]
    R = OracleRun(oracle, s["ref"], golden[matrix], s["circular"], s["k"], s["soft_mask"], repeat_filt=s.get("repeat_filt", 0),
                  just_outer_coords=s.get("just_outer_coords", 1))
    for i, (rd, exp) in enumerate(zip(s["reads"], s["pass1"])):
        if exp is None:
            continue
        p = R.pass1(rd, qual_sum=s["qual_sums"][i] if "qual_sums" in s else 0)
        for key, v in exp.items():
            if not exp["hits"] and key not in ("hits", "added"):
                continue
            assert p[key] == v, f"{name}: pass1 read {i} field {key}"
    R.end_pass1()
    for it, exp in enumerate(s["iters"]):
        cons, conv = R.iterate()
        assert [[f["score"], f["as_"], f["ae"], f["rc"]] for f in R.fsdb] == exp["reads"], f"{name}: iteration {it} reads"
        if s.get("repeat_filt"):
            assert [f["unique_best"] for f in R.fsdb] == exp["unique"], f"{name}: iteration {it} unique_best"
        slots = [[x["start"], x["end"], x["dropped"], x["segment"], x["seq"], x["smp"], x["ins"]] for x in oracle.asm_entries(R.asm)]
        assert slots == exp["slots"], f"{name}: iteration {it} AlnSeq list"
        assert np.flatnonzero(oracle.asm_gaps(R.asm, R.wrap_len)).tolist() == exp["gaps"]
        assert cons == exp["cons"], f"{name}: iteration {it} consensus"
        assert conv == exp["converged"]
    return len(s["iters"])


def test_session_reference_fixtures_circular(oracle, golden):
    assert _run_session(oracle, golden, "tr1_tf_c", "ancient") == 3


def test_session_reference_fixtures_linear(oracle, golden):
    _run_session(oracle, golden, "tr1_tf_lin", "ancient")


def test_session_reference_fixtures_kmer_softmask(oracle, golden):
    # exercises the k-mer filter with -M and the stale back-pointer behaviour (5 stale slots)
    _run_session(oracle, golden, "tr1_tf_c_k8_M", "ancient")


def test_session_synthetic_circular_kmer(oracle, golden):
    _run_session(oracle, golden, "synth2k_c_k10", "onepass")


def test_session_divergent_seed_kmer(oracle, golden):
    # BASELINE configs[3] in small: starting reference 10 % + indels away from the sample, k = 12
    _run_session(oracle, golden, "synth3k_div10_c_k12", "ancient")


def test_session_merged_pe_long_reads(oracle, golden):
    # BASELINE configs[2] in small: 30-140 bp reads, ancient.submat.solexa.pe
    _run_session(oracle, golden, "synth2k5_pe_long_c_k12", "pe")


def test_find_consensus_rules(oracle):
    # map_align.c:294-391: cov 0 -> N; gaps/cov >= 0.5 -> '-'; '>=' lets the later base win ties
    f = oracle.find_consensus
    assert f([0] * 10) == "N"
    assert f([1, 0, 0, 0, 1, 2, 100, 0, 0, 0]) == "-"
    assert f([1, 0, 0, 0, 1, 3, 100, 0, 0, 0]) == "A"
    assert f([1, 1, 1, 1, 0, 4, 50, 50, 50, 50]) == "T"
    assert f([1, 1, 0, 0, 0, 2, 50, 50, 40, 40]) == "C"
    assert f([1, 0, 0, 0, 0, 1, -400, -500, -500, -500]) == "N"      # MIN_SCORE_CONS = -399
    assert f([1, 0, 0, 0, 0, 1, -399, -500, -500, -500]) == "A"
    assert f([1, 0, 0, 0, 0, 1, -1, -2401, -3000, -3000], 2) == "N"  # cons_code 2: diff must EXCEED 2400
    assert f([1, 0, 0, 0, 0, 1, -1, -2402, -3000, -3000], 2) == "A"


def test_session_repeat_filter(oracle, golden):
    # -u: sort_fsdb + set_uniq_in_fsdb every round (the FSDB order itself changes), circular; -u -A, linear
    _run_session(oracle, golden, "synth1k5_dups_c_k10_u", "onepass")
    _run_session(oracle, golden, "synth1k5_dups_lin_k10_uA", "onepass")
    _run_session(oracle, golden, "synth1k5_dups_c_k10_U", "onepass")          # -U: duplicates decided by FragSeq.qual_sum


def test_repeat_filter_matches_reference_golden(oracle):
    # f1: the oracle's stable sort + set_uniq_in_fsdb against what the reference's own functions returned
    cases = json.load(open(os.path.join(G, "repeat_cases.json")))
    assert len(cases) == 40
    for i, c in enumerate(cases):
        order, uniq = oracle.repeat_filter(np.array(c["rc"], np.uint8), np.array(c["as_"], np.int32), np.array(c["ae"], np.int32),
                                           np.array(c["key4"], np.int32), np.array(c["trimmed"], np.uint8), c["just_outer"], c["tolerance"])
        assert order.tolist() == c["order"] and uniq.tolist() == c["unique"], i


def test_trim_matches_reference_golden(oracle):
    # f4: the oracle's trim_frag against what the reference's own function returned
    cases = json.load(open(os.path.join(G, "trim_cases.json")))
    assert len(cases) == 300
    for i, c in enumerate(cases):
        t = oracle.trim(c["read"], c["adapter"])
        assert [t[k] for k in ("trimmed", "trim_point", "score", "abr", "abc", "aer")] == c["out"], (i, c["read"])


# ---- round 2: the reference's FragSeq -> AlnSeq pointer behaviour (tests/golden/make_golden_r2.py)
def _load_r2(name):
    import gzip
    return json.load(gzip.open(os.path.join(G, "sessions_r2.json.gz"), "rt"))[name]


def _run_session_r2(oracle, golden, name, matrix):
    s = _load_r2(name)
    R = OracleRun(oracle, s["ref"], golden[matrix], s["circular"], s["k"], 0, distant_ref=s["distant_ref"])
    for i, (rd, exp) in enumerate(zip(s["reads"], s["pass1"])):
        p = R.pass1(rd)
        for key, v in exp.items():
            if not exp["hits"] and key not in ("hits", "added"):
                continue
            assert p[key] == v, f"{name}: pass1 read {i} field {key}"
    R.end_pass1()
    for it, exp in enumerate(s["iters"]):
        cons, conv = R.iterate()
        assert [[f["score"], f["as_"], f["ae"], f["rc"], f["strand_known"]] for f in R.fsdb] == exp["reads"], f"{name}: iteration {it} reads"
        slots = [[x["start"], x["end"], x["dropped"], x["segment"], x["seq"], x["smp"], x["ins"]] for x in oracle.asm_entries(R.asm)]
        assert slots == [x[1:] for x in exp["slots"]], f"{name}: iteration {it} AlnSeq list"
        assert np.flatnonzero(oracle.asm_gaps(R.asm, R.wrap_len)).tolist() == exp["gaps"]
        assert cons == exp["cons"], f"{name}: iteration {it} consensus"
        assert conv == exp["converged"]
    return R


def test_session_reads_scoring_exactly_2000(oracle, golden):
    # strand_known = 0 (mia.c:1653): never realigned, the pass-1 AlnSeq pointers alias other reads' slots or older content
    R = _run_session_r2(oracle, golden, "flat_2000_c", "flat")
    assert sum(1 for f in R.fsdb if not f["strand_known"]) == 14


@pytest.mark.parametrize("name", ["origin305_splitflip_c", "origin303_splitflip_c"])
def test_session_split_pattern_changes(oracle, golden, name):
    # reads flip between wrap-split and whole: slot numbers slide under the sticky flags, back_asp goes stale (mia_main.c:268-276)
    _run_session_r2(oracle, golden, name, "onepass")


@pytest.mark.parametrize("name", ["synth3k_div10_c_k12_D", "synth1k_N_lin_D"])
def test_session_distant_reference(oracle, golden, name):
    # mia -D: accept everything with a k-mer hit, retry strand-unknown reads on both strands of the whole reference from
    # iteration 2 on with whatever matrix the previous read left (H6), find_alignable_len in the cull
    _run_session_r2(oracle, golden, name, "ancient")


# ---- mia -h: the homopolymer-discounted gap candidates (tests/golden/make_golden_hp.py)
def _load_hp():
    import gzip
    return json.load(gzip.open(os.path.join(G, "hp.json.gz"), "rt"))


def test_align_homopolymer_discount_matches_reference_golden(oracle, golden):
    mats = {"flat": golden["flat"], "onepass": golden["onepass"], "ancient": golden["ancient"]}
    for i, c in enumerate(_load_hp()["align"]):
        a = oracle.align(c["ref"], c["read"], mats[c["matrix"]], c["sg5"], None, hp=1)
        assert [a["score"], a["abr"], a["abc"], a["aer"], a["aec"], a["ref_gapped"], a["read_gapped"]] == c["out"], i


@pytest.mark.parametrize("name", ["hp2k_c_k10_h", "hp2k_lin_k12_hD"])
def test_session_homopolymer_discount(oracle, golden, name):
    # mia -h and mia -h -D as whole sessions: pass 1 with the homopolymers of the whole strands, the rounds with those of the windows
    s = _load_hp()["sessions"][name]
    R = OracleRun(oracle, s["ref"], golden[s["matrix"]], s["circular"], s["k"], 0, distant_ref=s["distant_ref"], hp=1)
    for i, (rd, exp) in enumerate(zip(s["reads"], s["pass1"])):
        p = R.pass1(rd)
        for key, v in exp.items():
            if not exp["hits"] and key not in ("hits", "added"):
                continue
            assert p[key] == v, f"{name}: pass1 read {i} field {key}"
    R.end_pass1()
    for it, exp in enumerate(s["iters"]):
        cons, conv = R.iterate()
        assert [[f["score"], f["as_"], f["ae"], f["rc"], f["strand_known"]] for f in R.fsdb] == exp["reads"], f"{name}: iteration {it} reads"
        slots = [[x["start"], x["end"], x["dropped"], x["segment"], x["seq"], x["smp"], x["ins"]] for x in oracle.asm_entries(R.asm)]
        assert slots == [x[1:] for x in exp["slots"]], f"{name}: iteration {it} AlnSeq list"
        assert cons == exp["cons"], f"{name}: iteration {it} consensus"
        assert conv == exp["converged"]
