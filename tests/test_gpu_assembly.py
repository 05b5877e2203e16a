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


def _session_inputs(name):
    s = json.load(open(os.path.join(G, "sessions.json")))[name]
    reads = [r for r in s["reads"] if r]
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    return s, bases, off


@pytest.mark.parametrize("name,matrix", [("synth2k_c_k10", "onepass"), ("synth3k_div10_c_k12", "ancient"), ("synth2k5_pe_long_c_k12", "pe"),
                                         ("tr1_tf_lin", "ancient"), ("tr1_tf_c", "ancient"), ("tr1_tf_c_k8_M", "ancient")])
def test_resident_rounds_reproduce_reference_sessions(gpu, golden, name, matrix):
    # everything resident, one library call per round (score cut on the device): the reference's consensus of every round
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    s, bases, off = _session_inputs(name)
    A = driver.ResidentAssembler(gpu, s["ref"], golden[matrix], s["circular"], s["k"], s["soft_mask"])
    A.pass1(bases, off)
    for it, e in enumerate(s["iters"]):
        cons, conv = A.iterate(want_gaps=True)
        got = np.stack([A.score, A.as_, A.ae, A.rc.astype(np.int32)], 1).tolist()
        assert got == e["reads"], f"{name}: iteration {it + 1} per-read score/as/ae"
        assert np.flatnonzero(A.gaps).tolist() == [g for g in e["gaps"] if g < len(A.last)], f"{name}: iteration {it + 1} gaps"
        assert cons == e["cons"], f"{name}: iteration {it + 1} consensus"
        assert conv == e["converged"]


def _load_r2(name):
    import gzip
    return json.load(gzip.open(os.path.join(G, "sessions_r2.json.gz"), "rt"))[name]


@pytest.mark.parametrize("name,matrix", [("flat_2000_c", "flat"), ("origin305_splitflip_c", "onepass"), ("origin303_splitflip_c", "onepass"),
                                         ("synth3k_div10_c_k12_D", "ancient"), ("synth1k_N_lin_D", "ancient")])
def test_resident_rounds_follow_the_reference_pointers(gpu, golden, name, matrix):
    # the reference's FragSeq -> AlnSeq pointer behaviour in the one-call resident rounds (tests/golden/make_golden_r2.py): reads that
    # score exactly 2000 (strand_known = 0), split patterns that change (slot-indexed sticky flags, stale back pointers), -D
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    s = _load_r2(name)
    reads = s["reads"]
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    A = driver.ResidentAssembler(gpu, s["ref"], golden[matrix], s["circular"], s["k"], 0, distant_ref=s["distant_ref"])
    p = A.pass1(bases, off)
    for i, e in enumerate(s["pass1"]):
        assert int(p["hits"][i]) == e["hits"], (name, i)
        if e["hits"]:
            got = [int(p[k][i]) for k in ("score", "rc", "as_", "ae", "start")]
            assert got == [e[k] for k in ("score", "rc", "as_", "ae", "start")], (name, i)
    assert A.fsdb_idx.tolist() == s["iters"][0]["ids"]
    stale = 0
    for it, e in enumerate(s["iters"]):
        cons, conv = A.iterate(want_gaps=True)
        got = np.stack([A.score, A.as_, A.ae, A.rc.astype(np.int32), A.strand_known.astype(np.int32)], 1).tolist()
        assert got == e["reads"], f"{name}: iteration {it + 1} per-read score/as/ae/rc/strand_known"
        st = gpu.last_fsdb_stats()
        stale += st["stale_pointers"]
        n_drop_exp = sum(x[3] for x in e["slots"])
        assert np.flatnonzero(A.gaps).tolist() == [g for g in e["gaps"] if g < len(A.last)], f"{name}: iteration {it + 1} gaps"
        assert cons == e["cons"], f"{name}: iteration {it + 1} consensus ({st}, {n_drop_exp} AlnSeqs dropped in the reference)"
        assert conv == e["converged"]
    assert stale > 0, "the session was made to exercise stale pointers"


@pytest.mark.parametrize("name,matrix,parts", [("flat_2000_c", "flat", 2), ("origin305_splitflip_c", "onepass", 2), ("origin303_splitflip_c", "onepass", 3),
                                               ("origin305_splitflip_c", "onepass", 3), ("synth3k_div10_c_k12_D", "ancient", 2),
                                               ("synth3k_div10_c_k12_D", "ancient", -3), ("synth1k_N_lin_D", "ancient", -2)])
def test_sharded_rounds_follow_the_reference_pointers(gpu, golden, name, matrix, parts):
    # the pointer state with the reads on several shards (contexts of one process, collectives emulated by device copies): global slot
    # numbers, slot flags replicated, stale pointers resolved on the shard that holds them -- the reference's rounds, read by read.
    # A stale pointer that crosses a shard boundary is refused (miagpu.h): the test moves the boundaries until none does.
    # -D: the whole-reference attempts of every shard's strand-unknown reads, the matrix state (H6) handed from shard to shard
    # (miagpu_distant_retry_begin / _end), find_alignable_len of all ranks' reads in the pass-1 cull.  Under -D every strand-unknown
    # read holds stale pass-1 pointers whose targets slide by the number of unknown reads before them, so with more shards some
    # always cross a boundary: parts < 0 marks the cases where either outcome is accepted -- the reference's rounds, or the refusal.
    may_refuse, parts = parts < 0, abs(parts)
    import _pkg
    _pkg.load()
    from mia_b200 import api, driver, shard
    s = _load_r2(name)
    reads = s["reads"]
    n = len(reads)
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    refused, done = [], False
    for shift in (0.0, 0.07, -0.07, 0.13, -0.13, 0.21):
        cuts = [0] + [int(n * (r / parts + shift / parts)) for r in range(1, parts)] + [n]
        ctxs = [api.MiaGpu(0) for _ in range(parts)]
        try:
            asms = [driver.ResidentAssembler(g, s["ref"], golden[matrix], s["circular"], s["k"], 0, pointer_state=True, distant_ref=s["distant_ref"])
                    for g in ctxs]
            for r, a in enumerate(asms):
                lo, hi = cuts[r], cuts[r + 1]
                a.pass1(np.ascontiguousarray(bases[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo]), defer_cull=True)
                a._manual_retry = True
            all_sl = np.concatenate([a.seq_len for a in asms])
            all_sc = np.concatenate([a.score for a in asms])
            all_sp = np.concatenate([a.split for a in asms]).astype(np.uint8)
            all_cl = np.concatenate([a.cull_len() for a in asms]) if s["distant_ref"] else None
            los = np.concatenate([[0], np.cumsum([len(a.seq_len) for a in asms])])
            for r, a in enumerate(asms):
                a.pass1_cull(all_sl, all_sc, all_sp, int(los[r]), all_cl)
            L = shard.LocalShards(ctxs)
            stale = 0
            for it, e in enumerate(s["iters"]):
                for a in asms:
                    a.begin_round()
                if s["distant_ref"]:
                    after = [a.retry_begin() for a in asms]
                    for r, a in enumerate(asms):
                        a.retry_end(after, r)
                res = L.resident(max(len(a.seq_len) for a in asms), dropped=[a.dropped for a in asms])
                outs = [a.end_round(*r) for a, r in zip(asms, res)]
                stale += sum(g.last_fsdb_stats()["stale_pointers"] for g in ctxs)
                for cons, conv in outs:
                    assert cons == e["cons"], f"{name}: iteration {it + 1} consensus (cuts {cuts})"
                    assert conv == e["converged"]
                got = np.concatenate([np.stack([a.score, a.as_, a.ae, a.rc.astype(np.int32), a.strand_known.astype(np.int32)], 1) for a in asms]).tolist()
                assert got == e["reads"], f"{name}: iteration {it + 1} per-read results (cuts {cuts})"
            assert stale > 0, "the session was made to exercise stale pointers"
            done = True
        except api.MiaGpuError as ex:
            if "another rank" not in str(ex):
                raise
            refused.append((cuts, str(ex)[:120]))
        finally:
            for g in ctxs:
                g.close()
        if done:
            break
    assert done or (may_refuse and refused), f"every sharding was refused: {refused}"


@pytest.mark.parametrize("name,matrix,parts", [("synth3k_div10_c_k12", "ancient", 2), ("synth2k5_pe_long_c_k12", "pe", 3)])
def test_sharded_assembly_to_convergence(gpu, golden, name, matrix, parts):
    # BASELINE configs[2] / [3] in small: pass 1 and every round with the reads sharded (contexts of one process, collectives
    # emulated by device copies) -- the reference's consensus of every round on every shard
    import _pkg
    _pkg.load()
    from mia_b200 import api, driver, shard
    s, bases, off = _session_inputs(name)
    n = len(off) - 1
    ctxs = [api.MiaGpu(0) for _ in range(parts)]
    try:
        asms = [driver.ResidentAssembler(g, s["ref"], golden[matrix], s["circular"], s["k"], s["soft_mask"], pointer_state=False) for g in ctxs]
        for r, a in enumerate(asms):
            lo, hi = n * r // parts, n * (r + 1) // parts
            a.pass1(np.ascontiguousarray(bases[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo]), defer_cull=True)
        all_sl = np.concatenate([a.seq_len for a in asms])
        all_sc = np.concatenate([a.score for a in asms])
        for a in asms:
            a.pass1_cull(all_sl, all_sc)
        L = shard.LocalShards(ctxs)
        for it, e in enumerate(s["iters"]):
            for a in asms:
                a.begin_round()
            res = L.resident(max(len(a.seq_len) for a in asms), dropped=[a.dropped for a in asms])
            outs = [a.end_round(*r) for a, r in zip(asms, res)]
            for cons, conv in outs:
                assert cons == e["cons"], f"{name}: iteration {it + 1} consensus"
                assert conv == e["converged"]
            got = np.concatenate([np.stack([a.score, a.as_, a.ae, a.rc.astype(np.int32)], 1) for a in asms]).tolist()
            assert got == e["reads"], f"{name}: iteration {it + 1} per-read score/as/ae"
    finally:
        for g in ctxs:
            g.close()


@pytest.mark.parametrize("name", ["synth1k5_dups_c_k10_u", "synth1k5_dups_lin_k10_uA", "synth1k5_dups_c_k10_U"])
def test_repeat_filter_sessions_reproduce_reference(gpu, golden, name):
    # mia -u (and -u -A): the FSDB is re-sorted every round, duplicates are left out, sticky flags stay with FSDB positions --
    # per round: the reads in the reference's FSDB order, their unique_best flags, the consensus
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    s, bases, off = _session_inputs(name)
    # repeat_filt 1 = -u (FragSeq.score decides between duplicates), 2 = -U (FragSeq.qual_sum, as read_fastq sums it)
    A = driver.RepeatFilterAssembler(gpu, s["ref"], golden["onepass"], s["circular"], s["k"], s["soft_mask"],
                                     just_outer_coords=s["just_outer_coords"], key="qual" if s["repeat_filt"] == 2 else "score")
    qs = None
    if "qual_sums" in s:
        qs = np.array([q for q, r in zip(s["qual_sums"], s["reads"]) if r], np.int32)
    A.pass1(bases, off, qual_sum=qs)
    for it, e in enumerate(s["iters"]):
        cons, conv = A.iterate()
        fo = A.order
        assert A.ids[fo].tolist() == e["ids"], f"{name}: iteration {it + 1} FSDB order"
        got = np.stack([A.score[fo], A.as_[fo], A.ae[fo], A.rc[fo].astype(np.int32)], 1).tolist()
        assert got == e["reads"], f"{name}: iteration {it + 1} per-read score/as/ae"
        assert A.unique[fo].tolist() == e["unique"], f"{name}: iteration {it + 1} unique_best"
        assert cons == e["cons"], f"{name}: iteration {it + 1} consensus"
        assert conv == e["converged"]
