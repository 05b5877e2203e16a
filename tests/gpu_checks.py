"""Parity checks of the CUDA path against the CPU oracle, shared by tests/ and smoke().
Everything goes through the C ABI (mia_b200.api)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_pssm(name="onepass"):
    return np.load(os.path.join(ROOT, "tests", "golden", "pssm.npz"))[name]


def make_case(n_reads, ref_len, seed, divergence=0.02, indel_rate=0.004, min_len=35, max_len=75, n_rate=0.002):
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(ref_len, seed=seed)
    genome = synth.diverge(ref, divergence, seed=seed + 1, indel_rate=indel_rate)
    bases, off, truth = synth.make_reads(genome, n_reads, min_len, max_len, seed=seed + 2, n_rate=n_rate)
    # stored orientation: reverse-strand reads are kept reverse-complemented (fsdb.c:209-227)
    rc = truth["strand"].astype(np.uint8)
    out = bases.copy()
    for i in np.flatnonzero(rc):
        out[off[i]:off[i + 1]] = synth.revcomp_bytes(bases[off[i]:off[i + 1]])
    # as/ae guesses: true start on the sample genome +- a few bases (indels shift them)
    rng = np.random.default_rng(seed + 3)
    as_ = (truth["start"] + rng.integers(-6, 7, n_reads)).astype(np.int32)
    as_ = np.clip(as_, 0, ref_len - 1)
    ae = (as_ + truth["length"] - 1 + rng.integers(-3, 4, n_reads)).astype(np.int32)
    return ref, out, off, rc, as_, ae


def check_realign(gpu, oracle, ref, bases, off, rc, as_, ae, sm, circular=1, sample=None, hp=0):
    """GPU realign vs oracle realign, read by read: score, as, ae, abr, gapped strings.  hp: mia -h on both sides."""
    from mia_b200 import api
    gpu.set_pssm(sm)
    gpu.set_reference(ref, circular=circular, with_rc=0)
    gpu.set_homopolymer(hp)
    try:
        out = gpu.realign_host(bases, off, rc, as_, ae)
    finally:
        gpu.set_homopolymer(0)
    ctx = oracle.ctx_new(ref, circular, sm, with_rc=0, k=0, hp=hp)
    wref = oracle.ctx_seq(ctx)
    n = len(off) - 1
    idx = range(n) if sample is None else sample
    bad = []
    for i in idx:
        read = bases[off[i]:off[i + 1]].tobytes().decode()
        o = oracle.realign(ctx, read, int(rc[i]), int(as_[i]), int(ae[i]))
        st = int(out["status"][i])
        if st & 0x80:
            bad.append((i, "unsupported window", st))
            continue
        rg, fg = api.expand_runs(wref, read, int(out["as_out"][i]), int(out["abr"][i]), out["runs"][i], int(out["n_runs"][i]))
        got = (int(out["score"][i]), int(out["as_out"][i]), int(out["ae_out"][i]), int(out["abr"][i]), rg, fg)
        exp = (o["score"], o["as_"], o["ae"], o["abr"], o["ref_gapped"], o["read_gapped"])
        if got != exp:
            bad.append((i, got, exp))
    oracle.ctx_free(ctx)
    return bad, out


def check_realign_small(n_reads=2000, ref_len=3000, seed=7):
    import _pkg
    _pkg.load()
    from mia_b200 import api
    from oracle.pyoracle import Oracle
    o = Oracle()
    g = api.MiaGpu(0)
    ref, bases, off, rc, as_, ae = make_case(n_reads, ref_len, seed)
    bad, _ = check_realign(g, o, ref, bases, off, rc, as_, ae, load_pssm("onepass"))
    g.close()
    assert not bad, f"{len(bad)} of {n_reads} reads differ from the oracle; first: {bad[0]}"


def oracle_round(oracle, ref, bases, off, rc, res, sm, circular, dropped_by_read=None, cons_code=1):
    """Oracle assembly of one round from per-read realign results (natural pointers):
    returns (consensus, gaps, counts, entries-as-slots)."""
    import numpy as np
    smr = oracle.revcom_pssm(sm)
    ctx = oracle.ctx_new(ref, circular, sm, with_rc=0, k=0)
    wrap = oracle.lib.orc_ctx_wrap_len(ctx)
    a = oracle.asm_new()
    oracle.asm_begin_round(a, len(ref), wrap)
    n = len(off) - 1
    front, back = np.full(n, -1, np.int32), np.full(n, -1, np.int32)
    for i in range(n):
        r = res[i]
        if r is None:
            continue
        f, b = oracle.asm_add(a, r["ref_gapped"], r["read_gapped"], r["as_"], r["ae"], int(rc[i]), r["score"])
        front[i] = f
        back[i] = -1 if b is None else b
    oracle.asm_pop_smp(a, front, back)
    seq_len = np.diff(off).astype(np.int32)
    if dropped_by_read is None:
        score = np.array([0 if r is None else r["score"] for r in res], np.int32)
        oracle.asm_cull(a, front, back, seq_len, score)
    else:
        # force the flags: score 1 / 0 against a hard cut of 1
        score = np.where(np.asarray(dropped_by_read) > 0, 0, 1).astype(np.int32)
        oracle.asm_cull(a, front, back, seq_len, score, hard_cut=1)
    cons, counts = oracle.asm_consensus(a, sm, smr, cons_code, len(ref), counts=True)
    gaps = oracle.asm_gaps(a, wrap)
    ent = oracle.asm_entries(a)
    oracle.asm_free(a)
    oracle.ctx_free(ctx)
    return cons, gaps, counts, ent


def check_consensus(gpu, oracle, ref, bases, off, rc, as_, ae, sm, circular=1, cons_code=1, drop_frac=0.1, seed=0):
    """GPU realign + consensus vs the oracle's round on the same inputs."""
    import numpy as np
    from mia_b200 import entries as E
    gpu.set_pssm(sm)
    gpu.set_reference(ref, circular=circular, with_rc=0)
    out = gpu.realign_host(bases, off, rc, as_, ae)
    n = len(off) - 1
    ctx = oracle.ctx_new(ref, circular, sm, with_rc=0, k=0)
    res = []
    for i in range(n):
        read = bases[off[i]:off[i + 1]].tobytes().decode()
        res.append(oracle.realign(ctx, read, int(rc[i]), int(as_[i]), int(ae[i])))
    oracle.ctx_free(ctx)
    rng = np.random.default_rng(seed)
    dropped = (rng.random(n) < drop_frac).astype(np.uint8)
    cons, gaps, counts, _ = oracle_round(oracle, ref, bases, off, rc, res, sm, circular, dropped, cons_code)
    ent, split, _ = E.natural_entries(out["as_out"], out["ae_out"], out["n_runs"], out["runs"], len(ref), dropped, dropped)
    gcons, ggaps, gcounts = gpu.consensus(ent, cons_code, want_counts=True)
    problems = []
    if not (ggaps == gaps[:len(ref)]).all():
        problems.append(("gaps", np.flatnonzero(ggaps != gaps[:len(ref)])[:5].tolist()))
    if not (gcounts == counts).all():
        w = np.argwhere(gcounts != counts)[:5].tolist()
        problems.append(("counts", w, [(gcounts[r].tolist(), counts[r].tolist()) for r, _ in w[:2]]))
    if gcons != cons:
        problems.append(("consensus", len(gcons), len(cons)))
    return problems, dict(n_split=int(split.sum()), n_ins_cols=int(ggaps.sum()), cons_len=len(gcons))


def check_pass1(gpu, oracle, ref, reads, sm, circular=1, k=0, soft_mask=0, hp=0):
    """GPU pass 1 (k-mer filter + both-strand DP + strand pick + traceback + coordinates) vs oracle.  hp: mia -h on both sides."""
    import numpy as np
    from mia_b200 import api
    gpu.set_pssm(sm)
    gpu.set_reference(ref, circular=circular, with_rc=1)
    gpu.build_kmers(k, soft_mask)
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    gpu.upload_reads(bases, off)
    gpu.set_homopolymer(hp)
    try:
        out = gpu.pass1()
    finally:
        gpu.set_homopolymer(0)
    ctx = oracle.ctx_new(ref, circular, sm, with_rc=1, k=k, soft_mask=soft_mask, hp=hp)
    wref = oracle.ctx_seq(ctx)
    bad = []
    for i, rd in enumerate(reads):
        o = oracle.pass1(ctx, rd)
        if int(out["hits"][i]) != o["hits"]:
            bad.append((i, "hits", int(out["hits"][i]), o["hits"]))
            continue
        if not o["hits"]:
            if not (int(out["status"][i]) & 2):
                bad.append((i, "not flagged skipped"))
            continue
        got = tuple(int(out[k2][i]) for k2 in ("score", "fw_score", "rc_score", "rc", "as_", "ae", "start", "end"))
        exp = (o["score"], o["fw_score"], o["rc_score"], o["rc"], o["as_"], o["ae"], o["start"], o["end"] if not o["split"] else o["b_end"])
        if got != exp:
            bad.append((i, got, exp))
            continue
        if int(out["status"][i]) & 1:                     # more than MIAGPU_MAX_RUNS runs: reported, the strings cannot be compared
            n_gap_runs = sum(1 for x, y in zip(" " + o["f_ref"] + o["b_ref"], o["f_ref"] + o["b_ref"]) if y == "-" and x != "-") + \
                sum(1 for x, y in zip(" " + o["f_frag"] + o["b_frag"], o["f_frag"] + o["b_frag"]) if y == "-" and x != "-")
            if 2 * n_gap_runs + 1 <= 24:
                bad.append((i, "runs overflow reported for an alignment of", 2 * n_gap_runs + 1, "runs"))
            continue
        # gapped strings in forward-reference orientation: the stored read is revcomp'd for rc
        stored = oracle.revcom(rd) if o["rc"] else rd
        rg, fg = api.expand_runs(wref, stored, int(out["start"][i]), int(out["abr"][i]), out["runs"][i], int(out["n_runs"][i]))
        if (rg, fg) != (o["f_ref"] + o["b_ref"], o["f_frag"] + o["b_frag"]):
            bad.append((i, "strings", rg, fg, o["f_ref"] + o["b_ref"], o["f_frag"] + o["b_frag"]))
    oracle.ctx_free(ctx)
    return bad, out


# ---------------------------------------------------------------- parity of a whole one-call round on a prefix of a workload
class Checker:
    """The CPU side of a parity check: the unmodified reference (oracle/_ref) where it was built, else the oracle restatement.
    Per read it runs the reference's own call sequence on the window reiterate_assembly would cut (mia_main.c:190-235)."""

    def __init__(self):
        from oracle.pyoracle import Oracle, Ref, have_ref
        self.o = Oracle()
        self.r = Ref() if have_ref() else None
        self.kind = "oracle/_ref (unmodified reference)" if self.r else "oracle (restatement)"

    def realign(self, wrap, read, rc, as_, ae, sm, smr):
        L, W = len(read), len(wrap)
        rs = max(0, as_ - 50)
        re = W if ae + 51 > W else ae + 50
        if rs + L > re:
            rs, re = 0, W
        a = (self.r or self.o).align(wrap[rs:re], read, smr if rc else sm, 1)
        return dict(score=a["score"], as_=a["abc"] + rs, ae=a["aec"] + rs, abr=a["abr"], ref_gapped=a["ref_gapped"], read_gapped=a["read_gapped"])


def round_parity(gpu, ref, bases, off, rc, as_, ae, sm, circular=1, checker=None, want_consensus=True):
    """One resident one-call round (miagpu_iterate_resident under miagpu_set_fsdb) over the given reads against the CPU checker:
    every read's score / as / ae / abr / gapped strings, and the round's consensus after the checker's own score cut.
    -> dict(reads, per_read_equal, n_diff, first_diff, gapped_reads, consensus_equal, checker)"""
    from mia_b200 import api
    ck = checker or Checker()
    o = ck.o
    n = len(off) - 1
    seq_len = np.diff(off).astype(np.int32)
    gpu.set_pssm(sm)
    gpu.set_reference(ref, circular=circular, with_rc=0)
    gpu.upload_reads(bases, off)
    gpu.set_alignment_inputs(rc, as_, ae)
    gpu.set_fsdb(seq_len, np.full(n, 2001, np.int32))
    cons, fit, _ = gpu.iterate_resident()
    al = gpu.get_alignment()
    tot, _, _ = gpu.get_runs_packed()
    run_off, packed = np.zeros(n + 1, np.int64), np.zeros(max(tot, 1), np.uint16)
    gpu.get_runs_packed(run_off, packed)
    wrap = ref + (ref[:min(256, len(ref))] if circular else "")
    smr = o.revcom_pssm(sm)
    res, bad, gapped = [], [], 0
    for i in range(n):
        read = bases[off[i]:off[i + 1]].tobytes().decode()
        e = ck.realign(wrap, read, int(rc[i]), int(as_[i]), int(ae[i]), sm, smr)
        res.append(e)
        runs = packed[run_off[i]:run_off[i + 1]]
        gapped += len(runs) > 1
        rg, fg = api.expand_runs(wrap, read, int(al["as_out"][i]), int(al["abr"][i]), runs, len(runs))
        got = (int(al["score"][i]), int(al["as_out"][i]), int(al["ae_out"][i]), int(al["abr"][i]), rg, fg)
        exp = (e["score"], e["as_"], e["ae"], e["abr"], e["ref_gapped"], e["read_gapped"])
        if got != exp:
            bad.append((i, got[:4], exp[:4]))
    out = dict(reads=n, per_read_equal=not bad, n_diff=len(bad), first_diff=str(bad[0]) if bad else None, gapped_reads=int(gapped),
               checker=ck.kind)
    if want_consensus:
        ocons, _, _, _ = oracle_round(o, ref, bases, off, rc, res, sm, circular)
        out["consensus_equal"] = bool(ocons == cons)
        out["consensus_len"] = len(cons)
    out["consensus"] = cons
    return out


def assembly_parity(gpu, ref, reads, sm, circular=1, k=12, distant_ref=0, max_iter=30, hp=0):
    """A whole assembly (pass 1 with the k-mer filter, rounds until the consensus stops changing) of `reads` (list of str) through
    driver.ResidentAssembler on one GPU against the same assembly by the CPU checker -- the unmodified reference's own main loop
    (oracle/ref_harness.c: refh_sess_*) where oracle/_ref was built, else the oracle restatement (tests/oracle_driver.py): pass-1
    results of every read, per round every read's score / as / ae / rc / strand_known and the consensus.
    -> dict(reads, fsdb, rounds, pass1_equal, rounds_equal, consensus_equal, converged_equal, checker, first_diff)"""
    import tempfile
    from mia_b200 import driver
    from oracle.pyoracle import Oracle, Ref, have_ref
    n = len(reads)
    off = np.zeros(n + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    A = driver.ResidentAssembler(gpu, ref, sm, circular, k, 0, distant_ref=distant_ref, hp=hp)
    p = A.pass1(bases, off)
    first_diff = None
    exp_iters = []
    if have_ref():
        r = Ref()
        kind = "oracle/_ref (the unmodified reference's main loop)"
        with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
            f.write(">ref\n" + ref + "\n")
            path = f.name
        s = r.sess_new(path, circular, sm, k=k, soft_mask=0, distant_ref=distant_ref, hp=hp)
        r.set_hp(0)
        p1 = [r.sess_pass1(s, "r%d" % i, rd) for i, rd in enumerate(reads)]
        r.sess_end_pass1(s)
        for _ in range(max_iter):
            cons, conv = r.sess_iterate(s, sort=0)
            exp_iters.append((cons, conv, [[x["score"], x["as_"], x["ae"], x["rc"], x["strand_known"]] for x in r.sess_reads(s)]))
            if conv:
                break
        os.unlink(path)
    else:
        from oracle_driver import OracleRun
        o = Oracle()
        kind = "oracle (restatement)"
        R = OracleRun(o, ref, sm, circular, k, 0, distant_ref=distant_ref, hp=hp)
        p1 = [R.pass1(rd) for rd in reads]
        R.end_pass1()
        for _ in range(max_iter):
            cons, conv = R.iterate()
            exp_iters.append((cons, conv, [[f["score"], f["as_"], f["ae"], f["rc"], f["strand_known"]] for f in R.fsdb]))
            if conv:
                break
    pass1_equal = True
    for i, e in enumerate(p1):
        got = [int(p["hits"][i])] + ([int(p[k2][i]) for k2 in ("score", "rc", "as_", "ae")] if e["hits"] else [])
        exp = [e["hits"]] + ([e[k2] for k2 in ("score", "rc", "as_", "ae")] if e["hits"] else [])
        if got != exp:
            pass1_equal = False
            first_diff = first_diff or f"pass 1 read {i}: {got} != {exp}"
    rounds_equal = cons_equal = conv_equal = True
    for it, (cons, conv, rd) in enumerate(exp_iters):
        gc, gconv = A.iterate()
        got = np.stack([A.score, A.as_, A.ae, A.rc.astype(np.int32), A.strand_known.astype(np.int32)], 1).tolist()
        if got != rd:
            rounds_equal = False
            first_diff = first_diff or f"round {it + 1}: per-read results differ"
        if gc != cons:
            cons_equal = False
            first_diff = first_diff or f"round {it + 1}: consensus differs"
        conv_equal &= gconv == conv
    if hp:
        gpu.set_homopolymer(0)
    return dict(reads=n, fsdb=len(A.seq_len), rounds=len(exp_iters), pass1_equal=pass1_equal, rounds_equal=rounds_equal, consensus_equal=cons_equal,
                converged_equal=bool(conv_equal), checker=kind, first_diff=first_diff)
