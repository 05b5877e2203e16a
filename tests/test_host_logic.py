"""CPU: host-side logic of the product (entry-list construction) against the oracle's AlnSeq
bookkeeping, without a GPU.  Alignments come from the oracle; the run lists are derived from
its gapped strings exactly as the device would emit them."""
import numpy as np

import gpu_checks


def _runs_from_strings(rg, fg):
    runs, i = [], 0
    while i < len(rg):
        t = 1 if rg[i] == "-" else 2 if fg[i] == "-" else 0
        j = i
        while j < len(rg) and (1 if rg[j] == "-" else 2 if fg[j] == "-" else 0) == t:
            j += 1
        runs.append((t << 14) | (j - i))
        i = j
    return runs


def test_natural_entries_match_oracle_slots(oracle):
    import _pkg
    _pkg.load()
    from mia_b200 import api, entries as E
    sm = gpu_checks.load_pssm("onepass")
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(1200, 1500, seed=71, divergence=0.03, indel_rate=0.02)
    n = len(off) - 1
    ctx = oracle.ctx_new(ref, 1, sm, with_rc=0, k=0)
    res, runs, n_runs = [], np.zeros((n, api.MAX_RUNS), np.uint16), np.zeros(n, np.int32)
    for i in range(n):
        r = oracle.realign(ctx, bases[off[i]:off[i + 1]].tobytes().decode(), int(rc[i]), int(as_[i]), int(ae[i]))
        res.append(r)
        rr = _runs_from_strings(r["ref_gapped"], r["read_gapped"])
        assert len(rr) <= api.MAX_RUNS
        runs[i, :len(rr)] = rr
        n_runs[i] = len(rr)
        # expand_runs is the inverse
        assert api.expand_runs(oracle.ctx_seq(ctx), bases[off[i]:off[i + 1]].tobytes().decode(), r["as_"], r["abr"], runs[i], n_runs[i]) == \
            (r["ref_gapped"], r["read_gapped"])
    oracle.ctx_free(ctx)
    as_out = np.array([r["as_"] for r in res], np.int32)
    ae_out = np.array([r["ae"] for r in res], np.int32)
    ent, split, nslots = E.natural_entries(as_out, ae_out, n_runs, runs, len(ref))
    _, _, _, slots = gpu_checks.oracle_round(oracle, ref, bases, off, rc, res, sm, 1)
    assert len(ent) == len(slots) and split.sum() > 0
    for e, s in zip(ent, slots):
        assert e["ref_pos"] == s["start"] and e["col_count"] == len(s["seq"])
        assert e["back_formula"] == (1 if s["segment"] == "b" else 0)
    # asp_len bookkeeping: total_len = columns + inserted bases over both segments
    for i in np.flatnonzero(split)[:20]:
        a = int(nslots[:i].sum())
        ins = lambda s: sum(len(x.split(":")[1]) for x in s["ins"].split(";") if x)
        fl = len(slots[a]["seq"]) + ins(slots[a])
        bl = len(slots[a + 1]["seq"]) + ins(slots[a + 1])
        assert ent[a]["front_len"] == fl and ent[a]["total_len"] == fl + bl == ent[a + 1]["total_len"]
        assert ent[a + 1]["col_begin"] == len(slots[a]["seq"])


def test_distant_retry_state_is_folded_over_shards_in_rank_order():
    # -D over shards (driver.ResidentAssembler.retry_begin / retry_end): the matrix the first forward attempt of a shard runs with is
    # what the last read of the shard before it left (mia_main.c:120-174, H6).  Every shard reports the state it leaves as a function
    # of the state it is entered with; the fold over the ranks must hand every shard its true entry state and keep the last shard's
    # exit state for the next round -- checked against walking the shards one after the other.
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    rng = np.random.default_rng(5)

    class StubGpu:
        def __init__(self, after):
            self.after, self.entered = after, None

        def distant_retry_begin(self):
            return 7, list(self.after)

        def distant_retry_end(self, state_in):
            self.entered = state_in
            return 3

    for _ in range(200):
        world = int(rng.integers(1, 7))
        # identity (a shard without reads, or with strand-unknown reads only in round 1), constant 0 / 1, or the swap a chain cannot
        # produce but the fold must still treat as a function
        afters = [[(0, 1), (0, 0), (1, 1), (1, 0)][int(rng.integers(0, 4))] for _ in range(world)]
        carried = int(rng.integers(0, 2))
        want_in, s = [], carried
        for r in range(world):
            want_in.append(s)
            s = afters[r][s]
        for rank in range(world):
            A = object.__new__(driver.ResidentAssembler)
            A.g, A.matrix_state = StubGpu(afters[rank]), carried
            assert A.retry_begin() == list(afters[rank])
            A.retry_end(np.array(afters, np.int32), rank)
            assert A.g.entered == want_in[rank]
            assert A.matrix_state == s and A.retried == (7, 3)
