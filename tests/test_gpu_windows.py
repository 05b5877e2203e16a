"""-m gpu: miagpu_align_windows (the bare dyn_prog client sequence of ccheck.cc:571-603): every read against its own
reference stretch, no window rule.  Checked (a) against the UNMODIFIED reference's outputs for the fuzzed alignments of
tests/golden/align_cases.json that are unmasked (score, abr, abc, aer, aec and both gapped strings), with sg5 = 1 and sg5 = 0 and
with the strand-reversed matrix picked through rc = 1; (b) against the oracle on 4,000 seeded ccheck-shaped pairs (read vs a
piece of consensus about its own length, flat matrix), windows shorter than their reads included."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(gpu, api, refs, reads, sm, rc, sg5):
    ref = "".join(refs)
    ws = np.zeros(len(refs), np.int32)
    np.cumsum([len(r) for r in refs[:-1]], out=ws[1:])
    wl = np.array([len(r) for r in refs], np.int32)
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    gpu.set_pssm(sm)
    gpu.set_reference(ref, circular=0, with_rc=0)
    gpu.upload_reads(np.frombuffer("".join(reads).encode(), np.uint8), off)
    out = gpu.align_windows(np.full(len(reads), rc, np.uint8), ws, wl, sg5)
    assert (out["status"] == 0).all()
    res = []
    runs = out["runs"].view(np.uint16).reshape(-1, api.MAX_RUNS)
    for i in range(len(reads)):
        rg, fg = api.expand_runs(ref, reads[i], int(out["as_out"][i]), int(out["abr"][i]), runs[i], int(out["n_runs"][i]))
        aer = int(out["abr"][i]) + sum(1 for ch in fg if ch != "-") - 1
        res.append([int(out["score"][i]), int(out["abr"][i]), int(out["as_out"][i] - ws[i]), aer, int(out["ae_out"][i] - ws[i]), rg, fg])
    return res


@pytest.mark.parametrize("sg5", [1, 0])
def test_align_windows_equals_reference_golden(gpu, golden, sg5):
    import _pkg
    _pkg.load()
    from mia_b200 import api
    cases = [c for c in json.load(open(os.path.join(G, "align_cases.json"))) if c["mask"] is None and c["sg5"] == sg5]
    assert len(cases) >= 80
    checked = 0
    for mat, base, rc in (("flat", "flat", 0), ("ancient", "ancient", 0), ("onepass", "onepass", 0), ("ancient_rc", "ancient", 1)):
        sel = [c for c in cases if c["mat"] == mat]
        got = _run(gpu, api, [c["ref"] for c in sel], [c["read"] for c in sel], golden[base], rc, sg5)
        for k, (g, c) in enumerate(zip(got, sel)):
            assert g == c["out"], (mat, k, g, c["out"])
        checked += len(sel)
    assert checked == len(cases)


def test_align_windows_ccheck_shaped_pairs_equal_oracle(gpu, golden, oracle):
    import random
    import _pkg
    _pkg.load()
    from mia_b200 import api
    rng = random.Random(77)
    refs, reads = [], []
    for _ in range(4000):
        n = rng.randint(20, 120)
        s = [rng.choice("ACGT") for _ in range(n)]
        r = []
        for ch in s:
            x = rng.random()
            if x < 0.04:
                r.append(rng.choice("ACGT"))
            elif x < 0.06:
                continue
            elif x < 0.08:
                r.append(ch + rng.choice("ACGT"))
            else:
                r.append(ch)
        lo = rng.randint(0, 3)
        hi = n - rng.randint(0, 3)
        piece = "".join(s[lo:hi])
        if rng.random() < 0.05:
            piece = piece[: max(5, len(piece) // 2)]            # window much shorter than the read
        if rng.random() < 0.1:
            piece = piece.replace(piece[len(piece) // 2], "N", 1)  # ccheck turns non-ACGT consensus characters into N
        refs.append(piece)
        reads.append("".join(r)[:256] or "A")
    got = _run(gpu, api, refs, reads, golden["flat"], 0, 1)
    for k in range(len(refs)):
        a = oracle.align(refs[k], reads[k], golden["flat"], 1, None)
        want = [a["score"], a["abr"], a["abc"], a["aer"], a["aec"], a["ref_gapped"], a["read_gapped"]]
        assert got[k] == want, (k, refs[k], reads[k], got[k], want)


def test_align_windows_rejects_windows_outside_the_reference(gpu, golden):
    import _pkg
    _pkg.load()
    from mia_b200 import api
    gpu.set_pssm(golden["flat"])
    gpu.set_reference("ACGTACGTAC", circular=0, with_rc=0)
    gpu.upload_reads(np.frombuffer(b"ACGT", np.uint8), np.array([0, 4], np.int64))
    with pytest.raises(api.MiaGpuError, match="leaves the reference"):
        gpu.align_windows(np.zeros(1, np.uint8), np.array([8], np.int32), np.array([5], np.int32))
