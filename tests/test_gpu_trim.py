"""-m gpu: adapter trimming (miagpu_trim = trim_frag, mia.c:1318-1368) against the oracle, which
tests/test_oracle_vs_ref.py and tests/golden/trim_cases.json pin to the reference's own trim_frag.  Bit-exact: the best
score of the last column, abr / abc / aer, trimmed, trim_point."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NEAND = "GTCAGACACGCAACAGGGGATAGGCAAGGCACACAGGGGATAGG"            # mia_main.c:462


def make_reads(n, adapter, seed, max_len=120):
    rng = random.Random(seed)
    reads = []
    for _ in range(n):
        L = rng.randint(1, max_len)
        rd = "".join(rng.choice("ACGT") for _ in range(L))
        x = rng.random()
        if x < 0.6:                                              # a (damaged) adapter prefix at the 3' end
            frag = list(adapter[: rng.randint(1, len(adapter))])
            for i in range(len(frag)):
                y = rng.random()
                if y < 0.05:
                    frag[i] = rng.choice("ACGT")
                elif y < 0.07:
                    frag[i] = ""
                elif y < 0.09:
                    frag[i] += rng.choice("ACGT")
            rd = (rd + "".join(frag))[:256]
        elif x < 0.7:                                            # adapter in the middle: nothing to trim at the end
            rd = (rd[: L // 2] + adapter[:20] + rd[L // 2:])[:256]
        if rng.random() < 0.05:
            rd = rd[: len(rd) // 2] + "N" + rd[len(rd) // 2 + 1:]
        reads.append(rd or "A")
    return reads


@pytest.mark.parametrize("adapter,max_len", [(NEAND, 60), (NEAND, 120), ("CTGAGACACGCAACAGGGGATAGGCAAGGCACACAGGGGATAGG", 250), ("ACGTTGCA", 100), ("A", 40)])
def test_trim_matches_oracle(gpu, oracle, adapter, max_len):
    reads = make_reads(1500, adapter, seed=len(adapter) + max_len, max_len=max_len)
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    bases = np.frombuffer("".join(reads).encode(), np.uint8)
    out = gpu.trim(bases, off, adapter)
    bad = []
    for i, rd in enumerate(reads):
        o = oracle.trim(rd, adapter)
        got = {k: int(out[k][i]) for k in ("trimmed", "trim_point", "score", "abr", "abc", "aer")}
        if got != o:
            bad.append((i, rd, got, o))
    assert not bad, f"{len(bad)} reads differ; first {bad[0]}"
    assert 0.2 < out["trimmed"].mean() < 0.99


def test_trim_argument_checks(gpu):
    from mia_b200 import api
    off = np.array([0, 4], np.int64)
    with pytest.raises(api.MiaGpuError, match="adapter"):
        gpu.trim(np.frombuffer(b"ACGT", np.uint8), off, "A" * 128)
    with pytest.raises(api.MiaGpuError, match="bases"):
        gpu.trim(np.frombuffer(b"ACGT", np.uint8), np.array([0, 0, 4], np.int64), NEAND)
