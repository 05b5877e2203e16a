"""-m gpu: FASTQ text -> miagpu_fastx (f2) -> pass 1 + rounds on the device -> miagpu_write_maln (f3) against the `.maln` files the
UNMODIFIED reference binary wrote for the same FASTQ and flags (tests/golden/maln_session.json.gz; line 1, a time stamp, is left
out): every iteration's file byte for byte -- reference, gaps, matrices, every AlnSeq's id / score / start / end / flags / seq /
smp / inserts, and the list order."""
import gzip
import json
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,matrix", [("circ_k10", "ancient"), ("lin_pe", "pe")])
def test_fastq_to_maln_files_equal_reference(gpu, golden, name, matrix, tmp_path):
    import _pkg
    _pkg.load()
    from mia_b200 import api, driver
    s = json.load(gzip.open(os.path.join(HERE, "golden", "maln_session.json.gz"), "rt"))["sessions"][name]
    flags = s["flags"]
    circular = int("-c" in flags)
    k = int(flags[flags.index("-k") + 1]) if "-k" in flags else 0
    p = tmp_path / "reads.fq"
    p.write_text(s["fastq"])
    rdr = api.FastxReader(path=str(p))
    batch = rdr.next()
    assert rdr.next() is None
    A = driver.ResidentAssembler(gpu, s["ref"], golden[matrix], circular, k, 0)
    A.pass1(batch["bases"], batch["offsets"])
    conv = False
    for it, body in enumerate(s["malns"]):
        assert not conv
        _, conv = A.iterate(want_gaps=True)
        out = str(tmp_path / f"out.{it + 1}")
        A.write_maln(out, batch, s["ref_id"], s["ref_desc"])
        got = open(out).read().split("\n", 1)[1]
        if got != body:
            ga, gb = got.split("\n"), body.split("\n")
            for ln, (x, y) in enumerate(zip(ga, gb)):
                assert x == y, f"{name} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(ga) == len(gb)
    assert conv and A.split_changes == 0


@pytest.mark.parametrize("name,key", [("dups_c_k10_u", "score"), ("dups_c_k10_U", "qual")])
def test_repeat_filter_sessions_fastq_to_maln_files_equal_reference(gpu, golden, name, key, tmp_path):
    # mia -u / mia -U: the FSDB is re-sorted every round, duplicates (by score / by the FASTQ quality sum the reader computed) are left
    # out of the file, AlnSeq.dropped follows the slot-indexed sticky flags -- every iteration's file byte for byte
    import _pkg
    _pkg.load()
    from mia_b200 import api, driver
    s = json.load(gzip.open(os.path.join(HERE, "golden", "maln_session.json.gz"), "rt"))["sessions"][name]
    p = tmp_path / "reads.fq"
    p.write_text(s["fastq"])
    rdr = api.FastxReader(path=str(p))
    batch = rdr.next()
    A = driver.RepeatFilterAssembler(gpu, s["ref"], golden["onepass"], 1, 10, 0, key=key)
    A.pass1(batch["bases"], batch["offsets"], qual_sum=batch["qual_sum"])
    conv = False
    for it, body in enumerate(s["malns"]):
        assert not conv
        _, conv = A.iterate()
        out = str(tmp_path / f"out.{it + 1}")
        A.write_maln(out, batch, s["ref_id"], s["ref_desc"])
        got = open(out).read().split("\n", 1)[1]
        if got != body:
            ga, gb = got.split("\n"), body.split("\n")
            for ln, (x, y) in enumerate(zip(ga, gb)):
                assert x == y, f"{name} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(ga) == len(gb)
    assert conv
