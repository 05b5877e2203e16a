"""CPU: the host formats either side of the path against the UNMODIFIED reference's outputs
(tests/golden/maln_session.json.gz, written by tests/golden/make_maln_golden.py from oracle/_ref):

  * miagpu_fastx_*    vs the reference's find_input_type + read_next_seq loop on awkward FASTQ / FASTA texts;
  * miagpu_write_maln vs the reference's own `.maln` files: every AlnSeq of a golden file is turned back into what the
    device returns per read (score, as, ae, abr, run list, stored bases) and the writer must reproduce the file byte
    for byte after line 1 -- seq / ins / smp / "_f" "_b" ids / SEG / list order are all re-derived by the writer.
The same writer is driven from the device's real outputs in tests/test_gpu_maln.py."""
import gzip
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return json.load(gzip.open(os.path.join(HERE, "golden", "maln_session.json.gz"), "rt"))


@pytest.fixture(scope="module")
def api():
    import _pkg
    _pkg.load()
    from mia_b200 import api
    return api


def _records(api, text, batch):
    r = api.FastxReader(text=text)
    recs = []
    while True:
        b = r.next(batch)
        if b is None:
            break
        ids, descs = r.strings(b["ids"], b["id_off"]), r.strings(b["descs"], b["desc_off"])
        for i in range(b["n"]):
            recs.append([ids[i], descs[i], b["bases"][b["offsets"][i]:b["offsets"][i + 1]].tobytes().decode("latin-1"), int(b["qual_sum"][i])])
    r.close()
    return recs


@pytest.fixture(params=["sequential", "parallel"])
def parse_mode(request, monkeypatch):
    """parallel: the speculative multi-threaded parse (taken by itself only above 8 MB) forced onto every text, 4 cursors"""
    if request.param == "parallel":
        monkeypatch.setenv("MIAGPU_FASTX_PAR_MIN", "1")
        monkeypatch.setenv("MIAGPU_FASTX_THREADS", "4")
    return request.param


@pytest.mark.parametrize("batch", [1 << 20, 3, 1])
def test_fastx_reader_equals_reference_reader(api, gold, batch, parse_mode):
    for k, c in enumerate(gold["reader_cases"]):
        if batch == 1 and len(c["records"]) > 50:
            continue
        got = _records(api, c["text"], batch)
        assert len(got) == len(c["records"]), (k, len(got), len(c["records"]))
        for a, b in zip(got, c["records"]):
            assert a == b, (k, a, b)


def test_fastx_reader_from_file(api, gold, tmp_path):
    c = gold["reader_cases"][-1]
    p = tmp_path / "reads.fq"
    p.write_bytes(c["text"].encode("latin-1"))
    r = api.FastxReader(path=str(p))
    assert r.format == 1
    b = r.next()
    assert b["n"] == len(c["records"]) and r.next() is None
    assert r.strings(b["ids"], b["id_off"]) == [x[0] for x in c["records"]]
    with pytest.raises(api.MiaGpuError):
        api.FastxReader(path=str(tmp_path / "missing.fq"))


# ---------------------------------------------------------------------------------------------------------------- .maln
def parse_maln(body):
    """the reference's file (after line 1) -> header dict + AlnSeq dicts, in file order"""
    L = body.split("\n")
    h = dict(nas=int(L[0].split()[1]), size=int(L[1].split()[1]), coc=int(L[2].split()[1]))
    assert L[3] == "__REFERENCE__"
    h["ref_id"], h["ref_desc"] = L[4][3:], L[5][5:]
    h["len"], h["ref_size"] = int(L[6].split()[1]), int(L[7].split()[1])
    h["seq"] = L[8][4:]
    h["gaps"] = np.array(L[9].split()[1:], np.int32)
    i = L.index("__PSSM__")
    j = L.index("__ALNSEQS__")
    nums = [int(x) for ln in L[i + 1:j] for x in ln.split() if x.lstrip("-").isdigit()]
    assert nums[0] == 15 and len(nums) == 1 + 2 * 775
    h["fpsm"], h["rpsm"] = np.array(nums[1:776], np.int32), np.array(nums[776:], np.int32)
    seqs = []
    k = j + 1
    while k + 12 < len(L):
        a = dict(id=L[k][3:], desc=L[k + 1][5:], score=int(L[k + 2][6:]), num_inputs=int(L[k + 3].split()[1]), start=int(L[k + 4][6:]),
                 end=int(L[k + 5][4:]), rc=int(L[k + 6][3:]), tr=int(L[k + 7][3:]), dr=int(L[k + 8][3:]), seg=L[k + 9][4:], seq=L[k + 10][4:],
                 smp=L[k + 11][4:])
        t = L[k + 12].split()[1:]
        a["ins"] = {int(t[x]): t[x + 1] for x in range(0, len(t), 2)}
        seqs.append(a)
        k += 13
    assert len(seqs) == h["nas"]
    return h, seqs


def reads_from_alnseqs(h, seqs):
    """undo merge / split: one read per 'a' AlnSeq or per (f, b) pair -> the per-read arrays the device returns"""
    backs = {a["id"][:-2]: a for a in seqs if a["seg"] == "b"}
    reads = []
    for a in seqs:
        if a["seg"] == "b":
            continue
        segs = [a] + ([backs[a["id"][:-2]]] if a["seg"] == "f" else [])
        bases, runs = [], []

        def run(t, n):
            if runs and runs[-1][0] == t:
                runs[-1][1] += n
            else:
                runs.append([t, n])
        for s in segs:
            for c, ch in enumerate(s["seq"]):
                if c in s["ins"]:
                    bases.append(s["ins"][c]); run(1, len(s["ins"][c]))
                if ch == "-":
                    run(2, 1)
                else:
                    bases.append(ch); run(0, 1)
        rid = a["id"][:-2] if a["seg"] == "f" else a["id"]
        ae = a["end"] if a["seg"] == "a" else h["len"] + segs[1]["end"]
        reads.append(dict(id=rid, desc=a["desc"], bases="".join(bases), runs=runs, score=a["score"], as_=a["start"], ae=ae, rc=a["rc"],
                          df=a["dr"], db=segs[-1]["dr"], tr=a["tr"], ni=a["num_inputs"]))
    import re
    if reads and all(re.fullmatch(r"r\d{4}.*", r["id"]) for r in reads):
        reads.sort(key=lambda r: int(r["id"][1:5]))       # the synthetic sessions: FSDB order = input order = id order (ties of the sort)
    return reads


def write_from_reads(api, h, reads, path, circular):
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r["bases"]) for r in reads], out=off[1:])
    ids = b"".join(r["id"].encode() + b"\0" for r in reads)
    descs = b"".join(r["desc"].encode() + b"\0" for r in reads)
    id_off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r["id"]) + 1 for r in reads], out=id_off[1:])
    desc_off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r["desc"]) + 1 for r in reads], out=desc_off[1:])
    run_off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r["runs"]) for r in reads], out=run_off[1:])
    packed = np.array([(t << 14) | n for r in reads for t, n in r["runs"]], np.uint16)
    g = lambda k, dt: np.array([r[k] for r in reads], dt)
    rd = dict(bases=np.frombuffer("".join(r["bases"] for r in reads).encode(), np.uint8), offsets=off, ids=ids, id_off=id_off, descs=descs,
              desc_off=desc_off, rc=g("rc", np.uint8), trimmed=g("tr", np.uint8), num_inputs=g("ni", np.int32), score=g("score", np.int32),
              as_=g("as_", np.int32), ae=g("ae", np.int32), abr=np.zeros(len(reads), np.int32), run_off=run_off, packed=packed,
              dropped_front=g("df", np.uint8), dropped_back=g("db", np.uint8))
    return api.write_maln(path, h["ref_id"], h["ref_desc"], h["seq"], circular, h["size"], h["coc"], h["gaps"], h["fpsm"], h["rpsm"], rd)


# circ_k10_SN holds a read that starts beyond seq_len in rounds 2 and 3 (front AlnSeq of length -2, back smp two longer than its seq)
@pytest.mark.parametrize("name", ["circ_k10", "lin_pe", "tr1_tf_lin", "tr1_tf_c", "circ_k10_H", "circ_k10_SN"])
def test_write_maln_reproduces_reference_files(api, gold, name, tmp_path):
    s = gold["sessions"][name]
    circular = int("-c" in s["flags"])
    for it, body in enumerate(s["malns"]):
        h, seqs = parse_maln(body)
        assert h["ref_size"] == api.maln_ref_size(h["len"], circular)
        reads = reads_from_alnseqs(h, seqs)
        p = str(tmp_path / f"out.{it + 1}")
        n = write_from_reads(api, h, reads, p, circular)
        got = open(p).read()
        assert got.startswith("/* map_alignment [V1.0] */ ")
        got = got.split("\n", 1)[1]
        assert n == h["nas"]
        if got != body:
            ga, gb = got.split("\n"), body.split("\n")
            for k, (x, y) in enumerate(zip(ga, gb)):
                assert x == y, f"{name} iteration {it + 1}: line {k + 2}: {x[:120]!r} != {y[:120]!r}"
            assert len(ga) == len(gb)


def test_write_maln_rejects_bad_input(api, tmp_path):
    with pytest.raises(api.MiaGpuError):
        api.write_maln(str(tmp_path / "nodir" / "x"), "r", "", "ACGT", 0, 0, 1, np.zeros(4, np.int32), np.zeros(775, np.int32), np.zeros(775, np.int32),
                       dict(bases=np.zeros(0, np.uint8), offsets=np.zeros(1, np.int64), ids=b"", id_off=np.zeros(1, np.int64), rc=np.zeros(0, np.uint8),
                            score=np.zeros(0, np.int32), as_=np.zeros(0, np.int32), ae=np.zeros(0, np.int32), abr=np.zeros(0, np.int32),
                            run_off=np.zeros(1, np.int64), packed=np.zeros(0, np.uint16)))


def test_write_maln_threads_do_not_change_the_file(api, gold, tmp_path, monkeypatch):
    # 30 x the golden reads (about 12,000 AlnSeqs): worker threads format ranges of the sorted list; the file must not depend on them
    s = gold["sessions"]["circ_k10"]
    h, seqs = parse_maln(s["malns"][1])
    reads = reads_from_alnseqs(h, seqs)
    big = [dict(r, id=f"{r['id']}x{rep}") for rep in range(30) for r in reads]
    outs = []
    for threads in ("1", "3", "16"):
        monkeypatch.setenv("MIAGPU_MALN_THREADS", threads)
        p = str(tmp_path / f"t{threads}")
        n = write_from_reads(api, h, big, p, 1)
        assert n == 30 * h["nas"]
        outs.append(open(p).read().split("\n", 1)[1])
    assert outs[0] == outs[1] == outs[2]
    monkeypatch.delenv("MIAGPU_MALN_THREADS")
    p = str(tmp_path / "auto")
    write_from_reads(api, h, big, p, 1)
    assert open(p).read().split("\n", 1)[1] == outs[0]


def _ref_records(ref, text, tmp_path):
    """the reference's own reader loop (oracle/ref_harness.c: refh_read_seqs) on `text`"""
    import ctypes as C
    fi, fo = tmp_path / "in.txt", tmp_path / "out.txt"
    fi.write_bytes(text.encode("latin-1"))
    ref.lib.refh_read_seqs.restype = C.c_longlong
    ref.lib.refh_read_seqs.argtypes = [C.c_char_p, C.c_char_p]
    n = ref.lib.refh_read_seqs(str(fi).encode(), str(fo).encode())
    recs = []
    for line in fo.read_bytes().decode("latin-1").split("\n")[:-1]:
        rid, rest = line.split("\t", 1)
        desc, seq, qs = rest.rsplit("\t", 2)
        recs.append([rid, desc, seq, int(qs)])
    assert n == len(recs)
    return recs


def test_fastx_reader_fuzz_against_reference_reader(api, ref, tmp_path, parse_mode):
    # 400 seeded texts assembled from record-shaped pieces and noise: whatever the reference's reader makes of them, ours must too
    import random
    rng = random.Random(2026)
    alpha = "ACGTNacgtn"

    def word(n, chars):
        return "".join(rng.choice(chars) for _ in range(n))

    def record(fastq):
        rid = word(rng.choice([1, 5, 20, 99, 100, 101, 140]), "abcXYZ019_-.:")
        desc = rng.choice(["", " ", " d", "\tx y", "  two  blanks ", " " + word(rng.choice([10, 127, 128, 129, 200]), "abc def")])
        n = rng.choice([0, 1, 30, 75, 255, 256, 257, 300])
        seq = word(n, alpha)
        if not fastq:
            w = rng.choice([60, 80, 1000])
            body = "\n".join(seq[i:i + w] for i in range(0, len(seq), w))
            return f">{rid}{desc}\n{body}" + rng.choice(["\n", "\n\n", ""])
        qual = word(n if rng.random() < 0.93 else max(0, n - 1), "!#5AIh~")
        plus = rng.choice(["+", "+", "+" + rid, "-"])
        return f"@{rid}{desc}\n{seq}\n{plus}\n{qual}" + rng.choice(["\n", "\n", "\n\n", ""])
    checked = 0
    for t in range(400):
        fastq = rng.random() < 0.6
        text = "".join(record(fastq) for _ in range(rng.randint(0, 6)))
        if rng.random() < 0.2:
            k = rng.randint(0, len(text))
            text = text[:k]                                  # cut anywhere
        if rng.random() < 0.1:
            text = rng.choice(["\n", " ", "x", ">", "@"]) + text
        want = _ref_records(ref, text, tmp_path)
        got = _records(api, text, 1 << 20 if parse_mode == "parallel" else rng.choice([1, 2, 1 << 20]))
        # the dump splits on tabs: a record whose id / sequence holds a tab cannot be told apart there -- none is generated
        assert got == want, (t, text[:300], got[:3], want[:3])
        checked += len(want)
    assert checked > 500


def test_fastx_parallel_parse_equals_sequential_on_a_large_file(api, tmp_path, monkeypatch):
    # 60,000 well-formed records (7 MB): every piece is kept, and the merged batch equals the one-cursor parse array for array
    import _pkg
    _pkg.load()
    from mia_b200 import synth
    ref = synth.random_reference(5000, seed=5)
    bases, off, _ = synth.make_reads(ref, 60000, 35, 75, seed=6)
    p = tmp_path / "big.fq"
    p.write_bytes(synth.fastq_text(bases, off))
    out = []
    for threads in ("1", "7"):
        monkeypatch.setenv("MIAGPU_FASTX_PAR_MIN", "1")
        monkeypatch.setenv("MIAGPU_FASTX_THREADS", threads)
        r = api.FastxReader(path=str(p))
        b = r.next(1 << 40)
        assert r.next(1 << 40) is None
        r.close()
        out.append(b)
    a, b = out
    assert a["n"] == b["n"] == 60000
    for k in ("bases", "offsets", "id_off", "desc_off", "qual_sum"):
        assert (a[k] == b[k]).all(), k
    assert a["ids"] == b["ids"] and a["descs"] == b["descs"]
    assert (a["bases"] == bases).all() and (a["offsets"] == off).all()
