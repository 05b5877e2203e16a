"""-m gpu: the plain-C host (host/mia_gpu.c, gcc, linked against libmiagpu.so only) run as a program on the inputs the
UNMODIFIED reference binary was run on (tests/golden/maln_session.json.gz): same FASTA, FASTQ, matrix file and flags ->
the same number of `.maln` files, each byte-identical after line 1."""
import gzip
import json
import os
import subprocess

import pytest

from test_host_c import HOST, ensure_host, matrix_text

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name,matrix", [("circ_k10", "ancient"), ("lin_pe", "pe"), ("tr1_tf_lin", "ancient"), ("tr1_tf_lin_k8", "ancient"),
                                         ("tr1_tf_c", "ancient"), ("dups_c_k10_u", "onepass"), ("dups_c_k10_U", "onepass"),
                                         ("circ_k10_H", "ancient"), ("circ_k10_SN", "ancient"), ("circ_k10_p2", "ancient")])
def test_c_host_writes_the_reference_maln_files(golden, name, matrix, tmp_path):
    s = json.load(gzip.open(os.path.join(HERE, "golden", "maln_session.json.gz"), "rt"))["sessions"][name]
    # tr1_tf_*: the reference's own fixtures test/tr1.fna + test/tf.fna (FASTA reads, a lower-case stretch, a 236-base read)
    (tmp_path / "ref.fa").write_text(s.get("ref_text") or f">{s['ref_id']} {s['ref_desc']}\n{s['ref']}\n")
    (tmp_path / "reads.fq").write_text(s["fastq"])
    (tmp_path / "m.txt").write_text(matrix_text(golden[matrix]))
    ensure_host()
    r = subprocess.run([HOST, "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + s["flags"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for it, body in enumerate(s["malns"]):
        got = open(tmp_path / f"out.{it + 1}").read().split("\n", 1)[1]
        if got != body:
            for ln, (x, y) in enumerate(zip(got.split("\n"), body.split("\n"))):
                assert x == y, f"{name} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(got) == len(body)
    assert not os.path.exists(tmp_path / f"out.{len(s['malns']) + 1}")
    assert "Assembly convergence" in r.stderr


@pytest.mark.parametrize("name", ["flat_2000_c", "origin305_splitflip_c", "synth3k_div10_c_k12_D", "synth1k_N_lin_D"])
def test_c_host_follows_the_reference_pointers(golden, name, tmp_path):
    # tests/golden/make_maln_golden_r2.py: reads that score exactly 2000 (strand_known = 0), split patterns that change, -D.  The
    # reference's `.maln` files list an AlnSeq once per FragSeq pointer that reaches it -- stale pointers included.
    s = json.load(gzip.open(os.path.join(HERE, "golden", "maln_session_r2.json.gz"), "rt"))[name]
    (tmp_path / "ref.fa").write_text(s["ref_text"])
    (tmp_path / "reads.fq").write_text(s["fastq"])
    (tmp_path / "m.txt").write_text(matrix_text(golden[s["matrix"]]))
    ensure_host()
    r = subprocess.run([HOST, "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + s["flags"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for it, body in enumerate(s["malns"]):
        got = open(tmp_path / f"out.{it + 1}").read().split("\n", 1)[1]
        if got != body:
            for ln, (x, y) in enumerate(zip(got.split("\n"), body.split("\n"))):
                assert x == y, f"{name} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(got) == len(body)
    assert not os.path.exists(tmp_path / f"out.{len(s['malns']) + 1}")


@pytest.mark.parametrize("name", ["hp2k_c_k10_h", "hp2k_lin_k12_hD"])
def test_c_host_homopolymer_discount(golden, name, tmp_path):
    # mia -h / mia -h -D (tests/golden/make_golden_hp.py): FASTQ -> device -> the `.maln` files of the unmodified reference binary
    s = json.load(gzip.open(os.path.join(HERE, "golden", "hp.json.gz"), "rt"))["sessions"][name]
    (tmp_path / "ref.fa").write_text(s["ref_text"])
    (tmp_path / "reads.fq").write_text(s["fastq"])
    (tmp_path / "m.txt").write_text(matrix_text(golden[s["matrix"]]))
    ensure_host()
    r = subprocess.run([HOST, "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + s["flags"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for it, body in enumerate(s["malns"]):
        got = open(tmp_path / f"out.{it + 1}").read().split("\n", 1)[1]
        if got != body:
            for ln, (x, y) in enumerate(zip(got.split("\n"), body.split("\n"))):
                assert x == y, f"{name} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(got) == len(body)
    if s["complete"]:
        assert not os.path.exists(tmp_path / f"out.{len(s['malns']) + 1}")


@pytest.mark.parametrize("gpus", [1, 2, 4])
@pytest.mark.parametrize("name,matrix", [("circ_k10", "ancient"), ("lin_pe", "pe"), ("circ_k10_SN", "ancient")])
def test_c_host_for_the_gpus_of_one_box(golden, name, matrix, gpus, tmp_path):
    # host/mia_gpu_mg.c: one process, one context + one thread per GPU, NCCL between them (ncclCommInitAll), the sharded protocol of
    # include/miagpu.h -> the `.maln` files the unmodified reference binary wrote.  -g 1 runs the same code on a one-GPU box.
    import _pkg
    _pkg.load()
    from mia_b200 import api
    if api.load_library().miagpu_device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    s = json.load(gzip.open(os.path.join(HERE, "golden", "maln_session.json.gz"), "rt"))["sessions"][name]
    (tmp_path / "ref.fa").write_text(s.get("ref_text") or f">{s['ref_id']} {s['ref_desc']}\n{s['ref']}\n")
    (tmp_path / "reads.fq").write_text(s["fastq"])
    (tmp_path / "m.txt").write_text(matrix_text(golden[matrix]))
    ensure_host()
    mg = os.path.join(os.path.dirname(HOST), "mia_gpu_mg")
    if not os.path.exists(mg):
        pytest.skip("host/mia_gpu_mg not built (needs nccl.h / libnccl at build time)")
    r = subprocess.run([mg, "-g", str(gpus), "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + s["flags"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    for it, body in enumerate(s["malns"]):
        got = open(tmp_path / f"out.{it + 1}").read().split("\n", 1)[1]
        if got != body:
            for ln, (x, y) in enumerate(zip(got.split("\n"), body.split("\n"))):
                assert x == y, f"{name} -g {gpus} iteration {it + 1}: line {ln + 2}: {x[:160]!r} != {y[:160]!r}"
            assert len(got) == len(body)
    assert not os.path.exists(tmp_path / f"out.{len(s['malns']) + 1}")
    assert "Assembly convergence" in r.stderr
