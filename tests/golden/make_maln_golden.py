"""Generates tests/golden/maln_session.json.gz by running the UNMODIFIED reference binary
(oracle/_ref/mia, built from /root/reference/src by oracle/Makefile) on a small FASTQ data set and
keeping every iteration's `.maln` file (without line 1, which is a time stamp: map_alignment.c:299).
Run here, in the build container; the GPU box only reads the committed output.

    python tests/golden/make_maln_golden.py

The FASTQ text itself is kept too (with its deliberately awkward records: descriptions, lower case,
an over-long ID, an over-long read, blank-separated IDs) because the reference's read_fastq
(io.c:46-167) is what the streaming parser is pinned against."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import synth  # noqa: E402

MIA = os.path.join(ROOT, "oracle", "_ref", "mia")


def fastq_text(bases, off, seed):
    rng = np.random.default_rng(seed)
    out = []
    n = len(off) - 1
    for i in range(n):
        s = bases[off[i]:off[i + 1]].tobytes().decode()
        rid = f"r{i:04d}"
        head = rid
        x = rng.random()
        if x < 0.10:
            head = rid + " some description  with blanks"
        elif x < 0.13:
            head = rid + "\tlen=%d" % len(s)
        elif x < 0.15:
            s = s.lower()
        elif x < 0.16:
            head = rid + "_" + "x" * 120            # longer than MAX_ID_LEN (100)
        elif x < 0.17:
            s = (s * 6)[:300]                       # longer than INIT_ALN_SEQ_LEN (256)
        q = "".join(chr(33 + int(v)) for v in rng.integers(2, 41, len(s)))
        out.append(f"@{head}\n{s}\n+\n{q}\n")
    return "".join(out)


def run_session(name, ref, fq, flags, matrix, ref_text=None):
    with tempfile.TemporaryDirectory() as d:
        ref_text = ref_text or ">refseq a small circle\n" + ref + "\n"
        open(os.path.join(d, "ref.fa"), "w").write(ref_text)
        open(os.path.join(d, "reads.fq"), "w").write(fq)
        cmd = [MIA, "-r", "ref.fa", "-f", "reads.fq", "-s", matrix, "-m", "out"] + flags
        subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        malns = []
        i = 1
        while os.path.exists(os.path.join(d, f"out.{i}")):
            body = open(os.path.join(d, f"out.{i}")).read().split("\n", 1)[1]
            malns.append(body)
            i += 1
    return dict(name=name, ref=ref, ref_id="refseq", ref_desc="a small circle", ref_text=ref_text, fastq=fq, flags=flags, matrix=matrix,
                malns=malns)


def reader_case(text):
    """the reference's own find_input_type + read_next_seq loop (oracle/ref_harness.c: refh_read_seqs) on `text`"""
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libmia_ref.so"))
    lib.refh_read_seqs.restype = C.c_longlong
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.txt"), os.path.join(d, "out.txt")
        open(fi, "wb").write(text.encode("latin-1"))
        n = lib.refh_read_seqs(fi.encode(), fo.encode())
        recs = []
        for line in open(fo, "rb").read().decode("latin-1").split("\n")[:-1]:
            rid, rest = line.split("\t", 1)
            desc, seq, qs = rest.rsplit("\t", 2)
            recs.append([rid, desc, seq, int(qs)])
        assert n == len(recs), (n, len(recs))
    return dict(text=text, records=recs)


TRICKY = [
    # FASTQ: description with blanks and tabs, lower case, id of exactly 100 / 101 / 130 chars, read of 256 / 257 / 300 bases,
    # blank line inside, missing '+', unequal quality length (ends the input), no trailing newline
    "@a\nACGT\n+\nIIII\n@b desc here\nacgtn\n+b\nIIII!\n@c\t tabbed  desc \nAC GT\n+\nII II\n",
    "@" + "i" * 100 + "\nACGT\n+\nIIII\n@" + "j" * 101 + "\nACGT\n+\nIIII\n@" + "k" * 130 + " d\nACGT\n+\nIIII\n@z\nAC\n+\nII\n",
    "@l256\n" + "ACGT" * 64 + "\n+\n" + "I" * 256 + "\n@l257\n" + "ACGT" * 64 + "A\n+\n" + "I" * 257 + "\n@l300\n" + "ACGTA" * 60 + "\n+\n" + "5" * 300 + "\n@after\nAC\n+\nII\n",
    "@x " + "d" * 140 + "\nACGT\n+\nIIII\n@y\nACGT\n+\nIIII\n",
    "@p\nACGT\n+\nIIII\n@q\nACGT\nIIII\n@r\nACGT\n+\nIIII\n",
    "@p\nACGT\n+\nIIII\n@q\nACGT\n+\nIII\n@r\nACGT\n+\nIIII\n",
    "@p\nACGT\n+\nIIII\n\n@q\nACGT\n+\nIIII\n",
    "@p\nACGT\n+\nIIII\n@q\nACG",
    "@p\nACGT\n+\nIIII\n@q",
    "@only\nACGT\n+\nI#5~",
    # FASTA: multi-line, descriptions (first character repeated by the reference), lower case, long reads, '>' inside, empty
    ">s1\nACGT\nacgt\n>s2 some desc\nAC\nGT\n\n>s3\t x\nNNAC\n",
    ">" + "i" * 100 + "\nACGT\n>" + "j" * 120 + " dd\nACGT\n>l300 long\n" + "ACGTA" * 60 + "\n>after\nAC\n",
    ">e1\n>e2\nAC\n>e3 " + "d" * 200 + "\nACGT\n>e4 \nAC",
    "ACGT\n>s\nAC\n",
    "",
]


def main():
    out = {}
    ref = synth.random_reference(1500, seed=11)
    genome = synth.diverge(ref, 0.03, seed=12, indel_rate=0.006)
    bases, off, _ = synth.make_reads(genome, 400, 35, 75, seed=13, circular=True)
    fq = fastq_text(bases, off, 14)
    out["circ_k10"] = run_session("circ_k10", ref, fq, ["-c", "-k", "10", "-i"], "ancient.submat.txt")
    # the user's score cuts (-H; -S / -N) and the second consensus rule (-p 2) on the same reads
    out["circ_k10_H"] = run_session("circ_k10_H", ref, fq, ["-c", "-k", "10", "-i", "-H", "9000"], "ancient.submat.txt")
    out["circ_k10_SN"] = run_session("circ_k10_SN", ref, fq, ["-c", "-k", "10", "-i", "-S", "190", "-N", "-400"], "ancient.submat.txt")
    out["circ_k10_p2"] = run_session("circ_k10_p2", ref, fq, ["-c", "-k", "10", "-i", "-p", "2"], "ancient.submat.txt")
    bases, off, _ = synth.make_reads(genome, 300, 30, 140, seed=15, circular=False)
    fq = fastq_text(bases, off, 16)
    out["lin_pe"] = run_session("lin_pe", ref, fq, ["-i"], "ancient.submat.solexa.pe.txt")
    # -u / -U: PCR-style duplicates (same molecule, sometimes one base changed), qualities differ per read
    bases, off, _ = synth.make_reads(genome, 220, 35, 75, seed=17, circular=True)
    rng = np.random.default_rng(18)
    rl = [bases[off[i]:off[i + 1]].copy() for i in range(220)]
    for _ in range(130):
        r = rl[int(rng.integers(0, 220))].copy()
        if rng.random() < 0.4:
            q = int(rng.integers(0, len(r)))
            r[q] = b"ACGT"[(b"ACGT".index(bytes([r[q]])) + 1) % 4]
        rl.append(r)
    rl = [rl[i] for i in rng.permutation(len(rl))]
    db = np.concatenate(rl)
    do = np.concatenate([[0], np.cumsum([len(r) for r in rl])]).astype(np.int64)
    fq = fastq_text(db, do, 19)
    out["dups_c_k10_u"] = run_session("dups_c_k10_u", ref, fq, ["-c", "-k", "10", "-i", "-u"], "ancient.submat.solexa.onepass.txt")
    out["dups_c_k10_U"] = run_session("dups_c_k10_U", ref, fq, ["-c", "-k", "10", "-i", "-U"], "ancient.submat.solexa.onepass.txt")
    # the reference's own fixtures (test/tr1.fna, test/tf.fna: FASTA reads, lower-case reference stretch, an over-long read)
    fx = "/root/reference/test"
    tr1, tf = open(os.path.join(fx, "tr1.fna")).read(), open(os.path.join(fx, "tf.fna")).read()
    for name, flags in (("tr1_tf_lin", ["-i"]), ("tr1_tf_lin_k8", ["-k", "8", "-i"]), ("tr1_tf_c", ["-c", "-i"])):
        out[name] = run_session(name, None, tf, flags, "ancient.submat.txt", ref_text=tr1)
    cases = [reader_case(t) for t in TRICKY]
    cases.append(reader_case(tf))
    cases.append(reader_case(out["circ_k10"]["fastq"]))
    print("reader cases:", [len(c["records"]) for c in cases])
    for k, v in out.items():
        print(k, "iterations:", len(v["malns"]), "bytes:", [len(m) for m in v["malns"]])
    with gzip.open(os.path.join(HERE, "maln_session.json.gz"), "wt", compresslevel=9) as f:
        json.dump(dict(sessions=out, reader_cases=cases), f)
    print("written", os.path.getsize(os.path.join(HERE, "maln_session.json.gz")), "bytes")


if __name__ == "__main__":
    main()
