"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run here, in the build container; the GPU box only
reads the committed outputs.   python tests/golden/make_golden.py"""
import json
import os
import random
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import synth  # noqa: E402

REFROOT = "/root/reference"


def fuzz_align_cases(n, seed):
    rng = random.Random(seed)
    cases = []
    for _ in range(n):
        n1 = rng.randint(5, 200)
        ref = "".join(rng.choice("ACGT" if rng.random() < 0.97 else "N") for _ in range(n1))
        n2 = rng.randint(1, 90)
        if rng.random() < 0.7 and n1 > n2:
            s = rng.randint(0, n1 - n2)
            rd = list(ref[s:s + n2])
            for i in range(len(rd)):
                x = rng.random()
                if x < 0.05:
                    rd[i] = rng.choice("ACGT")
                elif x < 0.07:
                    rd[i] = ""
                elif x < 0.09:
                    rd[i] = rd[i] + rng.choice("ACGT")
                elif x < 0.095:
                    rd[i] = "N"
            rd = "".join(rd)[:256] or "A"
        else:
            rd = "".join(rng.choice("ACGT") for _ in range(n2))
        mask = None
        if rng.random() < 0.5:
            mask = [0] * n1
            for _ in range(rng.randint(1, 3)):
                a = rng.randint(0, n1 - 1)
                b = rng.randint(a, min(n1 - 1, a + rng.randint(0, 90)))
                for i in range(a, b + 1):
                    mask[i] = 1
        cases.append(dict(ref=ref, read=rd, mask=mask, sg5=rng.randint(0, 1), mat=rng.choice(["flat", "ancient", "ancient_rc", "onepass"])))
    return cases


def read_fixture_reads(path):
    reads, ids = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            ids.append(line[1:].split()[0])
            reads.append("")
        elif reads:
            reads[-1] += line.upper()
    return ids, [r[:256] for r in reads]


def session(r, ref, reads, sm, circular, k, soft_mask, max_iter=30, repeat_filt=0, just_outer_coords=1, qual_sums=None):
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(">ref\n" + ref + "\n")
        path = f.name
    s = r.sess_new(path, circular, sm, k=k, soft_mask=soft_mask)
    if repeat_filt:
        r.sess_set_repeat(s, repeat_filt, just_outer_coords)
    p1 = []
    for i, rd in enumerate(reads):
        if not rd:
            p1.append(None)
            continue
        d = r.sess_pass1(s, "r%d" % i, rd, qual_sum=0 if qual_sums is None else qual_sums[i])
        p1.append({k2: d[k2] for k2 in ("hits", "added", "score", "rc", "as_", "ae", "strand_known", "fw_score", "rc_score",
                                       "start", "end", "split", "b_start", "b_end", "f_ref", "f_frag", "b_ref", "b_frag")})
    r.sess_end_pass1(s)
    iters = []
    for _ in range(max_iter):
        cons, conv = r.sess_iterate(s, sort=0)
        rd = r.sess_reads(s)
        slots = r.sess_slots(s)
        iters.append(dict(cons=cons, converged=conv, reads=[[x["score"], x["as_"], x["ae"], x["rc"]] for x in rd],
                          ids=[int(x["id"][1:]) for x in rd], unique=[x["unique_best"] for x in rd],
                          slots=[[x["start"], x["end"], x["dropped"], x["segment"], x["seq"], x["smp"], x["ins"]] for x in slots],
                          gaps=np.flatnonzero(r.sess_gaps(s)).tolist()))
        if conv:
            break
    os.unlink(path)
    return dict(ref=ref, reads=reads, circular=circular, k=k, soft_mask=soft_mask, pass1=p1, iters=iters, repeat_filt=repeat_filt,
                just_outer_coords=just_outer_coords, **({} if qual_sums is None else {"qual_sums": [int(q) for q in qual_sums]}))


def main():
    r = Ref()
    m = {"ancient": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.txt"),
         "onepass": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.solexa.onepass.txt"),
         "pe": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.solexa.pe.txt"),
         "flat": r.flat_pssm()}
    for k in list(m):
        m[k + "_rc"] = r.revcom_pssm(m[k])
    np.savez_compressed(os.path.join(HERE, "pssm.npz"), **m)
    # matrix text of one file, to pin the oracle's parser
    open(os.path.join(HERE, "onepass_matrix_fixture.json"), "w").write(json.dumps(
        {"text": open(f"{REFROOT}/matrices/ancient.submat.solexa.onepass.txt").read()}))

    # a5-a7: the reference's answers on fuzzed alignments
    cases = fuzz_align_cases(400, seed=17)
    for c in cases:
        mask = None if c["mask"] is None else np.array(c["mask"], np.uint8)
        a = r.align(c["ref"], c["read"], m[c["mat"]], c["sg5"], mask)
        c["out"] = [a["score"], a["abr"], a["abc"], a["aer"], a["aec"], a["ref_gapped"], a["read_gapped"]]
    json.dump(cases, open(os.path.join(HERE, "align_cases.json"), "w"))

    # sessions: the reference's own fixtures + a small synthetic circular case
    tr1 = "".join(open(f"{REFROOT}/test/tr1.fna").read().split("\n")[1:])
    _, tf = read_fixture_reads(f"{REFROOT}/test/tf.fna")
    sess = {"tr1_tf_c": session(r, tr1, tf, m["ancient"], 1, 0, 0),
            "tr1_tf_lin": session(r, tr1, tf, m["ancient"], 0, 0, 0),
            "tr1_tf_c_k8_M": session(r, tr1, tf, m["ancient"], 1, 8, 1)}
    ref = synth.random_reference(2000, seed=1)
    g = synth.diverge(ref, 0.02, seed=3, indel_rate=0.003)
    b, off, _ = synth.make_reads(g, 250, 35, 75, seed=4)
    reads = [synth.read_str(b, off, i) for i in range(250)]
    sess["synth2k_c_k10"] = session(r, ref, reads, m["onepass"], 1, 10, 0)
    # BASELINE configs[3] in small: k-mer filter + a starting reference 10 % (+ indels) away from the sample -> many rounds
    ref = synth.random_reference(3000, seed=11)
    g = synth.diverge(ref, 0.10, seed=12, indel_rate=0.005)
    b, off, _ = synth.make_reads(g, 500, 35, 75, seed=13)
    reads = [synth.read_str(b, off, i) for i in range(500)]
    sess["synth3k_div10_c_k12"] = session(r, ref, reads, m["ancient"], 1, 12, 0)
    # BASELINE configs[2] in small: merged paired-end reads (30-140 bp), ancient.submat.solexa.pe, to convergence
    ref = synth.random_reference(2500, seed=21)
    g = synth.diverge(ref, 0.03, seed=22, indel_rate=0.004)
    b, off, _ = synth.make_reads(g, 300, 30, 140, seed=23)
    reads = [synth.read_str(b, off, i) for i in range(300)]
    sess["synth2k5_pe_long_c_k12"] = session(r, ref, reads, m["pe"], 1, 12, 0)
    # -u (repeat filter): a small genome at high coverage with PCR-style duplicates, some of them with a sequencing error
    ref = synth.random_reference(1500, seed=31)
    g = synth.diverge(ref, 0.03, seed=32, indel_rate=0.003)
    b, off, _ = synth.make_reads(g, 260, 35, 75, seed=33)
    reads = [synth.read_str(b, off, i) for i in range(260)]
    rng = np.random.default_rng(34)
    for _ in range(140):                                   # duplicates: same molecule, sometimes one base changed
        rd = reads[int(rng.integers(0, 260))]
        if rng.random() < 0.4:
            q = int(rng.integers(0, len(rd)))
            rd = rd[:q] + "ACGT"[("ACGT".index(rd[q]) + 1) % 4 if rd[q] in "ACGT" else 0] + rd[q + 1:]
        reads.append(rd)
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    sess["synth1k5_dups_c_k10_u"] = session(r, ref, reads, m["onepass"], 1, 10, 0, repeat_filt=1)
    sess["synth1k5_dups_lin_k10_uA"] = session(r, ref, reads, m["onepass"], 0, 10, 0, repeat_filt=1, just_outer_coords=0)
    # -U: the same reads with FASTQ quality sums (read_fastq: sum(q - 33)); duplicates are decided by quality, not by score
    qs = [int(v) for v in np.random.default_rng(35).integers(20, 41, len(reads)) * np.array([len(x) for x in reads])]
    sess["synth1k5_dups_c_k10_U"] = session(r, ref, reads, m["onepass"], 1, 10, 0, repeat_filt=2, qual_sums=qs)
    json.dump(sess, open(os.path.join(HERE, "sessions.json"), "w"))
    # f1: the reference's sort_fsdb[_qscore] + set_uniq_in_fsdb on small FSDBs full of ties
    rng = np.random.default_rng(7)
    rep = []
    for n in (1, 2, 9, 60, 400):
        for use_q in (0, 1):
            for jo in (0, 1):
                for tol in (0, 2):
                    rc = rng.integers(0, 2, n); as_ = rng.integers(0, 12, n); ae = as_ + rng.integers(30, 34, n)
                    k4 = rng.integers(2000, 2004, n); tr = rng.integers(0, 2, n)
                    order, uniq = r.repeat_filter(rc, as_, ae, k4, tr, jo, tol, use_qscore=use_q)
                    rep.append(dict(rc=rc.tolist(), as_=as_.tolist(), ae=ae.tolist(), key4=k4.tolist(), trimmed=tr.tolist(), use_qscore=use_q,
                                    just_outer=jo, tolerance=tol, order=order.tolist(), unique=uniq.tolist()))
    json.dump(rep, open(os.path.join(HERE, "repeat_cases.json"), "w"))
    # f4: the reference's trim_frag on reads with (damaged) adapter prefixes at the 3' end
    import random as _r
    rg = _r.Random(11)
    trims = []
    for adapter in ("GTCAGACACGCAACAGGGGATAGGCAAGGCACACAGGGGATAGG", "ACGTTGCA"):
        for _ in range(150):
            L = rg.randint(1, 90)
            rd = "".join(rg.choice("ACGT") for _ in range(L))
            if rg.random() < 0.7:
                frag = list(adapter[: rg.randint(1, len(adapter))])
                for i in range(len(frag)):
                    y = rg.random()
                    if y < 0.06:
                        frag[i] = rg.choice("ACGTN")
                    elif y < 0.08:
                        frag[i] = ""
                rd = rd + "".join(frag)
            t = r.trim(rd, adapter)
            trims.append(dict(read=rd, adapter=adapter, out=[t[k] for k in ("trimmed", "trim_point", "score", "abr", "abc", "aer")]))
    json.dump(trims, open(os.path.join(HERE, "trim_cases.json"), "w"))
    print("wrote pssm.npz, align_cases.json, sessions.json, repeat_cases.json, trim_cases.json")


if __name__ == "__main__":
    main()
