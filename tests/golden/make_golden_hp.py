"""Generates tests/golden/hp.json.gz by running the UNMODIFIED reference (oracle/_ref) with -h, the homopolymer-discounted gap
candidates of dyn_prog (mia.c:882-905, hp_discount_penalty mia.c:1096-1134, pop_hpl_and_hps mia.c:1193-1234):

  align     250 alignments (reference dyn_prog + max_sg_score + traceback with hp_special = 1) of homopolymer-rich pairs
  sessions  whole `mia -h` sessions (library form: per read pass 1, per iteration reads / AlnSeq list / consensus) and the `.maln`
            files the reference BINARY writes for the same inputs:  hp2k_c_k10_h (-c -k 10 -h), hp2k_lin_k12_hD (-k 12 -h -D)

One subprocess per session (the reference library does not survive several sessions in one process).
    python tests/golden/make_golden_hp.py"""
import gzip
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle.pyoracle import Ref  # noqa: E402
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import synth  # noqa: E402
from hp_data import hp_reference, hp_reads  # noqa: E402
from make_golden_r2 import session as _unused  # noqa: E402,F401  (same record layout)

MIA = os.path.join(ROOT, "oracle", "_ref", "mia")
MATS = {"onepass": "/root/reference/matrices/ancient.submat.solexa.onepass.txt", "ancient": "/root/reference/matrices/ancient.submat.txt"}
SESSIONS = {"hp2k_c_k10_h": dict(matrix="onepass", circular=1, k=10, distant=0, div=0.02, indel=0.01, hi=110, flags=["-c", "-k", "10", "-i", "-h"]),
            "hp2k_lin_k12_hD": dict(matrix="ancient", circular=0, k=12, distant=1, div=0.04, indel=0.003, hi=80, flags=["-k", "12", "-i", "-h", "-D"])}


def fuzz_pair(rng):
    runs = [(rng.choice("ACGT"), max(1, int(rng.expovariate(0.45)))) for _ in range(rng.randint(4, 45))]
    ref = "".join(b * min(n, 9) for b, n in runs)
    a = rng.randint(0, max(0, len(runs) - 3))
    rd = []
    for b, n in runs[a:a + rng.randint(2, 25)]:
        x = rng.random()
        if x < 0.25:
            n = max(1, n + rng.choice((-2, -1, 1, 2)))
        elif x < 0.30:
            b = rng.choice("ACGT")
        rd.append(b * min(n, 9))
    return ref, "".join(rd)[:250] or "A"


def session_inputs(name):
    c = SESSIONS[name]
    ref = hp_reference(2200, 31)
    sample = synth.diverge(ref, c["div"], seed=3, indel_rate=c["indel"])
    reads, _ = hp_reads(sample, 300, 41, lo=35, hi=c["hi"], circular=bool(c["circular"]))
    return ref, reads


def run_session(name):
    c = SESSIONS[name]
    r = Ref()
    sm = r.read_pssm(MATS[c["matrix"]])
    ref, reads = session_inputs(name)
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(">ref\n" + ref + "\n")
        path = f.name
    s = r.sess_new(path, c["circular"], sm, k=c["k"], soft_mask=0, distant_ref=c["distant"], hp=1)
    p1 = []
    for i, rd in enumerate(reads):
        d = r.sess_pass1(s, "r%d" % i, rd)
        p1.append({k2: d[k2] for k2 in ("hits", "added", "score", "rc", "as_", "ae", "strand_known", "fw_score", "rc_score", "start", "end", "split")})
    r.sess_end_pass1(s)
    iters = []
    for _ in range(30):
        cons, conv = r.sess_iterate(s, sort=0)
        rd = r.sess_reads(s)
        slots = r.sess_slots(s)
        iters.append(dict(cons=cons, converged=conv, reads=[[x["score"], x["as_"], x["ae"], x["rc"], x["strand_known"]] for x in rd],
                          ids=[int(x["id"][1:]) for x in rd],
                          slots=[[x["id"], x["start"], x["end"], x["dropped"], x["segment"], x["seq"], x["smp"], x["ins"]] for x in slots],
                          gaps=np.flatnonzero(r.sess_gaps(s)).tolist()))
        if conv:
            break
    os.unlink(path)
    # the binary on the same inputs
    fq = "".join(f"@r{i}\n{rd}\n+\n{'I' * len(rd)}\n" for i, rd in enumerate(reads))
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "ref.fa"), "w").write(">ref\n" + ref + "\n")
        open(os.path.join(d, "reads.fq"), "w").write(fq)
        open(os.path.join(d, "m.txt"), "w").write(synth.matrix_text(sm))
        subprocess.run([MIA, "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + c["flags"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        malns, i = [], 1
        while os.path.exists(os.path.join(d, f"out.{i}")):
            malns.append(open(os.path.join(d, f"out.{i}")).read().split("\n", 1)[1])
            i += 1
    assert len(malns) == len(iters), (name, len(malns), len(iters))
    complete = len(iters) <= 8
    iters, malns = iters[:8], malns[:8]              # (the -h -D session oscillates until MAX_ITER: the first eight rounds are kept)
    return dict(ref=ref, reads=reads, circular=c["circular"], k=c["k"], distant_ref=c["distant"], matrix=c["matrix"], flags=c["flags"],
                pass1=p1, iters=iters, ref_text=">ref\n" + ref + "\n", fastq=fq, malns=malns, complete=complete)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--session":
        json.dump(run_session(sys.argv[2]), open(sys.argv[3], "w"))
        return
    r = Ref()
    rng = random.Random(77)
    mats = {k: r.read_pssm(v) for k, v in MATS.items()}
    mats["flat"] = r.flat_pssm()
    align = []
    for it in range(250):
        s1, s2 = fuzz_pair(rng)
        m = rng.choice(sorted(mats))
        sg5 = rng.randint(0, 1)
        a = r.align(s1, s2, mats[m], sg5, None, hp=1)
        align.append(dict(ref=s1, read=s2, matrix=m, sg5=sg5, out=[a["score"], a["abr"], a["abc"], a["aer"], a["aec"], a["ref_gapped"], a["read_gapped"]]))
    r.set_hp(0)
    out = dict(align=align, sessions={})
    for name in SESSIONS:
        with tempfile.NamedTemporaryFile(suffix=".json", delete=False) as f:
            tmp = f.name
        subprocess.run([sys.executable, os.path.abspath(__file__), "--session", name, tmp], check=True)
        out["sessions"][name] = json.load(open(tmp))
        os.unlink(tmp)
        s = out["sessions"][name]
        print(name, "fsdb", sum(p["added"] for p in s["pass1"]), "iterations", len(s["iters"]), "maln bytes", [len(m) for m in s["malns"]])
    with gzip.open(os.path.join(HERE, "hp.json.gz"), "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("written", os.path.getsize(os.path.join(HERE, "hp.json.gz")), "bytes")


if __name__ == "__main__":
    main()
