"""Generates tests/golden/sessions_r2.json.gz by running the UNMODIFIED reference (oracle/_ref): whole `mia` sessions that
exercise the FragSeq -> AlnSeq POINTER behaviour of the reference (mia_main.c:120-178, 268-276; mia.c:1614, 1653):

  flat_2000_c            reads that score exactly 2000 in pass 1: accepted with strand_known = 0 (mia.c:1653), never realigned
                         (mia_main.c:178), their pass-1 front_asp / back_asp keep pointing into the slot array for ever
  origin30x_splitflip_c  a small circular genome with indels at its origin and reads that begin / end there: reads flip between
                         wrap-split and whole, slot numbers (and the sticky AlnSeq.dropped flags in the slots) slide, back_asp goes stale
  synth3k_div10_c_k12_D  mia -D on the 10 %-divergent set: every read with a k-mer hit is accepted, strand-unknown reads are
                         re-tried against the whole reference on both strands from iteration 2 on (submat carry-over, H6)
  synth1k_N_lin_D        -D on a linear reference with N stretches (find_alignable_len, mia.c:69-91)

Per iteration: consensus, per-read (score, as, ae, rc, strand_known), the culled AlnSeq list in FSDB order (ids included, so
that aliased slots show), positions with gaps.  Run here, in the build container:  python tests/golden/make_golden_r2.py"""
import gzip
import json
import os
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


def session(r, ref, reads, sm, circular, k, distant_ref=0, max_iter=30):
    with tempfile.NamedTemporaryFile("w", suffix=".fa", delete=False) as f:
        f.write(">ref\n" + ref + "\n")
        path = f.name
    s = r.sess_new(path, circular, sm, k=k, soft_mask=0, distant_ref=distant_ref)
    p1 = []
    for i, rd in enumerate(reads):
        d = r.sess_pass1(s, "r%d" % i, rd)
        p1.append({k2: d[k2] for k2 in ("hits", "added", "score", "rc", "as_", "ae", "strand_known", "fw_score", "rc_score", "start", "end",
                                       "split")})
    r.sess_end_pass1(s)
    iters = []
    for _ in range(max_iter):
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
    return dict(ref=ref, reads=reads, circular=circular, k=k, distant_ref=distant_ref, pass1=p1, iters=iters)


def reads_of(g, n, lo, hi, seed, circular=True):
    b, off, _ = synth.make_reads(g, n, lo, hi, seed=seed, circular=circular)
    return [synth.read_str(b, off, i) for i in range(n)]


def origin_reads(seed):
    """a 600 bp circle whose sample has indels right at the origin, plus reads that begin or end within 3 bases of it: between
    rounds such reads flip between wrap-split and whole (searched: seeds whose reference run is clean and flips 4 to 9 reads)"""
    rng = np.random.default_rng(seed)
    ref = synth.random_reference(600, seed=seed)
    gl = list(synth.diverge(ref, 0.05, seed=seed + 1, indel_rate=0.004))
    for _ in range(2):
        gl.insert(len(gl) - int(rng.integers(0, 6)), "ACGT"[int(rng.integers(0, 4))])
    del gl[int(rng.integers(1, 5))]
    g = "".join(gl)
    reads = reads_of(g, 300, 35, 75, seed=seed + 2)
    Gn = len(g)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    for q in range(80):
        L = int(rng.integers(35, 76))
        if q % 2 == 0:
            st = (Gn + int(rng.integers(-3, 4)) - L) % Gn        # ends near the origin
        else:
            st = int(rng.integers(-3, 4)) % Gn                   # starts near the origin
        rd = "".join(g[(st + i) % Gn] for i in range(L))
        if q % 4 >= 2:
            rd = "".join(comp[c] for c in reversed(rd))
        reads.insert(int(rng.integers(0, len(reads))), rd)
    return ref, reads


def build(name):
    r = Ref()
    m = {"ancient": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.txt"),
         "onepass": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.solexa.onepass.txt"),
         "flat": r.flat_pssm()}
    rng = np.random.default_rng(5)
    if name == "flat_2000_c":
        # score == 2000 with the flat matrix: 10 matches, or 13 matches and one mismatch (13 * 200 - 600)
        ref = synth.random_reference(900, seed=41)
        g = synth.diverge(ref, 0.02, seed=42, indel_rate=0.004)
        reads = reads_of(g, 160, 30, 70, seed=43)
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        for q in range(14):                                    # the special reads, spread over the input (and over both strands)
            p = int(rng.integers(0, len(ref) - 20))
            if q % 2 == 0:
                rd = ref[p:p + 10]
            else:
                rd = list(ref[p:p + 14])
                rd[6] = "ACGT"[("ACGT".index(rd[6]) + 1) % 4]
                rd = "".join(rd)
            if q % 4 >= 2:
                rd = "".join(comp[c] for c in reversed(rd))
            if q in (4, 5):                                    # across the origin of the circular reference
                w = ref[-6:] + ref[:8]
                rd = w[:10] if q == 4 else w[:6] + "ACGT"[("ACGT".index(w[6]) + 1) % 4] + w[7:14]
            reads.insert(int(rng.integers(0, len(reads))), rd)
        return session(r, ref, reads, m["flat"], 1, 0)
    if name in ("origin305_splitflip_c", "origin303_splitflip_c"):
        ref, reads = origin_reads(int(name[6:9]))
        return session(r, ref, reads, m["onepass"], 1, 0, max_iter=8)
    if name == "synth3k_div10_c_k12_D":
        # -D on the divergent set of round 1 (sessions.json: synth3k_div10_c_k12)
        ref = synth.random_reference(3000, seed=11)
        g = synth.diverge(ref, 0.10, seed=12, indel_rate=0.005)
        return session(r, ref, reads_of(g, 500, 35, 75, seed=13), m["ancient"], 1, 12, distant_ref=1)
    if name == "synth1k_N_lin_D":
        # -D, linear, reference with N stretches, no k-mer filter, short reads that start out strand-unknown
        ref = list(synth.random_reference(1000, seed=61))
        for a, b in ((120, 150), (480, 500), (700, 760)):
            ref[a:b] = "N" * (b - a)
        g = synth.diverge(synth.random_reference(1000, seed=61), 0.06, seed=62, indel_rate=0.004)
        return session(r, "".join(ref), reads_of(g, 260, 18, 60, seed=63, circular=False), m["ancient"], 0, 0, distant_ref=1)
    raise KeyError(name)


NAMES = ["flat_2000_c", "origin305_splitflip_c", "origin303_splitflip_c", "synth3k_div10_c_k12_D", "synth1k_N_lin_D"]


def main():
    import subprocess
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        json.dump(build(sys.argv[2]), sys.stdout)
        return
    sess = {}
    for name in NAMES:                                         # one process per session: the reference never frees and now and then
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], check=True, capture_output=True)   # corrupts its heap
        sess[name] = json.loads(out.stdout)
    for name, s in sess.items():
        unk = [sum(1 for x in it["reads"] if not x[4]) for it in s["iters"]]
        print(name, "reads in FSDB", len(s["iters"][0]["reads"]), "iterations", len(s["iters"]), "strand-unknown per iteration", unk,
              "AlnSeqs", [len(it["slots"]) for it in s["iters"]])
    with gzip.open(os.path.join(HERE, "sessions_r2.json.gz"), "wt") as f:
        json.dump(sess, f)
    print("wrote sessions_r2.json.gz")


if __name__ == "__main__":
    main()
