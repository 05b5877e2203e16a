"""Generates tests/golden/maln_session_r2.json.gz: the UNMODIFIED reference binary (oracle/_ref/mia) on the inputs of
tests/golden/sessions_r2.json.gz (make_golden_r2.py) -- reads that score exactly 2000, split patterns that change, -D -- keeping
every iteration's `.maln` file without its first line (a time stamp).  These files show the reference's FragSeq -> AlnSeq
pointer behaviour in full: an AlnSeq that stale pointers reach is written once per pointer.

    python tests/golden/make_maln_golden_r2.py"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import synth  # noqa: E402
from oracle.pyoracle import Ref  # noqa: E402

MIA = os.path.join(ROOT, "oracle", "_ref", "mia")
CASES = [("flat_2000_c", "flat", ["-c", "-i"]), ("origin305_splitflip_c", "onepass", ["-c", "-i"]),
         ("synth3k_div10_c_k12_D", "ancient", ["-c", "-k", "12", "-i", "-D"]), ("synth1k_N_lin_D", "ancient", ["-i", "-D"])]


def main():
    r = Ref()
    mats = {"flat": r.flat_pssm(), "onepass": r.read_pssm("/root/reference/matrices/ancient.submat.solexa.onepass.txt"),
            "ancient": r.read_pssm("/root/reference/matrices/ancient.submat.txt")}
    sess = json.load(gzip.open(os.path.join(HERE, "sessions_r2.json.gz"), "rt"))
    out = {}
    for name, matrix, flags in CASES:
        s = sess[name]
        fq = "".join(f"@r{i}\n{rd}\n+\n{'I' * len(rd)}\n" for i, rd in enumerate(s["reads"]))
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "ref.fa"), "w").write(">ref\n" + s["ref"] + "\n")
            open(os.path.join(d, "reads.fq"), "w").write(fq)
            open(os.path.join(d, "m.txt"), "w").write(synth.matrix_text(mats[matrix]))
            subprocess.run([MIA, "-r", "ref.fa", "-f", "reads.fq", "-s", "m.txt", "-m", "out"] + flags, cwd=d, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            malns, i = [], 1
            while os.path.exists(os.path.join(d, f"out.{i}")):
                malns.append(open(os.path.join(d, f"out.{i}")).read().split("\n", 1)[1])
                i += 1
        assert len(malns) == len(s["iters"]), (name, len(malns), len(s["iters"]))
        out[name] = dict(ref_text=">ref\n" + s["ref"] + "\n", fastq=fq, matrix=matrix, flags=flags, malns=malns)
        print(name, "iterations", len(malns), "bytes", [len(m) for m in malns])
    with gzip.open(os.path.join(HERE, "maln_session_r2.json.gz"), "wt", compresslevel=9) as f:
        json.dump(out, f)
    print("written", os.path.getsize(os.path.join(HERE, "maln_session_r2.json.gz")), "bytes")


if __name__ == "__main__":
    main()
