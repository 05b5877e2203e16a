"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run here, in the build container; the GPU box only
reads the committed outputs.   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402

REFROOT = "/root/reference"


def main():
    r = Ref()
    m = {"ancient": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.txt"),
         "onepass": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.solexa.onepass.txt"),
         "pe": r.read_pssm(f"{REFROOT}/matrices/ancient.submat.solexa.pe.txt"),
         "flat": r.flat_pssm()}
    for k in list(m):
        m[k + "_rc"] = r.revcom_pssm(m[k])
    np.savez_compressed(os.path.join(HERE, "pssm.npz"), **m)
    print("wrote pssm.npz")


if __name__ == "__main__":
    main()
