"""pass 1 alone on a BASELINE shape (for launch lists): python scripts/gpu_pass1_shape.py c3 1250000 [repeats]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import api, driver, synth  # noqa: E402
import gpu_checks  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "scripts"))
from gpu_assembly import SHAPES  # noqa: E402


def main():
    sh = SHAPES[sys.argv[1]]
    n = int(sys.argv[2])
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ref = synth.random_reference(sh["ref_len"], seed=1 if sh["ref_len"] < 100000 else 321)
    genome = synth.diverge(ref, sh["div"], seed=3, indel_rate=sh["indel"])
    sm = gpu_checks.load_pssm(sh["matrix"])
    b, o = synth.make_reads(genome, n, sh["lens"][0], sh["lens"][1], seed=1000, circular=bool(sh["circular"]))[:2]
    g = api.MiaGpu(0)
    for r in range(reps):
        A = driver.ResidentAssembler(g, ref, sm, circular=sh["circular"], k=sh["k"], strand_unknown="drop", pointer_state=False)
        t0 = time.perf_counter()
        A.pass1(b, o)
        print("pass1 wall", round(time.perf_counter() - t0, 3), "kernel ms", g.last_timing()["ms_kernels"], "route", [int(x) for x in g.last_pass1_stats()],
              "cells", g.last_pass1_cells() if hasattr(g, "last_pass1_cells") else None)
    g.close()


if __name__ == "__main__":
    main()
