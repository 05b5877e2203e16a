"""One resident round on a BASELINE read shape, for an ncu launch list of that round (scripts/gpu_shape_kernels.sh):
    python scripts/gpu_shape_round.py c4|c3|c2 [reads] [rounds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import _pkg
_pkg.load()
from mia_b200 import api
import bench

shape = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw, mat = {"c4": (dict(divergence=0.10, indel_rate=0.005), "ancient"),
           "c3": (dict(divergence=0.005, indel_rate=0.0005, min_len=30, max_len=140), "pe"),
           "c2": (dict(), "onepass")}[shape]
ref, bases, off, rc, as_, ae = bench.make_workload(n, seed=3000, **kw)
g = api.MiaGpu(0)
g.set_pssm(bench.load_pssm(mat))
g.set_reference(ref, circular=1, with_rc=0)
g.upload_reads(bases, off)
g.set_alignment_inputs(rc, as_, ae)
g.set_cut_inputs(np.diff(off).astype(np.int32))
import time
for k in range(rounds):
    g.reset_dropped()
    t0 = time.perf_counter()
    g.iterate_resident()
    print("round", k, round((time.perf_counter() - t0) * 1e3, 3), "ms wall", file=sys.stderr)
g.close()
