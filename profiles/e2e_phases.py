#!/usr/bin/env python
"""Wall-clock split of one end-to-end step of bench.py (host buffers in, consensus out).  GPU box only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench, _pkg
_pkg.load()
from mia_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref, bases, off, rc, as_, ae = bench.make_workload(n, seed=1000)
g = api.MiaGpu(0); g.set_pssm(bench.load_pssm()); g.set_reference(ref, circular=1, with_rc=0)
pin = lambda a: torch.from_numpy(a).pin_memory()
hb, ho, hr, ha, he = pin(bases), pin(off), pin(rc), pin(as_), pin(ae)
out = api.MiaGpu.alloc_realign_outputs(n, pinned=True); del out["runs"]
packed = torch.empty(4 * n, dtype=torch.int16).pin_memory()
below = torch.zeros(n, dtype=torch.uint8).pin_memory()
seq_len = np.diff(off).astype(np.int32)
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for it in range(6):
    t = time.perf_counter(); g.upload_reads(hb, ho); tick("upload_reads", t)
    t = time.perf_counter(); g.realign(hr, ha, he, out) if hasattr(g, "realign") else None; tick("realign(+h2d rc/as/ae, d2h results)", t)
    tm = g.last_timing()
    T.setdefault("  of which kernels (events)", []).append(tm["ms_kernels"]); T.setdefault("  of which d2h (events)", []).append(tm["ms_d2h"]); T.setdefault("  of which h2d (events)", []).append(tm["ms_h2d"])
    t = time.perf_counter(); g.get_runs_packed(None, packed); tick("get_runs_packed", t)
    t = time.perf_counter(); api.cull_flags(seq_len, out["score"].numpy(), out=below.numpy()); tick("cull_flags (host)", t)
    t = time.perf_counter(); g.consensus_natural(below, below, 1, want_gaps=False); tick("consensus_natural", t)
for k, v in T.items():
    print(f"{k:45s} {np.median(v[2:]):8.3f} ms")
