"""which reads of the bench workload leave the pass-1 fast path, and what they cost in the general kernel"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import _pkg; _pkg.load()
from mia_b200 import api, synth
n = 1_000_000
ref, stored, off, rc, as_, ae = bench.make_workload(n, seed=1000)
comp = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"): comp[a] = b
rid = np.repeat(np.arange(n), np.diff(off)); pos = np.arange(len(stored)) - off[rid]
src = np.where(rc[rid] == 1, off[rid] + (off[rid + 1] - off[rid]) - 1 - pos, np.arange(len(stored)))
orig = np.ascontiguousarray(np.where(rc[rid] == 1, comp[stored[src]], stored), np.uint8)
g = api.MiaGpu(0)
g.set_pssm(bench.load_pssm("onepass")); g.set_reference(ref, circular=1, with_rc=1); g.build_kmers(12)
g.upload_reads(orig, off)
g.pass1(); out = g.pass1(); print("pass1 ms", g.last_timing()["ms_kernels"], g.last_pass1_stats())
route = g.last_pass1_route()
gen = np.flatnonzero(route >= 2)
L = np.diff(off)
print("general:", len(gen), "route2", int((route == 2).sum()), "route3", int((route == 3).sum()))
print("hits of general reads: ", np.sort(out["hits"][gen])[::-1][:20], "L", L[gen][:20], "n_runs", out["n_runs"][gen][:20])
# time the general kernel on just those reads
sub_off = np.zeros(len(gen) + 1, np.int64); np.cumsum(L[gen], out=sub_off[1:])
sub = np.concatenate([orig[off[i]:off[i + 1]] for i in gen])
os.environ["MIAGPU_PASS1_FAST"] = "0"
g.upload_reads(sub, sub_off); g.pass1(); g.pass1(); print("general kernel on these reads: ms", g.last_timing()["ms_kernels"])
for m in (1, 8, 64):
    g.upload_reads(sub[: sub_off[m]], sub_off[: m + 1]); g.pass1(); g.pass1(); print(m, "reads: ms", g.last_timing()["ms_kernels"])
