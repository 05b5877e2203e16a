"""C5-like timing (1 Mb linear reference): pass 1 with k = 12 / 14 and one resident round."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg; _pkg.load()
from mia_b200 import api, synth
import gpu_checks
L, n = 1_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ref = synth.random_reference(L, seed=321)
genome = synth.diverge(ref, 0.005, seed=322)
b, off, truth = synth.make_reads(genome, n, 35, 75, seed=323, circular=False)
g = api.MiaGpu(0)
g.set_pssm(gpu_checks.load_pssm("onepass"))
g.set_reference(ref, circular=0, with_rc=1)
g.upload_reads(b, off)
res = {}
for k in (12, 14):
    t0 = time.perf_counter(); g.build_kmers(k); tb = time.perf_counter() - t0
    g.pass1(); a = g.pass1(); t = g.last_timing()
    res[f"pass1_k{k}"] = dict(build_s=tb, kernel_ms=t["ms_kernels"], stats=g.last_pass1_stats(), accepted=int((a["score"] >= 2000).sum()))
ok = (a["hits"] > 0) & (a["score"] >= 2000)
g.compact_reads(ok.astype(np.uint8), (a["rc"] == 1).astype(np.uint8))
idx = np.flatnonzero(ok)
g.set_reference(ref, circular=0, with_rc=0)
g.set_alignment_inputs(a["rc"][idx].copy(), a["as_"][idx].copy(), a["ae"][idx].copy())
g.set_cut_inputs(np.diff(off).astype(np.int32)[idx])
import torch
for _ in range(3):
    g.reset_dropped(); g.iterate_resident()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    g.reset_dropped(); cons = g.iterate_resident()[0]
torch.cuda.synchronize(); res["round_ms"] = (time.perf_counter() - t0) / 5 * 1e3
res["reads"] = int(len(idx)); res["cons_len"] = len(cons)
print(json.dumps(res))
