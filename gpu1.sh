set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import _pkg; _pkg.load()
from mia_b200 import api, synth
import gpu_checks
g = api.MiaGpu(0)
print("int32 peak ops/s", g.int32_peak())
ref = synth.random_reference(16569, 1)
genome = synth.diverge(ref, 0.005, seed=2)
N=1000000
t=time.time(); bases, off, truth = synth.make_reads(genome, N, 35, 75, seed=2); print("gen", time.time()-t)
rc = truth["strand"].astype(np.uint8)
as_ = truth["start"].astype(np.int32); ae = (as_ + truth["length"] - 1).astype(np.int32)
g.set_pssm(gpu_checks.load_pssm("onepass")); g.set_reference(ref, 1, 0)
for it in range(4):
    t=time.time(); out = g.realign_host(bases, off, rc, as_, ae); dt=time.time()-t
    tm = g.last_timing()
    print("wall %.1f ms"%(dt*1e3), tm, "GCUPS(kernels) %.1f"%(tm["dp_cells"]/tm["ms_kernels"]/1e6), "reads/s %.3g"%(N/(tm["ms_kernels"]*1e-3)))
print("score mean", out["score"].mean(), "status nonzero", (out["status"]!=0).sum(), "nruns>1", (out["n_runs"]>1).sum())
PY
