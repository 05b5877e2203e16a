"""host/mia_gpu alone on the GPU (no other CUDA process): 1 M C2 reads, FASTQ -> final .maln, three runs; prints the program's own phase timing."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg; _pkg.load()
import bench
from mia_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = bench.make_workload(n, 2)
ref, bases, off, rc = w[0], w[1], w[2], w[3]
import numpy as np
comp = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTN", b"TGCAN"):
    comp[a] = b
rid = np.repeat(np.arange(n), np.diff(off)); pos = np.arange(len(bases)) - off[rid]
src = np.where(rc[rid] == 1, off[rid] + (off[rid + 1] - off[rid]) - 1 - pos, np.arange(len(bases)))
orig = np.ascontiguousarray(np.where(rc[rid] == 1, comp[bases[src]], bases), np.uint8)
with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "ref.fa"), "w").write(">ref synthetic\n" + ref + "\n")
    open(os.path.join(d, "m.txt"), "w").write(synth.matrix_text(bench.load_pssm()))
    open(os.path.join(d, "all.fq"), "wb").write(synth.fastq_text(orig, off))
    for rep in range(int(os.environ.get("REPS", "3"))):
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "host", "mia_gpu"), "-r", "ref.fa", "-f", "all.fq", "-s", "m.txt", "-m", "out", "-c", "-k", "12", "-F", "--drop-score-2000"],
                           cwd=d, capture_output=True, text=True)
        print("run", rep, "reads", n, "wall_s %.3f" % (time.perf_counter() - t0), "fastq_mb %.0f" % (os.path.getsize(os.path.join(d, "all.fq")) / 1e6), "maln_mb", [round(os.path.getsize(os.path.join(d, f)) / 1e6) for f in os.listdir(d) if f.startswith("out.")], [l for l in r.stderr.split("\n") if "timing" in l or "convergence" in l or "left out" in l])
