#!/bin/bash
# per-launch time / instructions / DRAM reads of the consensus kernels inside a resident bench step (ncu launch list; a number
# printed under ncu is never a bench value): scripts/gpu_cons_kernels.sh <tag>  ->  gpurun_out/<tag>_cons.csv
tag=${1:-cons}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"tile_kernel|ent_bin|natural_entries|undo_kernel" -s 15 -c 6 --csv --log-file gpurun_out/${tag}_cons.csv \
    python bench.py --no-cpu --no-pass1 --no-extras --no-parity --no-shapes --steps 2 --warmup 3 > /dev/null 2>&1
python - "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/{sys.argv[1]}_cons.csv") if l.startswith('"'))]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((r[iid], r[ik].split("(")[0]), {})[r[im].split(".")[0].split("__")[1]] = r[iv]
for k, v in out.items():
    print(k[0], k[1], v)
PY
