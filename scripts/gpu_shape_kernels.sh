#!/bin/bash
# ncu launch list (time, instructions, DRAM reads, issue rate per launch) of the LAST resident round of a read shape:
#   scripts/gpu_shape_kernels.sh c4 <tag>   ->  gpurun_out/<tag>_<shape>_launches.csv + a per-kernel table on stdout
shape=${1:-c4}; tag=${2:-shape}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none --csv --log-file gpurun_out/${tag}_${shape}_launches.csv python scripts/gpu_shape_round.py $shape 1000000 3 > /dev/null 2>&1
python - "$shape" "$tag" <<'PY'
import csv, sys
rows = [r for r in csv.reader(l for l in open(f"gpurun_out/{sys.argv[2]}_{sys.argv[1]}_launches.csv") if l.startswith('"'))]
h = rows[0]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
out = {}
for r in rows[1:]:
    out.setdefault((int(r[iid]), r[ik].split("(")[0]), {})[r[im].split(".")[0].split("__")[1]] = float(r[iv].replace(",", ""))
ids = sorted(out)
# the last round = the launches after the last cut_init_kernel
last = max(i for i, (k, name) in enumerate(ids) if "cut_init" in name)
tot = 0
for k in ids[last:]:
    v = out[k]
    tot += v["time_duration"]
    print(f"{k[1][:48]:48s} {v['time_duration']/1e3:9.1f} us  {v['inst_executed']/1e6:8.2f} M inst  {v['bytes_read']/1e6:7.1f} MB  issue {v['issue_active']:5.1f} %")
print("sum of kernels", round(tot / 1e6, 3), "ms")
PY
