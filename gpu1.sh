python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gcups')}, d['e2e'], d['roofline_int32'], d['cpu_baseline']['value'])"
