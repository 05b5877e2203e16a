python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gcups','gpu_launches','buckets')}); print(d['e2e']); print(d['roofline_int32']); print(d['clocks'])"
