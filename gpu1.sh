python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gcups')}, d['e2e'])"
