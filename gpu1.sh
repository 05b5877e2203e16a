python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gcups')}); print(d['pass1'])"
