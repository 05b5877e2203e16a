python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for b in 3 4 6 8; do echo "blocks/SM=$b"; MIAGPU_BLOCKS_PER_SM=$b python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gcups')}, d['buckets'], d['e2e']['ms_per_step'])"; done
