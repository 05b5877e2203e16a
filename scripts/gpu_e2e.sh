mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for ch in 1 2 4 6 8; do
MIAGPU_CHUNKS=$ch python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks=$ch', {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['consensus_matches_e2e'])" || tail -20 gpurun_out/bench_quick.log
done
