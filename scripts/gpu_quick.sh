mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['consensus_matches_e2e']); print(d['pass1'])" || tail -20 gpurun_out/bench_quick.log
