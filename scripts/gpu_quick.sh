mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for ch in 4 3 6; do
MIAGPU_CHUNKS=$ch python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunks=$ch', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['consensus_matches_e2e'])" || tail -20 gpurun_out/bench_quick.log
done
MIAGPU_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_trace.log 2>&1
grep "miagpu trace" gpurun_out/bench_trace.log | tail -16
