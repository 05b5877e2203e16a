mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gcups')}, d['e2e']['ms_per_step'], [(b['kernel'],b['reads'],round(b['ms'],3)) for b in d['buckets']], d['pair16'], d['consensus_matches_e2e'])" || tail -20 gpurun_out/bench_quick.log
