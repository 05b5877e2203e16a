mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','steps','gpu_launches','clocks')}); print('e2e', d['e2e']); print(d['roofline']); print(d['roofline_issue']); print(d['pass1']['k12']); print(d['cpu_baseline'])" || tail -20 gpurun_out/bench_default.log
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
