mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for t in 1 0; do
MIAGPU_CONS_TILES=$t python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tiles=$t', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['consensus_matches_e2e'])" || tail -20 gpurun_out/bench_quick.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_tmp.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_list.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_tmp.csv | grep -E "tile|entry|gaps|ent_pos|classify|layout"
