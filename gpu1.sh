mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.log 2>&1; tail -1 gpurun_out/bench_r1c.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.log 2>&1; tail -1 gpurun_out/bench_ref_r1c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:realign_kernel -s 10 -c 1 -o gpurun_out/prof_realign_r1c python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
