mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:realign_kernelILi6 -s 2 -c 1 -o gpurun_out/prof_realign_r1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:entry_kernelILi1 -s 2 -c 1 -o gpurun_out/prof_entry_r1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
