mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:realign_kernel -s 7 -c 1 -o gpurun_out/prof_realign_r1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:entry_kernel -s 7 -c 1 -o gpurun_out/prof_entry_r1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out; tail -3 gpurun_out/ncu_full.log
