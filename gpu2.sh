mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:realign_kernel -s 10 -c 1 -o gpurun_out/prof_realign_r1b python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5; grep -c "" gpurun_out/ncu_full.log
