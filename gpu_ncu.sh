mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair16_kernel -s 7 -c 1 -o gpurun_out/prof_pair16_b python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -3
