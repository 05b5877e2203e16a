# round-2 profile set A: launch list of the default step, host-side timeline of one resident round, ncu --set full of the dominant
# DP kernel (the K = 11 class: the largest at configs[1]) and of the consensus accumulation kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-rmt --no-extras --no-pass1 --no-parity --no-shapes > gpurun_out/r02a_ncu_list.log 2>&1
MIAGPU_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-rmt --no-extras --no-pass1 --no-parity --no-shapes > gpurun_out/r02a_trace.log 2>&1
MIAGPU_SERIAL_LAUNCH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair16_kernel -s 7 -c 1 -o gpurun_out/r02a_prof_pair16 python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras --no-parity --no-shapes > gpurun_out/r02a_ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/r02a_prof_tile python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras --no-parity --no-shapes > gpurun_out/r02a_ncu_full2.log 2>&1
ls -la gpurun_out | tail -6
