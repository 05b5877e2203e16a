mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-rmt --pass1-unmasked-reads 100000 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep16_kernel -s 1 -c 1 -o gpurun_out/prof_sweep16_e python bench.py --steps 1 --warmup 3 --no-cpu --no-rmt --pass1-unmasked-reads 100000 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out | tail -5
