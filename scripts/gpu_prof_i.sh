# round-1 final profile set (r01i): default bench line, reference arm, launch list of the same step, ncu --set full of the
# dominant DP kernel and of the consensus accumulation kernel
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_i.log 2>&1; tail -1 gpurun_out/bench_i.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_i.log 2>&1; tail -1 gpurun_out/bench_ref_i.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-rmt --no-extras --pass1-unmasked-reads 100000 > gpurun_out/ncu_list_i.log 2>&1
MIAGPU_CHUNKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair16_kernel -s 10 -c 1 -o gpurun_out/prof_pair16_i python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras > gpurun_out/ncu_full_i1.log 2>&1
MIAGPU_CHUNKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_i python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras > gpurun_out/ncu_full_i2.log 2>&1
ls -la gpurun_out | tail -8
