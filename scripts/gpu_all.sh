mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_main.log 2>&1; tail -1 gpurun_out/bench_main.log | cut -c1-600
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
MIAGPU_CHUNKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair16_kernel -s 10 -c 1 -o gpurun_out/prof_pair16_d python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_full.log 2>&1
MIAGPU_CHUNKS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 3 -c 1 -o gpurun_out/prof_tile_d python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -14
