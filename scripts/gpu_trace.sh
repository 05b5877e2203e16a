mkdir -p gpurun_out
MIAGPU_TRACE=1 MIAGPU_CHUNKS=4 python bench.py --steps 2 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_trace.log 2>&1
grep "miagpu trace" gpurun_out/bench_trace.log | tail -24
