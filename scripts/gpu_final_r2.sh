# round 2, final set on one GPU: all GPU tests, smoke, the default bench line and the reference arm, launch list, ncu --set full of the
# dominant DP kernel, BASELINE configs[2] at full size on one GPU
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; tail -3 gpurun_out/r02f_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; tail -c 600 gpurun_out/r02f_bench_reference.json
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 300 gpurun_out/r02f_bench.json; tail -2 gpurun_out/r02f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-rmt --no-extras --no-pass1 --no-parity --no-shapes > gpurun_out/r02f_ncu_list.log 2>&1
MIAGPU_SERIAL_LAUNCH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair16_kernel -s 7 -c 1 -o gpurun_out/r02f_prof_pair16 python bench.py --steps 1 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras --no-parity --no-shapes > gpurun_out/r02f_ncu_full1.log 2>&1
python scripts/gpu_assembly.py --shape c3 --reads 10000000 --prefix 2000 > gpurun_out/r02f_c3_n1.json 2> gpurun_out/r02f_c3_n1.err; tail -c 900 gpurun_out/r02f_c3_n1.json
ls -la gpurun_out | grep r02f
