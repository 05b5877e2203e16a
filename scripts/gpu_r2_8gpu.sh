# round 2, one 8-GPU box: the 1 -> 8 curve of the bench line, BASELINE configs[2] and configs[4] at full size, the C host on 4 GPUs
mkdir -p gpurun_out
bash scripts/gpu_scale_r2.sh > gpurun_out/r02c_scale.txt 2>&1; cat gpurun_out/r02c_scale.txt | cut -c1-400
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
$TR scripts/gpu_assembly.py --shape c3 --reads 10000000 --prefix 0 > gpurun_out/r02c_c3_n8.json 2> gpurun_out/r02c_c3_n8.err; tail -c 1500 gpurun_out/r02c_c3_n8.json
$TR scripts/gpu_assembly.py --shape c5 --reads 50000000 --prefix 0 > gpurun_out/r02c_c5_n8.json 2> gpurun_out/r02c_c5_n8.err; tail -c 1800 gpurun_out/r02c_c5_n8.json; tail -3 gpurun_out/r02c_c5_n8.err
python -m pytest tests/test_gpu_host_c.py -m gpu -q -k "one_box" 2>&1 | tail -3
