mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -8
python -m pytest tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_shard.log
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_shard.py 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['consensus_matches_e2e'], d['score_cut'])" || tail -20 gpurun_out/bench_n1.log
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['consensus_matches_e2e'], d['score_cut'])" || tail -30 gpurun_out/bench_n2.log
fi
