mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
for N in 1 2 4 8; do
  if [ "$N" -le "$NG" ]; then
    if [ "$N" -eq 1 ]; then
      python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-pass1 --no-rmt --no-extras > gpurun_out/scale_n$N.log 2>&1
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.log 2>&1
    fi
    tail -1 gpurun_out/scale_n$N.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=$N', 'value %.1f M reads/s' % (d['value']/1e6), 'ms %.3f' % d['ms_per_step'], 'e2e %.1f M reads/s' % (d['e2e']['value']/1e6), 'ms %.3f' % d['e2e']['ms_per_step'], d['consensus_matches_e2e'], d['score_cut'], d['clocks'])" || tail -20 gpurun_out/scale_n$N.log
  fi
done
