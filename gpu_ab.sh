mkdir -p gpurun_out
L=mapping-iterative-assembler_b200
for v in mb1 mb5 mb6; do
  cp $L/libmiagpu_$v.so $L/libmiagpu.so
  for bps in 8; do
  MIAGPU_PAIR_BLOCKS_PER_SM=$bps python bench.py --steps 5 --warmup 3 --no-cpu --no-pass1 > gpurun_out/bench_$v.log 2>&1; tail -1 gpurun_out/bench_$v.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', $bps, {k:round(d[k],3) for k in ('ms_per_step',)}, [(b['kernel'],b['reads'],round(b['ms'],3)) for b in d['buckets'][:3]], d['consensus_matches_e2e'])" || tail -5 gpurun_out/bench_$v.log
  done
done
