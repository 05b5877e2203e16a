cd mapping-iterative-assembler_b200
cp libmiagpu.so libmiagpu_base.so
for v in base six_pp six_ip five_ip; do
  cp libmiagpu_$v.so libmiagpu.so
  cd ..; python bench.py --steps 10 --no-cpu --no-rmt --no-pass1 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['consensus_matches_e2e'], [(b['K'], round(b['ms'],3)) for b in d['buckets'] if b['kernel'].startswith('pair16')])"; cd mapping-iterative-assembler_b200
done
cp libmiagpu_base.so libmiagpu.so
