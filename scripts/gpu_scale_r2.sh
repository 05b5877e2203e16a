# round 2: the 1 -> 8 curve of the default bench line (resident + e2e arms, parity block at every N) on one 8-GPU box
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu --no-rmt --no-extras --no-pass1 --no-shapes > gpurun_out/r02b_scale_n1.json 2> gpurun_out/r02b_scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02b_scale_n$n.json 2> gpurun_out/r02b_scale_n$n.err
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02b_scale_n{n}.json").read().strip().split("\n")[-1])
    except Exception as e:
        print(n, "failed", e); continue
    if n==1: base=(d["value"], d["e2e"]["value"])
    print(n, round(d["ms_per_step"],3), round(d["value"]/1e6,1), "eff", round(d["value"]/n/base[0],3), "| e2e", round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["value"]/1e6,1), "eff", round(d["e2e"]["value"]/n/base[1],3), "| parity", d.get("parity"))
PY
