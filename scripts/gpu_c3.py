"""BASELINE configs[2] at full size: N synthetic merged paired-end aDNA reads (30-140 bp, ancient.submat.solexa.pe) against a
16.5 kb circular reference, pass 1 (k = 12) and rounds to convergence with everything resident, reads sharded over the GPUs
of one box (one process per GPU, NCCL; launch with torchrun) or on one GPU (plain python).

The data set does not depend on the number of GPUs: it is 8 seeded pieces of N/8 reads, rank r of W takes pieces
[8r/W, 8(r+1)/W) -- so the md5 of every round's consensus must be the same at W = 1 and W = 8 (integer sums, exact score cut:
SURVEY 8e), and the converged consensus should be the sample genome the reads were drawn from.

    python scripts/gpu_c3.py [N]                                         # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/gpu_c3.py [N]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

_pkg.load()
from mia_b200 import api, driver, shard, synth  # noqa: E402
import gpu_checks  # noqa: E402

PIECES = 8


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ref = synth.random_reference(16569, seed=1)
    genome = synth.diverge(ref, 0.005, seed=3)
    per = n_total // PIECES
    t0 = time.perf_counter()
    parts = [synth.make_reads(genome, per, 30, 140, seed=1000 + p)[:2] for p in range(rank * PIECES // world, (rank + 1) * PIECES // world)]
    bases = np.concatenate([b for b, _ in parts])
    off = np.concatenate([[0]] + [o[1:] + sum(int(q[1][-1]) for q in parts[:i]) for i, (_, o) in enumerate(parts)]).astype(np.int64)
    t_gen = time.perf_counter() - t0

    class Exchange:
        rounds = None

        @staticmethod
        def all_gather_host(a):
            lst = [None] * world
            dist.all_gather_object(lst, a)
            return np.concatenate(lst)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    g = api.MiaGpu(local)
    A = driver.ResidentAssembler(g, ref, gpu_checks.load_pssm("pe"), circular=1, k=12, exchange=Exchange if world > 1 else None, strand_unknown="drop")
    if os.environ.get("C3_WARM"):                      # load every kernel once on a small prefix, so that the timed calls show steady cost
        W = driver.ResidentAssembler(g, ref, gpu_checks.load_pssm("pe"), circular=1, k=12, strand_unknown="drop")
        m = min(20000, len(off) - 1)
        W.pass1(bases[: off[m]], off[: m + 1])
        W.iterate()
    barrier()
    t0 = time.perf_counter()
    A.pass1(bases, off)
    barrier()
    t_pass1 = time.perf_counter() - t0
    p1_route = [int(x) for x in g.last_pass1_stats()]          # reads finished by the windowed pair kernels, general kernel, no k-mer hit
    n_local = len(A.seq_len)
    if world > 1:
        t = torch.tensor([n_local, A.strand_unknown_reads], device="cuda", dtype=torch.int64)
        mx = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        n_fsdb, n_unknown, n_max = int(t[0]), int(t[1]), int(mx[0])
        Exchange.rounds = shard.ShardedRounds(g, local, world, rank, n_max)
    else:
        n_fsdb, n_unknown = n_local, A.strand_unknown_reads
    rounds, conv = [], False
    while not conv and A.iter < 30:
        barrier()
        t0 = time.perf_counter()
        cons, conv = A.iterate()
        barrier()
        rounds.append(dict(ms=(time.perf_counter() - t0) * 1e3, cons_len=len(cons), md5=hashlib.md5(cons.encode()).hexdigest(),
                           dropped=int(A.dropped.sum())))
    if world > 1:
        lst = [None] * world
        dist.all_gather_object(lst, rounds[-1]["md5"])
        same = len(set(lst)) == 1
        d = torch.tensor([r["dropped"] for r in rounds], device="cuda", dtype=torch.int64)
        dist.all_reduce(d, op=dist.ReduceOp.SUM)
        for r, v in zip(rounds, d.tolist()):
            r["dropped"] = v
    else:
        same = True
    if rank == 0:
        ident = sum(a == b for a, b in zip(cons, genome)) / max(len(cons), len(genome)) if len(cons) == len(genome) else None
        total_s = t_pass1 + sum(r["ms"] for r in rounds) / 1e3
        print(json.dumps(dict(config="BASELINE configs[2]: merged PE reads 30-140 bp, ancient.submat.solexa.pe, 16,569 bp circular R-rand, k = 12",
                              n_gpus=world, reads=per * PIECES, reads_in_fsdb=n_fsdb, generate_s_per_rank=t_gen, pass1_s=t_pass1,
                              rounds=len(rounds), converged=bool(conv), per_round=rounds, assembly_s=total_s,
                              reads_per_s_whole_assembly=per * PIECES / total_s,
                              reads_per_s_per_round=per * PIECES / (sum(r["ms"] for r in rounds) / len(rounds) / 1e3),
                              all_ranks_same_consensus=same, consensus_equals_sample_genome=(cons == genome), identity_to_sample_genome=ident,
                              split_changes=int(A.split_changes),
                              reads_scoring_exactly_2000_left_out=n_unknown,
                              pass1_route_rank0=dict(zip(("pair_kernels", "general_kernel", "no_kmer_hit"), p1_route)))))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
