"""Whole assemblies at the sizes BASELINE.json names, with everything resident: pass 1 (k-mer filter) and rounds until the
consensus stops changing, on one GPU (plain python) or with the reads sharded over the GPUs of one box (torchrun, NCCL).

    --shape c3   configs[2]: merged paired-end reads 30-140 bp, ancient.submat.solexa.pe, 16,569 bp circle, k = 12
    --shape c4   configs[3]: reads of a sample 10 % + 0.5 % indels away from the 16,569 bp seed reference, ancient.submat, k = 12
                 (--distant: mia -D; over several GPUs only while no stale pointer crosses a shard boundary)
    --shape c5   configs[4]: 1 Mb linear reference, 35-75 bp reads, ancient.submat.solexa.onepass, k = 14

The data set does not depend on the number of GPUs (8 seeded pieces; rank r of W takes pieces [8r/W, 8(r+1)/W)), so the md5 of every
round's consensus must agree between W = 1 and W = 8.  Rank 0 then checks a PREFIX of the same reads (--prefix, default 10,000) as a
whole assembly on one GPU against the CPU checker (oracle/_ref: the unmodified reference's own main loop): SURVEY 8d / H11.

    python scripts/gpu_assembly.py --shape c3 --reads 10000000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/gpu_assembly.py --shape c5 --reads 50000000
"""
import argparse
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
SHAPES = {
    "c3": dict(ref_len=16569, circular=1, div=0.005, indel=0.0, lens=(30, 140), matrix="pe", k=12,
               name="BASELINE configs[2]: merged PE reads 30-140 bp, ancient.submat.solexa.pe, 16,569 bp circular R-rand, k = 12"),
    "c4": dict(ref_len=16569, circular=1, div=0.10, indel=0.005, lens=(35, 75), matrix="ancient", k=12,
               name="BASELINE configs[3]: sample 10 % + 0.5 % indels away from the 16,569 bp circular seed reference, ancient.submat, k = 12"),
    "c5": dict(ref_len=1_000_000, circular=0, div=0.005, indel=0.0, lens=(35, 75), matrix="onepass", k=14,
               name="BASELINE configs[4]: 1 Mb linear R-rand reference, 35-75 bp reads, ancient.submat.solexa.onepass, k = 14"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="c3", choices=sorted(SHAPES))
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--prefix", type=int, default=10_000)
    ap.add_argument("--distant", action="store_true")
    ap.add_argument("--legacy-shards", action="store_true", help="sharded rounds without the pointer state (round-1 behaviour: reads scoring exactly 2000 left out, split changes counted)")
    ap.add_argument("--parity-only", action="store_true", help="only the prefix check of the data set --reads names (one GPU)")
    args = ap.parse_args()
    sh = SHAPES[args.shape]
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ref = synth.random_reference(sh["ref_len"], seed=1 if sh["ref_len"] < 100000 else 321)
    genome = synth.diverge(ref, sh["div"], seed=3, indel_rate=sh["indel"])
    per = args.reads // PIECES
    if args.parity_only:
        t0 = time.perf_counter()
        sm = gpu_checks.load_pssm(sh["matrix"])
        b0, o0 = synth.make_reads(genome, per, sh["lens"][0], sh["lens"][1], seed=1000, circular=bool(sh["circular"]))[:2]
        P = min(args.prefix, per)
        reads = [synth.read_str(b0, o0, i) for i in range(P)]
        gp = api.MiaGpu(local)
        par = gpu_checks.assembly_parity(gp, ref, reads, sm, sh["circular"], sh["k"], int(args.distant))
        gp.close()
        par["wall_s"] = round(time.perf_counter() - t0, 1)
        print(json.dumps(dict(config=sh["name"] + (" -D" if args.distant else ""), data_set_reads=per * PIECES, prefix_parity=par)))
        return
    t0 = time.perf_counter()
    mine = range(rank * PIECES // world, (rank + 1) * PIECES // world)
    parts = [synth.make_reads(genome, per, sh["lens"][0], sh["lens"][1], seed=1000 + p, circular=bool(sh["circular"]))[:2] for p in mine]
    bases = np.concatenate([b for b, _ in parts])
    off = np.concatenate([[0]] + [o[1:] + sum(int(q[1][-1]) for q in parts[:i]) for i, (_, o) in enumerate(parts)]).astype(np.int64)
    t_gen = time.perf_counter() - t0
    sm = gpu_checks.load_pssm(sh["matrix"])

    class Exchange:
        rounds = None
        rank = int(os.environ.get("RANK", "0"))

        @staticmethod
        def all_gather_host(a):
            """per-read arrays of all ranks in rank order, through the device (NCCL all-gather of padded pieces)"""
            a = np.ascontiguousarray(a)
            cnt = torch.tensor([len(a)], device="cuda", dtype=torch.int64)
            cnts = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(cnts, cnt)
            cnts = [int(c) for c in cnts]
            pad = torch.zeros(max(cnts), dtype=torch.from_numpy(a[:1]).dtype, device="cuda")
            pad[: len(a)] = torch.from_numpy(a).cuda()
            out = torch.empty(world * max(cnts), dtype=pad.dtype, device="cuda")
            dist.all_gather_into_tensor(out, pad)
            out = out.cpu().numpy().reshape(world, -1)
            return np.concatenate([out[r, : cnts[r]] for r in range(world)])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    g = api.MiaGpu(local)
    mk = lambda gg, x=None: driver.ResidentAssembler(gg, ref, sm, circular=sh["circular"], k=sh["k"], exchange=x, strand_unknown="drop",
                                                     distant_ref=int(args.distant), pointer_state=None if (x is None or args.legacy_shards) else True)
    W = mk(g)                                          # load every kernel once on a small prefix: the timed calls show steady cost
    m = min(20000, len(off) - 1)
    W.pass1(np.ascontiguousarray(bases[: off[m]]), np.ascontiguousarray(off[: m + 1]))
    W.iterate()
    A = mk(g, Exchange if world > 1 else None)
    barrier()
    t0 = time.perf_counter()
    A.pass1(bases, off)
    barrier()
    t_pass1 = time.perf_counter() - t0
    p1_kernel_ms = g.last_timing()["ms_kernels"]
    p1_route = [int(x) for x in g.last_pass1_stats()]
    n_local = len(A.seq_len)
    if world > 1:
        t = torch.tensor([n_local, A.strand_unknown_reads], device="cuda", dtype=torch.int64)
        mx = t.clone()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        n_fsdb, n_unknown, n_max = int(t[0]), int(t[1]), int(mx[0])
        if A.fs:
            u = torch.tensor([int((~A.strand_known).sum())], device="cuda", dtype=torch.int64)
            dist.all_reduce(u)
            n_unknown = int(u[0])
        Exchange.rounds = shard.ShardedRounds(g, local, world, rank, n_max)
    else:
        n_fsdb, n_unknown = n_local, int((~A.strand_known).sum())
    rounds, conv = [], False
    while not conv and A.iter < 30:
        barrier()
        t0 = time.perf_counter()
        cons, conv = A.iterate()
        barrier()
        rounds.append(dict(ms=(time.perf_counter() - t0) * 1e3, cons_len=len(cons), md5=hashlib.md5(cons.encode()).hexdigest(),
                           dropped=int(A.dropped.sum()), strand_unknown=int((~A.strand_known).sum())))
    same = True
    extra = {}
    if world > 1:
        lst = [None] * world
        dist.all_gather_object(lst, [r["md5"] for r in rounds])
        same = all(x == lst[0] for x in lst)
        d = torch.tensor([r["dropped"] for r in rounds], device="cuda", dtype=torch.int64)
        dist.all_reduce(d, op=dist.ReduceOp.SUM)
        for r, v in zip(rounds, d.tolist()):
            r["dropped"] = v
        # what the round's largest collective costs: the SUM all-reduce of the column planes (10 int32 per column)
        words = 10 * len(cons)
        buf = torch.zeros(words, dtype=torch.int32, device="cuda")
        for _ in range(3):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(20):
            dist.all_reduce(buf)
        ev[1].record()
        torch.cuda.synchronize()
        ar_ms = ev[0].elapsed_time(ev[1]) / 20
        extra["allreduce_sum_planes"] = dict(bytes=4 * words, ms=ar_ms, share_of_last_round=ar_ms / rounds[-1]["ms"])
    # the column accumulation alone (entry_kernel<1> with global REDs beyond 64 tiles, tile_kernel below), local reads
    g.accumulate_gaps_natural()
    g.accumulate_counts()
    acc_ms = g.last_timing()["ms_kernels"]
    visits = int(A.seq_len.sum())
    extra["accumulate"] = dict(kernel="tile_kernel" if (len(cons) + 1791) // 1792 <= 64 else "entry_kernel<1>", ms=acc_ms, column_visits=visits,
                               g_visits_per_s=visits / (acc_ms * 1e-3) / 1e9)
    if rank == 0:
        ident = sum(a == b for a, b in zip(cons, genome)) / max(len(cons), len(genome)) if len(cons) == len(genome) else None
        total_s = t_pass1 + sum(r["ms"] for r in rounds) / 1e3
        out = dict(config=sh["name"] + (" -D" if args.distant else ""), n_gpus=world, reads=per * PIECES, reads_in_fsdb=n_fsdb, generate_s_per_rank=t_gen,
                   pass1_s=t_pass1, pass1_kernel_ms_rank0=p1_kernel_ms, rounds=len(rounds), converged=bool(conv), per_round=rounds, assembly_s=total_s,
                   reads_per_s_whole_assembly=per * PIECES / total_s,
                   reads_per_s_per_round=per * PIECES / (sum(r["ms"] for r in rounds) / len(rounds) / 1e3),
                   all_ranks_same_consensus=same, consensus_equals_sample_genome=(cons == genome), identity_to_sample_genome=ident,
                   pointer_state=bool(A.fs), split_changes_not_modelled=int(A.split_changes),
                   reads_scoring_exactly_2000_left_out=n_unknown if not A.fs else 0, strand_unknown_reads=n_unknown if A.fs else 0,
                   pass1_route_rank0=dict(zip(("pair_kernels", "general_kernel", "no_kmer_hit"), p1_route)), **extra)
        if args.prefix > 0:                            # the first reads of the data set as an assembly of their own, against the CPU checker
            t0 = time.perf_counter()
            b0, o0 = synth.make_reads(genome, per, sh["lens"][0], sh["lens"][1], seed=1000, circular=bool(sh["circular"]))[:2]
            P = min(args.prefix, per)
            reads = [synth.read_str(b0, o0, i) for i in range(P)]
            gp = api.MiaGpu(local)
            par = gpu_checks.assembly_parity(gp, ref, reads, sm, sh["circular"], sh["k"], int(args.distant))
            gp.close()
            par["wall_s"] = round(time.perf_counter() - t0, 1)
            out["prefix_parity"] = par
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
