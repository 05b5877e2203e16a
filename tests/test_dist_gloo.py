"""CPU, world_size 2 over gloo: the multi-GPU plan of SURVEY 8e -- reads sharded, consensus
replicated, gaps all-reduced with MAX and column planes with SUM, then every rank calls the
same bases.  The per-rank planes come from the oracle here (no GPU); the property under test
is that the reduction of shard accumulators reproduces the single-process consensus bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gpu_checks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard_planes(o, ref, bases, off, rc, as_, ae, sm, lo, hi, forced_gaps=None):
    """(gaps, counts, ins) of reads [lo,hi): realign + natural round via the oracle."""
    ctx = o.ctx_new(ref, 1, sm, with_rc=0, k=0)
    res = [None] * (len(off) - 1)
    for i in range(lo, hi):
        res[i] = o.realign(ctx, bases[off[i]:off[i + 1]].tobytes().decode(), int(rc[i]), int(as_[i]), int(ae[i]))
    o.ctx_free(ctx)
    dropped = np.zeros(len(off) - 1, np.uint8)
    cons, gaps, counts, _ = gpu_checks.oracle_round(o, ref, bases, off, rc, res, sm, 1, dropped)
    return cons, gaps[:len(ref)].copy(), counts


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle.pyoracle import Oracle
    o = Oracle()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sm = gpu_checks.load_pssm("onepass")
    # no indels: base-column planes only (insert columns depend on the reduced gaps layout, covered on the GPU)
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(400, 1200, seed=81, divergence=0.03, indel_rate=0.0)
    n = len(off) - 1
    lo, hi = rank * n // world, (rank + 1) * n // world
    _, gaps, counts = _shard_planes(o, ref, bases, off, rc, as_, ae, sm, lo, hi)
    tg, tc = torch.from_numpy(gaps.copy()), torch.from_numpy(counts.copy())
    dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    called = "".join(o.find_consensus(tc[p].numpy(), 1) for p in range(len(ref))).replace("-", "")
    if rank == 0:
        full_cons, full_gaps, full_counts = _shard_planes(o, ref, bases, off, rc, as_, ae, sm, 0, n)
        q.put((called == full_cons, bool((tc.numpy() == full_counts).all()), bool((tg.numpy() == full_gaps).all())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_accumulators_reduce_to_single_process_consensus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert res == (True, True, True)


def _chain_worker(rank, world, port, q):
    """-D over shards: driver.ResidentAssembler.begin_round hands every rank the matrix state the ranks before it leave (H6),
    all-gathered over the process group -- here gloo, a stub in place of the GPU context."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _pkg
    _pkg.load()
    from mia_b200 import driver
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(17)                       # the same stream on every rank: everybody knows every rank's function
    FUNCS = [(0, 1), (0, 0), (1, 1)]                      # identity (no reads / untouched reads), constant forward, constant reversed

    class Stub:
        after, entered = (0, 1), None

        def set_reference(self, *a, **k):
            pass

        def distant_retry_begin(self):
            return 5, list(self.after)

        def distant_retry_end(self, s):
            self.entered = s
            return 2

    class Exchange:
        rank = 0
        rounds = None

        @staticmethod
        def all_gather_host(a):
            t = torch.from_numpy(np.ascontiguousarray(a))
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return np.concatenate([x.numpy() for x in out])

    Exchange.rank = rank
    A = object.__new__(driver.ResidentAssembler)
    A.g, A.x, A.distant_ref, A.matrix_state, A._manual_retry = Stub(), Exchange, 1, 0, False
    A.cons, A.last, A.iter, A.circular = None, "ACGT", 0, 1
    ok, carried = True, 0
    for _ in range(12):
        funcs = [FUNCS[int(rng.integers(0, 3))] for _ in range(world)]
        A.g.after = funcs[rank]
        A.begin_round()
        s = carried
        for r in range(rank):
            s = funcs[r][s]
        ok &= A.g.entered == s
        for r in range(rank, world):
            s = funcs[r][s]
        ok &= A.matrix_state == s
        carried = s
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        q.put(all(flags))
    dist.barrier()
    dist.destroy_process_group()


def test_distant_retry_matrix_state_crosses_rank_boundaries():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_chain_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert res is True
