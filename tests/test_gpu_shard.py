"""GPU: sharded rounds (miagpu_shard_begin[_host] / _cut / _finish) against the single-GPU round over the
concatenated reads.  Bit-exact: slope and intercept as IEEE doubles, sticky flags, gaps, consensus, on every rank.
Several "ranks" are contexts of one process here (the collectives are device copies, mia_b200.shard.LocalShards); the
NCCL flavour of the same protocol runs when the box has two GPUs."""
import os
import struct
import sys

import numpy as np
import pytest

import gpu_checks

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bits(x):
    return struct.pack("<d", x)


def _case(n_reads, ref_len, seed):
    ref, bases, off, rc, as_, ae = gpu_checks.make_case(n_reads, ref_len, seed=seed, divergence=0.02, indel_rate=0.004)
    n = len(off) - 1
    seq_len = np.diff(off).astype(np.int32)
    sticky = (np.arange(n) % 97 == 0).astype(np.uint8)
    return ref, bases, off, rc, as_, ae, seq_len, sticky


def _single(gpu, ref, bases, off, rc, as_, ae, seq_len, sticky, unique=None, **kw):
    gpu.set_pssm(gpu_checks.load_pssm("onepass"))
    gpu.set_reference(ref, circular=1, with_rc=0)
    gpu.upload_reads(bases, off)
    gpu.set_alignment_inputs(rc, as_, ae)
    gpu.set_cut_inputs(seq_len, unique, sticky)
    d = np.zeros(len(seq_len), np.uint8)
    cons, fit, gaps = gpu.iterate_resident(dropped=d, want_gaps=True, **kw)
    return cons, fit, gaps, d


def _bounds(n, parts, ragged):
    if not ragged:
        return [n * r // parts for r in range(parts + 1)]
    w = np.array([1.0 + 0.6 * ((r * 7) % 3) for r in range(parts)])
    b = np.concatenate([[0], np.cumsum(w / w.sum() * n)]).astype(np.int64)
    b[-1] = n
    return b.tolist()


@pytest.mark.parametrize("n_reads,ref_len,seed,parts,ragged", [(9000, 3000, 4, 2, False), (120000, 6000, 5, 3, True), (40000, 2500, 6, 4, True)])
def test_sharded_resident_equals_single_gpu(gpu, n_reads, ref_len, seed, parts, ragged):
    from mia_b200 import api, shard
    ref, bases, off, rc, as_, ae, seq_len, sticky = _case(n_reads, ref_len, seed)
    cons, fit, gaps, drop = _single(gpu, ref, bases, off, rc, as_, ae, seq_len, sticky)
    bnd = _bounds(len(seq_len), parts, ragged)
    ctxs = [api.MiaGpu(0) for _ in range(parts)]
    try:
        for r, g in enumerate(ctxs):
            lo, hi = bnd[r], bnd[r + 1]
            g.set_pssm(gpu_checks.load_pssm("onepass"))
            g.set_reference(ref, circular=1, with_rc=0)
            g.upload_reads(np.ascontiguousarray(bases[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo]))
            g.set_alignment_inputs(rc[lo:hi].copy(), as_[lo:hi].copy(), ae[lo:hi].copy())
            g.set_cut_inputs(seq_len[lo:hi].copy(), None, sticky[lo:hi].copy())
        n_max = max(bnd[r + 1] - bnd[r] for r in range(parts))
        ds = [np.zeros(bnd[r + 1] - bnd[r], np.uint8) for r in range(parts)]
        L = shard.LocalShards(ctxs)
        res = L.resident(n_max, dropped=ds, want_gaps=True)
        for r, (c2, f2, g2) in enumerate(res):
            assert _bits(f2[0]) == _bits(fit[0]) and _bits(f2[1]) == _bits(fit[1]), (r, fit, f2)
            assert c2 == cons, r
            assert (g2 == gaps).all(), r
        assert (np.concatenate(ds) == drop).all()
        st = ctxs[0].last_cut_stats()
        assert st["serial_blocks"] >= 2                      # at least the first block of either chain
        # a second round: flags are sticky on the device, nothing changes
        res2 = L.resident(n_max, dropped=ds)
        assert res2[0][0] == cons and (np.concatenate(ds) == drop).all()
    finally:
        for g in ctxs:
            g.close()


def test_sharded_host_policy_variants(gpu):
    from mia_b200 import api, shard
    ref, bases, off, rc, as_, ae, seq_len, sticky = _case(30000, 4000, 33)
    n = len(seq_len)
    rng = np.random.default_rng(2)
    unique = (rng.random(n) < 0.8).astype(np.uint8)
    parts = 2
    bnd = _bounds(n, parts, True)
    ctxs = [api.MiaGpu(0) for _ in range(parts)]
    try:
        for g in ctxs:
            g.set_pssm(gpu_checks.load_pssm("onepass"))
            g.set_reference(ref, circular=1, with_rc=0)
        L = shard.LocalShards(ctxs)
        for kw in (dict(), dict(unique=unique), dict(hard_cut=9000), dict(score_cut=(150.0, -500.0))):
            u = kw.pop("unique", None)
            cons, fit, gaps, drop = _single(gpu, ref, bases, off, rc, as_, ae, seq_len, sticky, u, **kw)
            ref_out = gpu.realign_host(bases, off, rc, as_, ae)
            tot_ref = int(np.maximum(ref_out["n_runs"], 0).sum())
            shards = []
            for r in range(parts):
                lo, hi = bnd[r], bnd[r + 1]
                shards.append(dict(bases=np.ascontiguousarray(bases[off[lo]:off[hi]]), off=np.ascontiguousarray(off[lo:hi + 1] - off[lo]),
                                   rc=rc[lo:hi].copy(), as_=as_[lo:hi].copy(), ae=ae[lo:hi].copy(), seq_len=seq_len[lo:hi].copy(),
                                   dropped=sticky[lo:hi].copy(), out=api.MiaGpu.alloc_realign_outputs(hi - lo),
                                   packed=np.zeros(8 * (hi - lo), np.uint16), unique_best=None if u is None else u[lo:hi].copy()))
            res = L.host(max(bnd[r + 1] - bnd[r] for r in range(parts)), shards, want_gaps=True, **kw)
            for r, (c2, f2, tot, g2) in enumerate(res):
                assert c2 == cons and (g2 == gaps).all(), (kw, r)
                if not kw:
                    assert _bits(f2[0]) == _bits(fit[0]) and _bits(f2[1]) == _bits(fit[1])
            assert (np.concatenate([s["dropped"] for s in shards]) == drop).all(), kw
            assert (np.concatenate([s["out"]["score"] for s in shards]) == ref_out["score"]).all()
            assert sum(r[2] for r in res) == tot_ref
    finally:
        for g in ctxs:
            g.close()


def test_world_one_equals_iterate_resident(gpu):
    import torch
    from mia_b200 import shard
    ref, bases, off, rc, as_, ae, seq_len, sticky = _case(20000, 3000, 8)
    cons, fit, gaps, drop = _single(gpu, ref, bases, off, rc, as_, ae, seq_len, sticky)
    gpu.set_cut_inputs(seq_len, None, sticky)
    S = shard.ShardedRounds(gpu, 0, 1, 0, len(seq_len))
    d = np.zeros(len(seq_len), np.uint8)
    c2, f2, g2 = S.resident(dropped=d, want_gaps=True)
    assert c2 == cons and (g2 == gaps).all() and (d == drop).all()
    assert _bits(f2[0]) == _bits(fit[0]) and _bits(f2[1]) == _bits(fit[1])


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import _pkg
    _pkg.load()
    from mia_b200 import api, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ref, bases, off, rc, as_, ae, seq_len, sticky = _case(60000, 5000, 12)
    n = len(seq_len)
    g = api.MiaGpu(rank)
    single = None
    if rank == 0:
        single = _single(g, ref, bases, off, rc, as_, ae, seq_len, sticky)
    lo, hi = n * rank // world, n * (rank + 1) // world
    g.set_pssm(gpu_checks.load_pssm("onepass"))
    g.set_reference(ref, circular=1, with_rc=0)
    g.upload_reads(np.ascontiguousarray(bases[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo]))
    g.set_alignment_inputs(rc[lo:hi].copy(), as_[lo:hi].copy(), ae[lo:hi].copy())
    g.set_cut_inputs(seq_len[lo:hi].copy(), None, sticky[lo:hi].copy())
    S = shard.ShardedRounds(g, rank, world, rank, (n + world - 1) // world)
    d = np.zeros(hi - lo, np.uint8)
    cons, fit, gaps = S.resident(dropped=d, want_gaps=True)
    if rank == 0:
        ok = cons == single[0] and _bits(fit[0]) == _bits(single[1][0]) and _bits(fit[1]) == _bits(single[1][1])
        ok = ok and bool((gaps == single[2]).all()) and bool((d == single[3][lo:hi]).all())
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_resident_nccl_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert ok
