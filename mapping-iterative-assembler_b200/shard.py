"""Sharded rounds: reads partitioned over the GPUs of one box, one process and one
libmiagpu context per GPU, consensus replicated.  The library has no communication dependency; this
module runs the three collectives of a round (all-gather of the regression keys, all-reduce MAX of
the insert maxima + per-length best scores, all-reduce SUM of the column planes) with
torch.distributed / NCCL on the library's own stream, so that a round is one stream-ordered
sequence without host synchronisation between the DP and the collectives.  A C host makes the same
calls with ncclAllGather / ncclAllReduce.

`LocalShards` runs the same protocol for several contexts that live in ONE process on one GPU and
emulates the collectives with device copies: the cross-rank logic is then testable on a single GPU.
"""
import numpy as np


class _Raw:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3}


def dev_tensor(ptr, words, device, typestr="<i4"):
    """torch view of `words` 32-bit words of device memory owned by the library"""
    import torch
    return torch.as_tensor(_Raw(ptr, words, typestr), device=device)


class ShardedRounds:
    """One rank's side of the protocol.  gpu: api.MiaGpu; group: torch.distributed process group (None = default)."""

    def __init__(self, gpu, device, world, rank, n_max, group=None):
        import torch
        self.g, self.world, self.rank, self.n_max, self.group = gpu, world, rank, int(n_max), group
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.stream = torch.cuda.ExternalStream(gpu.lib.miagpu_stream(gpu.h), device=self.device)

    # the collectives, ordered on the library's stream
    def _after_begin(self, b):
        import torch
        import torch.distributed as dist
        send = dev_tensor(b["send"][0], b["send"][1], self.device)
        recv = dev_tensor(b["recv"][0], b["recv"][1], self.device)
        mx = dev_tensor(b["max"][0], b["max"][1], self.device)
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                recv.copy_(send)
            else:
                dist.all_gather_into_tensor(recv, send, group=self.group)
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)

    def _after_cut(self, sb):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            planes = dev_tensor(sb[0], sb[1], self.device)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(planes, op=dist.ReduceOp.SUM, group=self.group)

    def resident(self, cons_code=1, hard_cut=0, score_cut=None, dropped=None, want_gaps=False):
        """miagpu_iterate_resident for a shard: -> (consensus, (slope, intercept), gaps)"""
        b = self.g.shard_begin(self.world, self.rank, self.n_max, hard_cut, score_cut)
        self._after_begin(b)
        fit, sb = self.g.shard_cut()
        self._after_cut(sb)
        cons, gaps, _ = self.g.shard_finish(cons_code, dropped, None, want_gaps)
        return cons, fit, gaps

    def host(self, bases, offsets, rc, as_, ae, seq_len, dropped, out, packed=None, cons_code=1, unique_best=None, hard_cut=0,
             score_cut=None, want_gaps=False):
        """miagpu_iterate_host for a shard: -> (consensus, (slope, intercept), total_runs, gaps); dropped updated in place"""
        b = self.g.shard_begin_host(self.world, self.rank, self.n_max, bases, offsets, rc, as_, ae, seq_len, dropped, out, unique_best,
                                    hard_cut, score_cut)
        self._after_begin(b)
        fit, sb = self.g.shard_cut()
        self._after_cut(sb)
        cons, gaps, tot = self.g.shard_finish(cons_code, dropped, packed, want_gaps, want_total_runs=True)
        return cons, fit, tot, gaps


class LocalShards:
    """Several contexts in one process; the collectives are emulated with torch ops between device-wide syncs."""

    def __init__(self, gpus, device=0):
        import torch
        self.gpus, self.device = gpus, torch.device("cuda", device)
        self.world = len(gpus)

    def _sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    def _exchange_begin(self, bs):
        import torch
        self._sync()
        sends = [dev_tensor(b["send"][0], b["send"][1], self.device) for b in bs]
        allk = torch.cat(sends)
        for b in bs:
            dev_tensor(b["recv"][0], b["recv"][1], self.device).copy_(allk)
        mx = [dev_tensor(b["max"][0], b["max"][1], self.device) for b in bs]
        m = torch.stack(mx).max(0).values
        for t in mx:
            t.copy_(m)
        self._sync()

    def _exchange_cut(self, sbs):
        import torch
        self._sync()
        ts = [dev_tensor(sb[0], sb[1], self.device) for sb in sbs]
        total = torch.stack(ts).sum(0, dtype=torch.int32)
        for t in ts:
            t.copy_(total)
        self._sync()

    def resident(self, n_max, cons_code=1, hard_cut=0, score_cut=None, dropped=None, want_gaps=False):
        """dropped: list of uint8 arrays (one per shard) or None.  -> list of (consensus, fit, gaps) per shard"""
        bs = [g.shard_begin(self.world, r, n_max, hard_cut, score_cut) for r, g in enumerate(self.gpus)]
        self._exchange_begin(bs)
        cuts = [g.shard_cut() for g in self.gpus]
        self._exchange_cut([c[1] for c in cuts])
        res = []
        for r, g in enumerate(self.gpus):
            cons, gaps, _ = g.shard_finish(cons_code, None if dropped is None else dropped[r], None, want_gaps)
            res.append((cons, cuts[r][0], gaps))
        return res

    def host(self, n_max, shards, cons_code=1, hard_cut=0, score_cut=None, want_gaps=False):
        """shards: list of dicts(bases, off, rc, as_, ae, seq_len, dropped, out, packed=None, unique_best=None)"""
        bs = [g.shard_begin_host(self.world, r, n_max, s["bases"], s["off"], s["rc"], s["as_"], s["ae"], s["seq_len"], s["dropped"], s["out"],
                                 s.get("unique_best"), hard_cut, score_cut) for r, (g, s) in enumerate(zip(self.gpus, shards))]
        self._exchange_begin(bs)
        cuts = [g.shard_cut() for g in self.gpus]
        self._exchange_cut([c[1] for c in cuts])
        res = []
        for r, (g, s) in enumerate(zip(self.gpus, shards)):
            cons, gaps, tot = g.shard_finish(cons_code, s["dropped"], s.get("packed"), want_gaps, want_total_runs=True)
            res.append((cons, cuts[r][0], tot, gaps))
        return res


def split_reads(bases, off, parts):
    """contiguous partition in input (FSDB) order: -> list of (lo, hi, bases_slice, off_slice)"""
    n = len(off) - 1
    out = []
    for r in range(parts):
        lo, hi = n * r // parts, n * (r + 1) // parts
        out.append((lo, hi, np.ascontiguousarray(bases[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo])))
    return out
