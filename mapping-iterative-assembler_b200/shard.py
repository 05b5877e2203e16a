"""Sharded rounds: reads partitioned over the GPUs of one box, one process and one
libmiagpu context per GPU, consensus replicated.  The library has no communication dependency; this
module runs the three collectives of a round (all-reduce MAX of the insert maxima + per-length best
scores + the ranks' integer sums, all-gather of the regression's block records, all-reduce SUM of the
column planes) with
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


_views = {}


def dev_tensor(ptr, words, device, typestr="<i4"):
    """torch view of `words` 32-bit words of device memory owned by the library (views are cached: a round hands out the same
    buffers again and again, and building a view costs tens of microseconds on the round's critical path)"""
    import torch
    key = (ptr, words, str(device), typestr)
    t = _views.get(key)
    if t is None:
        if len(_views) > 256:
            _views.clear()
        t = _views[key] = torch.as_tensor(_Raw(ptr, words, typestr), device=device)
    return t


class ShardedRounds:
    """One rank's side of the protocol.  gpu: api.MiaGpu; group: torch.distributed process group (None = default)."""

    def __init__(self, gpu, device, world, rank, n_max, group=None):
        import torch
        self.g, self.world, self.rank, self.n_max, self.group = gpu, world, rank, int(n_max), group
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.stream = torch.cuda.ExternalStream(gpu.lib.miagpu_stream(gpu.h), device=self.device)

    # the collectives, ordered on the library's stream
    def _after_begin(self, mb):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            mx = dev_tensor(mb[0], mb[1], self.device)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)

    def _after_fit(self, f):
        import torch
        import torch.distributed as dist
        words = f["send"][1]
        if not words:
            return
        send = dev_tensor(f["send"][0], words, self.device)
        recv = dev_tensor(f["recv"][0], words * self.world, self.device)
        with torch.cuda.stream(self.stream):
            if self.world == 1:
                recv.copy_(send)
            else:
                dist.all_gather_into_tensor(recv, send, group=self.group)

    def _after_cut(self, sb):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            planes = dev_tensor(sb[0], sb[1], self.device)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(planes, op=dist.ReduceOp.SUM, group=self.group)

    def _round(self):
        self._after_fit(self.g.shard_fit())
        fit, sb = self.g.shard_cut()
        self._after_cut(sb)
        return fit

    def _after_finish(self):
        """pointer state: the slot flags of all ranks (MAX), before the next round begins"""
        import torch
        import torch.distributed as dist
        ptr, nbytes = self.g.shard_flags()
        if nbytes and self.world > 1:
            fl = dev_tensor(ptr, nbytes, self.device, "|u1")
            with torch.cuda.stream(self.stream):
                dist.all_reduce(fl, op=dist.ReduceOp.MAX, group=self.group)

    def resident(self, cons_code=1, hard_cut=0, score_cut=None, dropped=None, want_gaps=False):
        """miagpu_iterate_resident for a shard: -> (consensus, (slope, intercept), gaps)"""
        self._after_begin(self.g.shard_begin(self.world, self.rank, self.n_max, hard_cut, score_cut))
        fit = self._round()
        cons, gaps, _ = self.g.shard_finish(cons_code, dropped, None, want_gaps)
        self._after_finish()
        return cons, fit, gaps

    def host(self, bases, offsets, rc, as_, ae, seq_len, dropped, out, packed=None, cons_code=1, unique_best=None, hard_cut=0,
             score_cut=None, want_gaps=False):
        """miagpu_iterate_host for a shard: -> (consensus, (slope, intercept), total_runs, gaps); dropped updated in place"""
        self._after_begin(self.g.shard_begin_host(self.world, self.rank, self.n_max, bases, offsets, rc, as_, ae, seq_len, dropped, out,
                                                  unique_best, hard_cut, score_cut))
        fit = self._round()
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

    def _exchange_begin(self, mbs):
        import torch
        self._sync()
        mx = [dev_tensor(mb[0], mb[1], self.device) for mb in mbs]
        m = torch.stack(mx).max(0).values
        for t in mx:
            t.copy_(m)
        self._sync()

    def _exchange_fit(self, fs):
        import torch
        self._sync()
        if fs[0]["send"][1]:
            sends = [dev_tensor(f["send"][0], f["send"][1], self.device) for f in fs]
            allk = torch.cat(sends)
            for f in fs:
                dev_tensor(f["recv"][0], f["send"][1] * self.world, self.device).copy_(allk)
        self._sync()

    def _exchange_cut(self, sbs):
        import torch
        self._sync()
        ts = [dev_tensor(sb[0], sb[1], self.device) for sb in sbs]
        total = torch.stack(ts).sum(0, dtype=torch.int32)
        for t in ts:
            t.copy_(total)
        self._sync()

    def _round(self):
        self._exchange_fit([g.shard_fit() for g in self.gpus])
        cuts = [g.shard_cut() for g in self.gpus]
        self._exchange_cut([c[1] for c in cuts])
        return cuts

    def _exchange_flags(self):
        import torch
        fl = [g.shard_flags() for g in self.gpus]
        if not fl[0][1]:
            return
        self._sync()
        ts = [dev_tensor(p, n, self.device, "|u1") for p, n in fl]
        m = torch.stack(ts).max(0).values
        for t in ts:
            t.copy_(m)
        self._sync()

    def resident(self, n_max, cons_code=1, hard_cut=0, score_cut=None, dropped=None, want_gaps=False):
        """dropped: list of uint8 arrays (one per shard) or None.  -> list of (consensus, fit, gaps) per shard"""
        self._exchange_begin([g.shard_begin(self.world, r, n_max, hard_cut, score_cut) for r, g in enumerate(self.gpus)])
        cuts = self._round()
        res = []
        for r, g in enumerate(self.gpus):
            cons, gaps, _ = g.shard_finish(cons_code, None if dropped is None else dropped[r], None, want_gaps)
            res.append((cons, cuts[r][0], gaps))
        self._exchange_flags()
        return res

    def host(self, n_max, shards, cons_code=1, hard_cut=0, score_cut=None, want_gaps=False):
        """shards: list of dicts(bases, off, rc, as_, ae, seq_len, dropped, out, packed=None, unique_best=None)"""
        self._exchange_begin([g.shard_begin_host(self.world, r, n_max, s["bases"], s["off"], s["rc"], s["as_"], s["ae"], s["seq_len"], s["dropped"],
                                                 s["out"], s.get("unique_best"), hard_cut, score_cut) for r, (g, s) in enumerate(zip(self.gpus, shards))])
        cuts = self._round()
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
