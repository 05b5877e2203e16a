"""Diagnostic: the block records of a sharded round (protocol 2) against the chains added read by read in numpy.
usage: python scripts/gpu_shard_diag.py [reads_per_shard] [shards]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg  # noqa: E402

_pkg.load()
import torch  # noqa: E402
from mia_b200 import api, driver, shard, synth  # noqa: E402
import gpu_checks  # noqa: E402


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 700_000
    parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ref = synth.random_reference(16569, seed=1)
    genome = synth.diverge(ref, 0.005, seed=3, indel_rate=0.0)
    sm = gpu_checks.load_pssm("pe")
    ctxs = [api.MiaGpu(0) for _ in range(parts)]
    asms = [driver.ResidentAssembler(g, ref, sm, circular=1, k=12, strand_unknown="drop", pointer_state=False) for g in ctxs]
    for r, a in enumerate(asms):
        b, o = synth.make_reads(genome, per - 13000 * r, 30, 140, seed=1000 + r, circular=True)[:2]
        a.pass1(b, o, defer_cull=True)
    all_sl = np.concatenate([a.seq_len for a in asms])
    all_sc = np.concatenate([a.score for a in asms])
    for a in asms:
        a.pass1_cull(all_sl, all_sc)
    n_max = max(len(a.seq_len) for a in asms)
    print("reads per shard", [len(a.seq_len) for a in asms], "n_max", n_max)
    L = shard.LocalShards(ctxs)
    rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    for rnd in range(rounds):
        print('round', rnd + 1)
        for a in asms:
            a.begin_round()
        L._exchange_begin([g.shard_begin(L.world, r, n_max) for r, g in enumerate(ctxs)])
        fs = [g.shard_fit() for g in ctxs]
        L._exchange_fit(fs)
        words = fs[0]["send"][1]
        nb = (n_max + 511) // 512
        rec_dt = np.dtype([("T", "<f8", 2), ("A", "<f8", 2), ("e", "<i4", 2), ("ok", "<i4", 2)])
        recv = shard.dev_tensor(fs[0]["recv"][0], words * parts, L.device).cpu().numpy()
        # the reads of every shard as the device has them now
        sl, sc = [], []
        for g, a in zip(ctxs, asms):
            al = g.get_alignment()
            sl.append(a.seq_len.astype(np.int64))
            sc.append(al["score"].astype(np.int64))
        used = [s >= 2000 for s in sc]
        sx = sum(int(l[u].sum()) for l, u in zip(sl, used))
        sy = sum(int(s[u].sum()) for s, u in zip(sc, used))
        cnt = sum(int(u.sum()) for u in used)
        xbar, ybar = sx / cnt, sy / cnt
        run = np.zeros(2)
        for r in range(parts):
            recs = recv[r * words: r * words + nb * rec_dt.itemsize // 4].view(rec_dt)
            pf = recv[r * words + nb * rec_dt.itemsize // 4:]
            dx = sl[r].astype(np.float64) - xbar
            a0 = np.where(used[r], dx * (sc[r].astype(np.float64) - ybar), 0.0)
            a1 = np.where(used[r], dx * dx, 0.0)
            for ch, a in enumerate((a0, a1)):
                c = np.cumsum(np.concatenate([[run[ch]], a]))       # sequential adds
                starts = c[0:len(a) + 1:512][:nb]
                starts = np.concatenate([starts, np.full(nb - len(starts), c[-1])])
                e_true = np.frexp(starts)[1] - 1
                ok = recs["ok"][:, ch]
                e = recs["e"][:, ch]
                mism = np.flatnonzero((e != e_true) & (ok != 0))
                print(f"rank {r} chain {ch}: blocks {nb}, not ok {int((ok == 0).sum())}, e mismatches among ok {len(mism)}, start {run[ch]:.6g}, end {c[-1]:.6g}, suspects {int(pf[0])}")
                for b in mism[:5]:
                    print("   block", int(b), "e", int(e[b]), "true e", int(e_true[b]), "true start", starts[b], "T", recs["T"][b, ch], "A", recs["A"][b, ch])
                run[ch] = c[-1]
        try:
            cuts = [g.shard_cut() for g in ctxs]
            print("shard_cut ok", cuts[0][0], [g.last_cut_stats() for g in ctxs])
            L._exchange_cut([c[1] for c in cuts])
            for g, a in zip(ctxs, asms):
                cons, gaps, _ = g.shard_finish(1, a.dropped, None, False)
                a.end_round(cons, cuts[0][0], gaps)
        except api.MiaGpuError as ex:
            print("shard_cut FAILED:", ex)
            break
    for g in ctxs:
        g.close()


if __name__ == "__main__":
    main()
