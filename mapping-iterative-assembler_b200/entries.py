"""Host-side construction of the culled AlnSeq entry list (include/miagpu.h, miagpu_entry)
from realign outputs -- the "natural" case where every read points at its own fresh
segments (no stale back pointers).  Mirrors mia_main.c:259-276 (end fix + split decision),
split_pwaln (mia.c:1376-1438) and asp_len (fsdb.c:518-530).  numpy, O(n)."""
import numpy as np

from .api import ENTRY_DTYPE, MAX_RUNS


def run_geometry(runs, n_runs):
    """cols (M+D), ins (I) per read from the run lists."""
    runs = np.asarray(runs).view(np.uint16).reshape(-1, MAX_RUNS)
    valid = np.arange(MAX_RUNS)[None, :] < np.asarray(n_runs)[:, None]
    t = runs >> 14
    ln = (runs & 0x3FFF).astype(np.int64) * valid
    cols = (ln * (t != 1)).sum(1)
    ins = (ln * (t == 1)).sum(1)
    return cols.astype(np.int32), ins.astype(np.int32)


def _front_ins(runs_i, n, cols_f):
    """inserted bases attached to alignment columns < cols_f"""
    col, tot = 0, 0
    for x in runs_i[:n]:
        x = int(x)
        t, ln = x >> 14, x & 0x3FFF
        if t == 1:
            if col < cols_f:
                tot += ln
        else:
            col += ln
    return tot


def natural_entries(as_out, ae_out, n_runs, runs, seq_len, dropped_front=None, dropped_back=None, keep=None):
    """Returns (entries, split[n], slots_per_read[n]).  Reads with n_runs <= 0 or keep==0 get no entry."""
    as_out, ae_out, n_runs = (np.asarray(a) for a in (as_out, ae_out, n_runs))
    n = len(as_out)
    runs = np.asarray(runs).view(np.uint16).reshape(-1, MAX_RUNS)
    cols, ins = run_geometry(runs, n_runs)
    end = np.where(ae_out > seq_len, ae_out - seq_len, ae_out)         # mia_main.c:259-263
    split = as_out > end                                                # 265
    ok = n_runs > 0
    if keep is not None:
        ok &= np.asarray(keep).astype(bool)
    split &= ok
    nslots = ok.astype(np.int32) + split
    first = np.zeros(n + 1, np.int64)
    np.cumsum(nslots, out=first[1:])
    ent = np.zeros(int(first[-1]), ENTRY_DTYPE)
    idx = np.flatnonzero(ok)
    f = first[idx]
    tot = (cols + ins)[idx]
    ent["read"][f] = idx
    ent["col_begin"][f] = 0
    ent["col_count"][f] = cols[idx]
    ent["ref_pos"][f] = as_out[idx]
    ent["front_len"][f] = tot
    ent["total_len"][f] = tot
    if dropped_front is not None:
        ent["dropped"][f] = np.asarray(dropped_front)[idx]
    for i in np.flatnonzero(split):
        cf_raw = seq_len - int(as_out[i])
        # a read that starts beyond seq_len: split_pwaln moves ALL of it to the back AlnSeq at START 0 (mia.c:1400-1422); the front
        # AlnSeq keeps the negative length asp_len computes from end - start + 1 (fsdb.c:522-523)
        cf = min(max(cf_raw, 0), int(cols[i]))
        fi = _front_ins(runs[i], int(n_runs[i]), cf_raw)
        fl = cf_raw if cf_raw < 0 else cf + fi
        bl = (int(cols[i]) + int(ins[i])) - fl
        a, b = first[i], first[i] + 1
        ent["col_count"][a] = cf
        ent["front_len"][a] = fl
        ent["total_len"][a] = fl + bl
        ent[b] = ent[a]
        ent["col_begin"][b] = cf
        ent["col_count"][b] = int(cols[i]) - cf
        ent["ref_pos"][b] = 0
        ent["back_formula"][b] = 1
        if dropped_back is not None:
            ent["dropped"][b] = dropped_back[i]
    return ent, split, nslots
