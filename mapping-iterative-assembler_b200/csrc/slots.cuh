// slots.cuh -- the reference's FragSeq -> AlnSeq pointer semantics on the device (sm_100a).
//
// In the reference an AlnSeq is an OBJECT in maln->AlnSeqArray[k] ("slot" k) that every round re-uses for the k-th merged
// segment in FSDB order (merge_pwaln_into_maln, map_align.c:866-954: one slot per read, two per wrap-split read), and a
// FragSeq holds POINTERS to its slots:
//   * AlnSeq.dropped lives in the slot and is only ever set (cull_maln_from_fsdb mia.c:471-478; merge copies every field
//     except it, map_align.c:885-893; H10): when the split pattern of the reads changes, every later read slides onto slots
//     whose flags other reads set;
//   * reiterate_assembly sets front_asp always but back_asp only when the read is wrap-split this round
//     (mia_main.c:265-276): a read that was split once keeps a STALE pointer to its old back slot;
//   * a read whose strand is unknown (pass-1 score <= 2000: mia.c:1653; -D accepts those, mia.c:1614) is not realigned at
//     all (mia_main.c:178), both its pass-1 pointers stay;
//   * pop_smp_from_FSDB (fsdb.c:542-619), cull_maln_from_fsdb (mia.c:418-506) and the consensus (mia.c:515-603) follow the
//     pointers: a stale pointer to a slot that is live this round counts that slot's alignment once more (and the slot's
//     smp codes are whatever the LAST pointer visited wrote); a stale pointer to a slot beyond this round's slot count sees
//     the content an earlier round left there, with its inserts freed (mia_main.c:80-92).
// Here: slot numbers come from an exclusive scan over the reads (fs_nsl_kernel + cub), the natural entry of every fresh
// segment is built by fs_entries_kernel (which also emits the list of stale pointers), flags are indexed by slot, and the
// few stale pointers are resolved by the host part (fs_resolve in miagpu.cu) into extra entries / patched smp parameters;
// content that is no longer live is kept as "frozen" alignments (fs_freeze_kernel) taken from the previous round's
// results, which stay in the ping-pong partners of the per-read alignment buffers.
#pragma once
#include "common.cuh"
#include "consensus.cuh"
#include "scorecut.cuh"

namespace miagpu {

constexpr int FS_CNT_STALE = 0, FS_CNT_STATUS = 1, FS_CNT_NSLOTS = 2, FS_CNT_WORDS = 8;
constexpr int FS_CNT_BASE = 3, FS_CNT_TOTAL = 4;   // sharded rounds: first slot of this rank's reads, slots of all ranks (this round)
constexpr int FZ_BASES = MAX_READ;                  // bytes of bases per frozen alignment

struct FsDev {
  const uint8_t* known;        // FragSeq.strand_known
  int32_t* front_slot;         // FragSeq.front_asp as a slot index
  int32_t* back_slot;          // FragSeq.back_asp, -1 = NULL
  uint8_t* slot_flag;          // AlnSeq.dropped per slot, sticky
  uint8_t* slot_new;           // flags this round's cull sets (committed by fs_commit_kernel)
  const int32_t* first;        // slot of the read's front segment this round (exclusive scan of nsl)
  int32_t* slot_owner;         // slot -> 2 * read + segment, this round
  int32_t* ent_slot;           // entry -> slot (-1: the entry is empty)
  int32_t* stale;              // [stale_cap][3]: read, kind (0 front_asp, 1 back_asp), slot
  int32_t* counters;           // FS_CNT_*
  int32_t stale_cap;
  const int32_t* slot_base;    // nullable; sharded rounds: *slot_base = the slot of this rank's first AlnSeq (slot numbers are global:
                               // the ranks' reads one after the other in FSDB order)
};

// Columns / inserted bases / deletions of an alignment and of its part in front of the wrap point
// (mia_main.c:259-276: end fix + split test; split_pwaln mia.c:1376-1438; asp_len fsdb.c:518-530).
struct SegGeom { int split, cols, ins, dels, fcols, fins, fdels, cf; };
__device__ __forceinline__ SegGeom seg_geom(int start, int ae, int seq_len, const uint16_t* r, int nr) {
  SegGeom g{};
  int end = ae;
  if (end > seq_len) end -= seq_len;
  g.split = start > end;
  g.cf = g.split ? seq_len - start : 0x7fffffff;    // alignment columns that stay in front; negative when the read starts beyond seq_len
  for (int k = 0; k < nr; k++) {
    const int x = r[k], t = x >> 14, len = x & 0x3fff;
    if (t == MIAGPU_RUN_I) { g.ins += len; if (g.cols < g.cf) g.fins += len; }
    else {
      if (t == MIAGPU_RUN_D) { g.dels += len; g.fdels += min(max(g.cf - g.cols, 0), len); }
      g.cols += len;
    }
  }
  g.fcols = g.split ? min(max(g.cf, 0), g.cols) : g.cols;
  if (!g.split) { g.fins = g.ins; g.fdels = g.dels; }
  return g;
}
// the natural entries of a read's alignment: 2i = whole / front, 2i + 1 = back (col_count 0 unless split)
__device__ __forceinline__ void seg_entries(int rd, int start, const SegGeom& g, miagpu_entry& f, miagpu_entry& b) {
  f = miagpu_entry{}; b = miagpu_entry{};
  f.read = b.read = rd;
  // asp_len goes by end - start + 1: a front AlnSeq that starts beyond seq_len has a NEGATIVE length (fsdb.c:522-523)
  const int fl = (g.split && g.cf < 0) ? g.cf : g.fcols + g.fins;
  f.col_begin = 0; f.col_count = g.fcols; f.ref_pos = start; f.front_len = fl; f.total_len = g.cols + g.ins;
  if (g.split) {
    b = f;
    b.col_begin = g.fcols; b.col_count = g.cols - g.fcols; b.ref_pos = 0; b.back_formula = 1;
  }
}

// slots the read takes this round: known reads merge one AlnSeq, two when wrap-split; strand-unknown reads none
__global__ void fs_nsl_kernel(int64_t n, const uint8_t* __restrict__ known, const int32_t* __restrict__ as_in, const int32_t* __restrict__ ae_in,
                              int32_t* as_out, int32_t* ae_out, const int32_t* __restrict__ n_runs, const uint8_t* __restrict__ status,
                              int seq_len, int32_t* nsl, int32_t* counters) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int st = 0;
  if (i < n) {
    int v = 0;
    if (!known || known[i]) {
      st = status[i] | (n_runs[i] <= 0 ? MIAGPU_ST_RUNS_OVERFLOW : 0);
      int end = ae_out[i];
      if (end > seq_len) end -= seq_len;
      v = 1 + (as_out[i] > end);
    } else {                                         // not realigned: fs->as / fs->ae stay what they were
      as_out[i] = as_in[i]; ae_out[i] = ae_in[i];
    }
    nsl[i] = v;
  }
  st = __reduce_or_sync(0xffffffffu, st);
  if ((threadIdx.x & 31) == 0 && st) atomicOr(&counters[FS_CNT_STATUS], st);
}

// natural entries + slot bookkeeping of one round (see the header).  unique (nullable): a read that is not unique_best keeps
// its slots but is absent from the culled list (mia.c:466).
__global__ void fs_entries_kernel(int64_t n, FsDev f, const int32_t* __restrict__ as_out, const int32_t* __restrict__ ae_out,
                                  const int32_t* __restrict__ n_runs, const uint16_t* __restrict__ runs, int seq_len,
                                  const uint8_t* __restrict__ unique, miagpu_entry* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  miagpu_entry ef{}, eb{};
  ef.read = eb.read = (int32_t)i;
  int sf = -1, sb = -1;
  const int old_back = f.back_slot[i];
  const bool listed = !unique || unique[i];
  auto stale = [&](int kind, int slot) {
    const int q = atomicAdd(&f.counters[FS_CNT_STALE], 1);
    if (q < f.stale_cap) { f.stale[3 * q] = (int32_t)i; f.stale[3 * q + 1] = kind; f.stale[3 * q + 2] = slot; }
  };
  if (!f.known || f.known[i]) {
    const int nr = n_runs[i];
    const int s0 = f.first[i] + (f.slot_base ? *f.slot_base : 0);
    const SegGeom g = seg_geom(as_out[i], ae_out[i], seq_len, runs + i * MAX_RUNS, nr > 0 ? nr : 0);
    seg_entries((int)i, as_out[i], g, ef, eb);
    f.front_slot[i] = s0;
    f.slot_owner[s0] = (int32_t)(2 * i);
    ef.dropped = f.slot_flag[s0];
    sf = s0;
    if (g.split) {
      f.back_slot[i] = s0 + 1;
      f.slot_owner[s0 + 1] = (int32_t)(2 * i + 1);
      eb.dropped = f.slot_flag[s0 + 1];
      sb = s0 + 1;
    } else if (old_back >= 0) {
      stale(1, old_back);                            // mia_main.c:273-276: back_asp is not cleared
    }
    if (!listed) { ef.col_count = 0; eb.col_count = 0; sf = sb = -1; }
  } else {                                           // mia_main.c:178: not realigned, both pass-1 pointers stay
    stale(0, f.front_slot[i]);
    if (old_back >= 0) stale(1, old_back);
  }
  out[2 * i] = ef; out[2 * i + 1] = eb;
  f.ent_slot[2 * i] = sf; f.ent_slot[2 * i + 1] = sb;
}

// What a stale pointer points at, for the host's resolution.  rec[8 * q ..]: read, kind, slot | live owner (2 * read + seg, -1) |
// previous round's owner (-1) | holder known | holder's front_asp | spare.  geo[8 * q ..]: cb, cc, ref_pos, ins, dels, rb of the LIVE content and the
// natural smp parameters are taken from the entries by the host (it downloads the owners' entries); here only the owners.
__global__ void fs_gather_kernel(int n_stale, const int32_t* __restrict__ stale, const int32_t* __restrict__ slot_owner, int64_t n_slots,
                                 const int32_t* __restrict__ slot_owner_prev, int64_t n_slots_prev, const uint8_t* __restrict__ known,
                                 const int32_t* __restrict__ front_slot, int32_t* rec) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_stale) return;
  const int i = stale[3 * q], kind = stale[3 * q + 1], k = stale[3 * q + 2];
  rec[8 * q] = i; rec[8 * q + 1] = kind; rec[8 * q + 2] = k;
  rec[8 * q + 3] = (k >= 0 && k < n_slots) ? slot_owner[k] : -1;
  rec[8 * q + 4] = (k >= 0 && k < n_slots_prev && slot_owner_prev) ? slot_owner_prev[k] : -1;
  rec[8 * q + 5] = (!known || known[i]) ? 1 : 0;
  rec[8 * q + 6] = front_slot[i];
  rec[8 * q + 7] = (k >= 0 && k < n_slots) ? 1 : 0;      // live this round (sharded rounds: possibly on another rank, then rec[3] = -1)
}

// Geometry of the segments of a list of alignments (2 * read + seg), from the current or the previous round's results:
// out[8 * q ..] = cb, cc, ref_pos, ins (attached to those columns), dels, rb (read bases consumed before cb), front_len, total_len
struct AlnView { const int32_t* as_out; const int32_t* ae_out; const int32_t* n_runs; const uint16_t* runs; const int32_t* abr; int seq_len; };
__global__ void fs_geom_kernel(int m, const int32_t* __restrict__ which, AlnView v, int32_t* out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int e = which[q], j = e >> 1, seg = e & 1;
  const int nr = v.n_runs[j];
  const SegGeom g = seg_geom(v.as_out[j], v.ae_out[j], v.seq_len, v.runs + (int64_t)j * MAX_RUNS, nr > 0 ? nr : 0);
  miagpu_entry ef, eb;
  seg_entries(j, v.as_out[j], g, ef, eb);
  const miagpu_entry& x = seg ? eb : ef;
  int32_t* o = out + 8 * q;
  o[0] = x.col_begin; o[1] = x.col_count; o[2] = x.ref_pos;
  o[3] = seg ? g.ins - g.fins : g.fins;
  o[4] = seg ? g.dels - g.fdels : g.fdels;
  o[5] = seg ? g.fcols - g.fdels + g.fins : 0;
  o[6] = x.front_len; o[7] = x.total_len;
}

// Freeze the content of segments (2 * read + seg of the view) as stand-alone alignments without their inserts (the inserts
// of every slot are freed when the next round begins, mia_main.c:80-92): bases = the aligned read bases of the segment's
// M columns, runs = its M / D runs, first row 0.  flip (nullable, per read): the stored read is still in its original
// orientation although the alignment is a reverse-strand one (strand-unknown reads, fsdb.c:209-210): the bases are
// reverse-complemented while they are copied.  dst[q] = index of the frozen alignment that receives list item q.
__global__ void fs_freeze_kernel(int m, const int32_t* __restrict__ which, const int32_t* __restrict__ dst, AlnView v,
                                 const uint8_t* __restrict__ bases, const int64_t* __restrict__ off, const uint8_t* __restrict__ rc,
                                 const uint8_t* __restrict__ flip, const int32_t* __restrict__ score_prev, uint8_t* fz_bases, uint16_t* fz_runs,
                                 int32_t* fz_nruns, uint8_t* fz_rc, int32_t* geo) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int e = which[q], j = e >> 1, seg = e & 1, d = dst[q];
  const int nr = max(v.n_runs[j], 0);
  const uint16_t* r = v.runs + (int64_t)j * MAX_RUNS;
  const SegGeom g = seg_geom(v.as_out[j], v.ae_out[j], v.seq_len, r, nr);
  const int cb = seg ? g.fcols : 0, ce = seg ? g.cols : g.fcols;
  const uint8_t* read = bases + off[j];
  const int L = (int)(off[j + 1] - off[j]);
  const bool fl = flip && flip[j];
  uint8_t* ob = fz_bases + (int64_t)d * FZ_BASES;
  uint16_t* orun = fz_runs + (int64_t)d * MAX_RUNS;
  int col = 0, row = v.abr[j], nb = 0, no = 0, dels = 0;
  for (int k = 0; k < nr; k++) {
    const int x = r[k], t = x >> 14, len = x & 0x3fff;
    if (t == MIAGPU_RUN_I) { row += len; continue; }
    const int lo = max(col, cb), hi = min(col + len, ce);
    if (hi > lo) {
      if (t == MIAGPU_RUN_M)
        for (int c = lo; c < hi; c++) {
          const int rr = row + (c - col);
          uint8_t b = 'N';
          if (rr >= 0 && rr < L) {
            b = fl ? read[L - 1 - rr] : read[rr];
            if (fl) b = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : b == 'T' ? 'A' : 'N';
          }
          if (nb < FZ_BASES) ob[nb++] = b;
        }
      else dels += hi - lo;
      if (no > 0 && (orun[no - 1] >> 14) == t) orun[no - 1] = (uint16_t)((t << 14) | ((orun[no - 1] & 0x3fff) + (hi - lo)));
      else if (no < MAX_RUNS) orun[no++] = (uint16_t)((t << 14) | (hi - lo));
    }
    col += len;
    if (t == MIAGPU_RUN_M) row += len;
  }
  fz_nruns[d] = no;
  fz_rc[d] = rc[j];
  int32_t* o = geo + 8 * q;
  o[0] = seg ? 0 : v.as_out[j];                      // AlnSeq.start
  o[1] = ce - cb;                                    // columns
  o[2] = dels;
  o[3] = rc[j];
  o[4] = score_prev ? score_prev[j] : 0;             // AlnSeq.score
  o[5] = g.split ? (seg ? 'b' : 'f') : 'a';          // AlnSeq.segment
  o[6] = j; o[7] = 0;
}

// patches of the host's resolution: pat[6 * q ..] = entry, front_len, total_len, act_bias, back_formula, slot (ent_slot; -2 = keep)
__global__ void fs_patch_kernel(int m, const int32_t* __restrict__ pat, miagpu_entry* entries, int32_t* ent_slot, const uint8_t* __restrict__ slot_flag) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const int32_t* p = pat + 6 * q;
  miagpu_entry& e = entries[p[0]];
  e.front_len = p[1]; e.total_len = p[2]; e.act_bias = p[3]; e.back_formula = (uint8_t)p[4];
  if (p[5] != -2) { ent_slot[p[0]] = p[5]; e.dropped = p[5] >= 0 ? slot_flag[p[5]] : 0; }
}

// cull_maln_from_fsdb's per-read test (mia.c:452-479) through the pointers: a read below the cut flags whatever its front_asp
// and back_asp point at.  len_for_thr: seq_len, or find_alignable_len under -D (nprefix = number of 'N' in ref[0, x), nullable).
__global__ void fs_flags_kernel(int64_t n, const int32_t* __restrict__ seq_len, const int32_t* __restrict__ score, const double* __restrict__ thr,
                                const uint8_t* __restrict__ unique, FsDev f, const int32_t* __restrict__ as_out, const int32_t* __restrict__ ae_out,
                                const int32_t* __restrict__ nprefix, int wrap_len, CutStatsDev* st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int l = seq_len[i];
  if (l < 0 || l > MAX_READ) { atomicMin(&st->bad, (long long)i); return; }
  if (unique && !unique[i]) return;
  if (nprefix) {                                     // mia.c:69-91
    const long long a = as_out[i], e = min((long long)ae_out[i], (long long)wrap_len);
    if (a >= 0 && e > a) l -= nprefix[e] - nprefix[a];
    l = max(l, 15);                                  // MIN_ALIGNABLE_LEN, params.h:38
  }
  if (!((double)score[i] < thr[l])) return;
  const int kf = f.front_slot[i], kb = f.back_slot[i];
  if (kf >= 0 && !f.slot_flag[kf]) f.slot_new[kf] = 1;
  if (kb >= 0 && !f.slot_flag[kb]) f.slot_new[kb] = 1;
}

// entries whose slot this round's cull flagged: their base columns come back out of the planes (they were accumulated while the
// host was still stitching the regression) and they take the flag.  A lane looks at one entry, the warp walks the flagged ones.
__global__ void __launch_bounds__(256) fs_undo_kernel(ConsParams p, const int32_t* __restrict__ ent_slot, const uint8_t* __restrict__ slot_new,
                                                      miagpu_entry* entries) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
  const int64_t idx = w0 + lane;
  bool want = false;
  if (idx < p.n_entries) {
    const int s = ent_slot[idx];
    want = s >= 0 && slot_new[s] && !entries[idx].dropped && entries[idx].col_count > 0;
    if (s >= 0 && slot_new[s]) entries[idx].dropped = 1;
  }
  unsigned m = __ballot_sync(0xffffffffu, want);
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    miagpu_entry e = entries[w0 + b];
    e.dropped = 0;
    walk_entry<2>(p, e, lane, NegGlobalAdder{p.acc, p.n_cols});
  }
}

// flags of this round become sticky; per read (nullable): AlnSeq.dropped of what its front_asp points at
__global__ void fs_commit_kernel(int64_t n_slots, uint8_t* slot_flag, uint8_t* slot_new) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n_slots && slot_new[k]) { slot_flag[k] = 1; slot_new[k] = 0; }
}
__global__ void fs_read_flags_kernel(int64_t n, const int32_t* __restrict__ front_slot, const int32_t* __restrict__ back_slot,
                                     const uint8_t* __restrict__ slot_flag, uint8_t* dropped_front, uint8_t* dropped_back) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kf = front_slot[i], kb = back_slot[i];
  if (dropped_front) dropped_front[i] = kf >= 0 ? slot_flag[kf] : 0;
  if (dropped_back) dropped_back[i] = kb >= 0 ? slot_flag[kb] : 0;
}

// sharded rounds: the ranks' slot counts of this round came back in the header rows of the MAX all-reduce (word 11 of a row)
__global__ void fs_base_kernel(int world, int rank, const int32_t* __restrict__ hdr, int hdr_stride, int32_t* counters) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int base = 0, total = 0;
  for (int r = 0; r < world; r++) { const int v = hdr[r * hdr_stride + 11]; if (r < rank) base += v; total += v; }
  counters[FS_CNT_BASE] = base; counters[FS_CNT_TOTAL] = total;
}

// legacy per-read flags (miagpu_set_cut_inputs) become the flags of the slots the reads take at the first numbering
__global__ void fs_seed_flags_kernel(int64_t n, const int32_t* __restrict__ first, const int32_t* __restrict__ nsl, const uint8_t* __restrict__ read_flag,
                                     uint8_t* slot_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !read_flag[i]) return;
  for (int s = 0; s < nsl[i]; s++) slot_flag[first[i] + s] = 1;
}

// slot -> owner of the pass-1 numbering the host hands over with miagpu_set_fsdb
__global__ void fs_owner_kernel(int64_t n, const int32_t* __restrict__ front_slot, const int32_t* __restrict__ back_slot, int64_t n_slots, int32_t* owner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int kf = front_slot[i], kb = back_slot[i];
  if (kf >= 0 && kf < n_slots) owner[kf] = (int32_t)(2 * i);
  if (kb >= 0 && kb < n_slots) owner[kb] = (int32_t)(2 * i + 1);
}

// pass-1 results of the reads that stay (miagpu_compact_reads), in FSDB order, as "the previous round": PWAlnFrag start / end
// after the end fix (mia.c:1606-1610), runs in forward reference orientation; a reverse-strand alignment begins at row 0 of the
// reverse-complemented read (the soft clip is at its far end).  flip[i]: the stored read was NOT reverse-complemented.
__global__ void fs_prev_from_pass1_kernel(int64_t m, const int32_t* __restrict__ src, const uint8_t* __restrict__ revcomp, const int32_t* __restrict__ start,
                                          const int32_t* __restrict__ end, const int32_t* __restrict__ abr, const int32_t* __restrict__ n_runs,
                                          const uint16_t* __restrict__ runs, const uint8_t* __restrict__ rc_out, int32_t* as_prev, int32_t* ae_prev,
                                          int32_t* abr_prev, int32_t* nruns_prev, uint16_t* runs_prev, uint8_t* flip) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= m) return;
  const int i = src[w];
  const bool r = rc_out[i] != 0;
  as_prev[w] = start[i]; ae_prev[w] = end[i];
  abr_prev[w] = r ? 0 : abr[i];
  nruns_prev[w] = n_runs[i];
  for (int k = 0; k < MAX_RUNS; k++) runs_prev[w * MAX_RUNS + k] = runs[(int64_t)i * MAX_RUNS + k];
  flip[w] = r && !revcomp[w];
}

// -D: the three whole-reference attempts of a strand-unknown read (mia_main.c:120-174) as reads of a scratch batch:
// item 3q = the stored read (forward matrix), 3q + 1 = the stored read (strand-reversed matrix: whatever a->submat was left
// pointing at, H6), 3q + 2 = its reverse complement (strand-reversed matrix)
__global__ void fs_retry_reads_kernel(int m, const int32_t* __restrict__ list, const uint8_t* __restrict__ bases, const int64_t* __restrict__ off,
                                      const int64_t* __restrict__ off_new, uint8_t* out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= 3 * m) return;
  const int i = list[w / 3], var = w % 3;
  const int64_t o = off[i], L = off[i + 1] - o, d = off_new[w];
  for (int64_t j = lane; j < L; j += 32) {
    uint8_t b = var == 2 ? bases[o + L - 1 - j] : bases[o + j];
    if (var == 2) {                                  // revcom_char, map_align.c:418-431
      const char* from = "ABCDGHKMNRSTUVWXY";
      const char* to = "TVGHCDMKNYSAABWXR";
      uint8_t r = 'N';
      for (int q = 0; q < 17; q++) if (from[q] == b) r = to[q];
      b = b == '-' ? '-' : r;
    }
    out[d + j] = b;
  }
}
// what the host's resolution of the attempts decided: upd[6 * q ..] = read, rc, as, ae, score, reverse-complement the stored read
__global__ void fs_apply_kernel(int m, const int32_t* __restrict__ upd, uint8_t* known, uint8_t* rc, int32_t* as, int32_t* ae, int32_t* score,
                                uint8_t* bases, const int64_t* __restrict__ off) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= m) return;
  const int32_t* u = upd + 6 * w;
  const int i = u[0];
  if (lane == 0) { known[i] = 1; rc[i] = (uint8_t)u[1]; as[i] = u[2]; ae[i] = u[3]; score[i] = u[4]; }
  if (!u[5]) return;
  const int64_t o = off[i], L = off[i + 1] - o;
  auto comp = [](uint8_t b) -> uint8_t {
    const char* from = "ABCDGHKMNRSTUVWXY";
    const char* to = "TVGHCDMKNYSAABWXR";
    uint8_t r = 'N';
    for (int q = 0; q < 17; q++) if (from[q] == b) r = to[q];
    return b == '-' ? '-' : r;
  };
  for (int64_t j = lane; j < (L + 1) / 2; j += 32) {
    const uint8_t a = bases[o + j], z = bases[o + L - 1 - j];
    bases[o + j] = comp(z);
    bases[o + L - 1 - j] = comp(a);
  }
}

}  // namespace miagpu
