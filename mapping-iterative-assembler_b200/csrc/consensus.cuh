// consensus.cuh -- per-column accumulation and base calling (sm_100a).
//
// Replaces consensus_assembly_string (mia.c:515-603) and what feeds it:
// merge_pwaln_into_maln's seq/ins/gaps (map_align.c:866-954), pop_smp_from_FSDB
// (fsdb.c:542-619), cull_maln_from_fsdb's gaps rescan (mia.c:486-504), add_base
// (map_align.c:229-263), find_ins_cons (444-510) and find_consensus (294-391).
//
// The reference scans ALL AlnSeqs for EVERY reference position: O(ref_len x N).
// Here one warp walks one AlnSeq entry (O(sum read_len) total): lanes take
// consecutive reference columns, locate them in the run list, derive the base,
// the PSSM depth code (smp) and the four column scores, and add into plane-major
// int32 accumulators with coalesced RED.ADDs (lane i -> address base+i).
// Integer sums commute, so the result is bit-identical for any schedule and any
// number of GPUs (the planes are all-reduced between accumulate and call).
#pragma once
#include "common.cuh"

namespace miagpu {

constexpr int NPLANE = MIAGPU_COUNTS_PER_COL;   // As,Cs,Gs,Ts,gaps,cov,sA,sC,sG,sT
constexpr int PL_GAPS = 4, PL_COV = 5, PL_SCORE = 6;

struct ConsParams {
  const miagpu_entry* entries;
  int64_t n_entries;
  const uint8_t* bases;
  const int64_t* off;
  const uint8_t* rc;
  const int32_t* abr;
  const int32_t* n_runs;
  const uint16_t* runs;
  const int32_t* sm;          // [2][775] forward then strand-reversed, reference layout sm[d][ref][read]
  int32_t seq_len;
  int32_t* gaps;              // [seq_len]
  const int32_t* ins_off;     // [seq_len+1] exclusive scan of gaps (gaps[0] forced to 0)
  int32_t* acc;               // [NPLANE][n_cols], n_cols = seq_len + ins_off[seq_len]
  int64_t n_cols;
};

__device__ __forceinline__ int smp_depth(const miagpu_entry& e, int act) {   // fsdb.c:569-580 / 598-609
  const int dfront = e.back_formula ? e.front_len + act : act;
  const int dback = e.total_len - act - 1;
  return dfront <= PSSM_DEPTH ? dfront : (dback < PSSM_DEPTH ? 2 * PSSM_DEPTH - dback : PSSM_DEPTH);
}

// add_base (map_align.c:229-263) into column `col` of the plane-major accumulators
__device__ __forceinline__ void add_base_dev(int32_t* acc, int64_t n_cols, int64_t col, int ch_code /*0..4, 5 = '-'*/,
                                             const int32_t* sm_strand, int depth) {
  atomicAdd(acc + PL_COV * n_cols + col, 1);
  if (ch_code == 5) { atomicAdd(acc + PL_GAPS * n_cols + col, 1); return; }
  if (ch_code < 4) atomicAdd(acc + ch_code * n_cols + col, 1);
  const int32_t* m = sm_strand + depth * 25 + ch_code;          // sm[depth][X][b]
#pragma unroll
  for (int x = 0; x < 4; x++) atomicAdd(acc + (PL_SCORE + x) * n_cols + col, m[x * 5]);
}

// MODE 0: ref->gaps = max insert length per position over the culled list (mia.c:486-504;
//         positions with start < pos <= end only, so an insert in front of an entry's first
//         column never counts)
// MODE 1: base columns + insert columns
template <int MODE>
__global__ void __launch_bounds__(256) entry_kernel(ConsParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= p.n_entries) return;
  const miagpu_entry e = p.entries[w];
  if (e.col_count <= 0) return;
  const int rd = e.read;
  const int nr = p.n_runs[rd];
  if (nr <= 0) return;
  const uint16_t* runs = p.runs + (int64_t)rd * MAX_RUNS;
  const uint8_t* read = p.bases + p.off[rd];
  const int32_t* sm_strand = p.sm + (p.rc[rd] ? MIAGPU_PSSM_INTS : 0);
  const int cb = e.col_begin, ce = e.col_begin + e.col_count;

  int colpos = 0;                 // reference columns before this run
  int rpos = p.abr[rd];           // read rows consumed before this run (absolute row)
  const int row0 = rpos;
  int pend = 0;                   // inserted bases waiting for the next column
  for (int k = 0; k < nr; k++) {
    const int x = runs[k];
    const int type = x >> 14, len = x & 0x3fff;
    if (type == MIAGPU_RUN_I) { pend = len; rpos += len; continue; }
    const int lo = max(colpos, cb), hi = min(colpos + len, ce);
    for (int i = lo + lane; i < hi; i += 32) {
      const int pos = e.ref_pos + (i - cb);
      if (pos >= p.seq_len) continue;                       // consensus loop stops at seq_len (mia.c:555)
      const bool isM = type == MIAGPU_RUN_M;
      const int row = rpos + (i - colpos);                  // read row of an M column
      const int q = (i == colpos) ? pend : 0;               // insert length in front of this column
      if (MODE == 0) {
        if (q > 0 && i > cb && pos > 0) atomicMax(p.gaps + pos, q);
      } else {
        // act = read bases consumed before the column, inserted ones included (fsdb.c:564-583)
        const int act = e.act_bias + (isM ? row - row0 : rpos - row0);
        const int depth = smp_depth(e, act);
        if (!e.dropped) {
          const int ch = isM ? base_code(read[row]) : 5;
          add_base_dev(p.acc, p.n_cols, pos + p.ins_off[pos + 1], ch, sm_strand, depth);   // base column sits after its insert columns
        }
        const int g = (i > cb && pos > 0) ? p.gaps[pos] : 0; // find_ins_cons: start < pos <= end, dropped NOT checked
        for (int j = 0; j < g; j++) {
          const int ch = j < q ? base_code(read[row - q + j]) : 5;
          add_base_dev(p.acc, p.n_cols, pos + p.ins_off[pos] + j, ch, sm_strand, depth);
        }
      }
    }
    colpos += len;
    if (type == MIAGPU_RUN_M) rpos += len;
    pend = 0;
  }
}

// find_consensus (map_align.c:294-391) for every column of the padded layout.
// out[col] = called character ('-' for a gap consensus, dropped on the host side).
__global__ void call_kernel(const int32_t* acc, int64_t n_cols, int cons_code, char* out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  const int cov = acc[PL_COV * n_cols + c], gaps = acc[PL_GAPS * n_cols + c];
  char r;
  if (cov == 0) r = 'N';
  else if ((double)gaps / (double)cov >= 0.5) r = '-';
  else {
    const int sA = acc[(PL_SCORE + 0) * n_cols + c], sC = acc[(PL_SCORE + 1) * n_cols + c];
    const int sG = acc[(PL_SCORE + 2) * n_cols + c], sT = acc[(PL_SCORE + 3) * n_cols + c];
    int top = sA, second = INT_MIN;
    char base = 'A';
    if (sC >= top) { second = top; top = sC; base = 'C'; } else second = sC;        // '>=': later base wins ties
    if (sG >= top) { second = top; top = sG; base = 'G'; } else if (sG >= second) second = sG;
    if (sT >= top) { second = top; top = sT; base = 'T'; } else if (sT >= second) second = sT;
    if (cons_code == 2) r = (top >= 0 || (top - 2400) > second) ? base : 'N';       // MIN_SC_DIFF_CONS
    else r = top >= -399 ? base : 'N';                                              // MIN_SCORE_CONS
  }
  out[c] = r;
}

// Device-side construction of the "natural" entry list: read i owns entries 2i (whole
// alignment or front part) and 2i+1 (wrapped back part, col_count = 0 when not split).
// Mirrors mia_main.c:259-276 (end fix, split test), split_pwaln (mia.c:1376-1438) and
// asp_len (fsdb.c:518-530); used when the host has no stale AlnSeq pointers to describe.
__global__ void natural_entries_kernel(int64_t n, const int32_t* as_out, const int32_t* ae_out, const int32_t* n_runs,
                                       const uint16_t* runs, const uint8_t* status, int seq_len, const uint8_t* dropped_front,
                                       const uint8_t* dropped_back, miagpu_entry* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  miagpu_entry f{}, b{};
  f.read = b.read = (int32_t)i;
  const int nr = n_runs[i];
  if (nr > 0 && !(status[i] & MIAGPU_ST_UNSUPPORTED)) {
    const int start = as_out[i];
    int end = ae_out[i];
    if (end > seq_len) end -= seq_len;
    const bool split = start > end;
    const int cf = split ? seq_len - start : 0x7fffffff;     // alignment columns that stay in front
    int cols = 0, ins = 0, fins = 0;
    const uint16_t* r = runs + i * MAX_RUNS;
    for (int k = 0; k < nr; k++) {
      const int x = r[k], t = x >> 14, len = x & 0x3fff;
      if (t == MIAGPU_RUN_I) { ins += len; if (cols < cf) fins += len; }
      else cols += len;
    }
    const int fcols = split ? cf : cols;
    const int fl = fcols + fins, bl = split ? (cols - fcols) + (ins - fins) : 0;
    f.col_begin = 0; f.col_count = fcols; f.ref_pos = start; f.front_len = fl; f.total_len = fl + bl;
    f.dropped = dropped_front ? dropped_front[i] : 0;
    if (split) {
      b = f;
      b.col_begin = fcols; b.col_count = cols - fcols; b.ref_pos = 0; b.back_formula = 1;
      b.dropped = dropped_back ? dropped_back[i] : f.dropped;
    }
  }
  out[2 * i] = f;
  out[2 * i + 1] = b;
}

// gaps[0] is ignored by the consensus (mia.c:557 "ref_pos > 0"); force it to 0 so the scan
// allocates no columns for it
__global__ void zero_first_kernel(int32_t* gaps) { gaps[0] = 0; }

}  // namespace miagpu
