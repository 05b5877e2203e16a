// realign.cuh -- windowed PSSM semi-global DP + on-device traceback (sm_100a).
//
// Replaces, for a whole batch of reads, the per-read body of reiterate_assembly
// (mia_main.c:178-257): pop_s2c_in_a / pop_s1c_in_a (mia.c:1243, 1054),
// dyn_prog (mia.c:740-981), max_sg_score (mia.c:1278-1302), find_align_begin
// (mia.c:612-637) and populate_pwaln_to_begin (mia.c:1440-1497).
//
// Parallelisation.  dyn_prog's cell (r,c) reads only rows < r:
//     D   = S[r-1][c-1]
//     G_c = max_{k<=c-2} S[r-1][k] - P(c-k-1)      (best_gap_col, running arg-max along row r-1)
//     G_r = max_{j<=r-2} S[j][c-1] - P(r-j-1)      (best_gap_row[c-1], running arg-max down column c-1)
// so a whole row is computed at once: ONE WARP PER READ, lane l owns K consecutive
// columns [l*K, l*K+K) for every row, all state in registers.  There is no
// anti-diagonal skew and no fill/drain loss; the only cross-lane traffic per row
// is 3 shuffles for the left neighbours and a 5-step warp max-scan that turns the
// per-lane arg-max of (S[r-1][k] + GEP*k) into the running prefix the reference
// keeps in best_gap_col.  Both arg-maxes are carried as single packed 32-bit keys
// (common.cuh) so "value then earliest index" is one IMNMX.
//
// Traceback.  Every cell stores a 16-bit trace word (the low 16 bits of the
// winning key: 2-bit move marker + 9-bit jump target) into a per-warp scratch
// matrix in global memory, written with one coalesced 128-bit (or 64-bit) store
// per lane per row.  Scratch is re-used by the persistent warp for every read it
// processes and sized to stay L2-resident.  The walk back reproduces the
// reference's quirks: a jump target of 0 is indistinguishable from "diagonal"
// (trace == 0, H2), the walk stops in row 0 / column 0 / at a start-new cell.
#pragma once
#include "common.cuh"

namespace miagpu {

struct RealignParams {
  // resident reads
  const uint8_t* bases;
  const int64_t* off;
  // per read inputs
  const uint8_t* rc;
  const int32_t* win_start;
  const int32_t* win_len;
  // work list of this bucket
  const int32_t* list;
  int32_t n_list;            // upper bound used to size the grid
  const int32_t* n_list_ptr; // device-side length of the list (the pair kernel appends to it before this launch)
  int32_t* counter;          // dynamic work fetch
  // reference codes (0..4), padded to 16 B
  const uint8_t* ref_codes;
  int32_t ref_bytes;         // padded length (multiple of 16)
  int32_t ref_in_smem;
  const int32_t* prof;       // PROF_INTS
  int32_t sg5;
  // outputs
  int32_t* score;
  int32_t* as_out;
  int32_t* ae_out;
  int32_t* abr;
  int32_t* n_runs;
  uint16_t* runs;
  uint8_t* status;
  // trace scratch
  uint32_t* scratch;
  int64_t scratch_words_per_warp;
  // TRIM (trim_frag, mia.c:1318-1368): every work item has the SAME rows (the adapter: bases[0 .. shared_rows)), item i is
  // read i, whose bases are the matrix columns (ref_codes + win_start[i], win_len[i] columns); the end cell is the first
  // maximum of the LAST COLUMN in row order; as_out / ae_out are matrix columns (abc, aec), aer_out the end row
  int32_t shared_rows;
  int32_t* aer_out;
  unsigned long long* cells_done;   // nullable: DP cells of the reads this launch computed (measurement)
  // JOB (pass 1 with the k-mer filter, pass1.cuh): a work item names a JOB -- the read job_read[job] & 0x7fffffff against one stretch
  // of unmasked columns of the strand bit 31 names (win_start / win_len indexed by job, in the two-strand code array) -- whose 16-bit
  // run found a winning path that is not one plain diagonal.  Forward matrix on both strands (H5); a stretch that does not begin at
  // the strand's column 0 starts a new alignment in its first column (pair16.cuh, JOB).  Outputs are sg_align's (mia.c:1568-1610),
  // per read.  An instantiation takes the jobs with job_wl_lo < win_len <= job_wl_hi and leaves the others to its siblings.
  const int32_t* job_read;
  int32_t strand_stride, job_wl_lo, job_wl_hi, seq_len;
  int32_t* start;
  int32_t* end;
  // JOB, per-job outputs instead (job_score != nullptr): jobs the 16-bit kernels could not finish exactly (pair16.cuh 5., "sunk")
  // are computed here before the merge looks at them -- score, abc / aec in strand coordinates, abr, and whether the path is one
  // plain diagonal (status MIAGPU_ST_OK) or not (0x40: the merge has the winner traced)
  int32_t* job_score; int32_t* job_abc; int32_t* job_aec; int32_t* job_abr;
  uint8_t* job_status;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// (a & b) | c in ONE LOP3 (b in a register so that ptxas does not split two immediates)
__device__ __forceinline__ int and_or(int a, int b, int c) {
  int d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int lds_s32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// TMA-style bulk copy global -> shared with mbarrier completion (UBLKCP in SASS)
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

template <int K>
struct TraceLayout {
  static constexpr int W = (K + 1) / 2;                            // 32-bit words of trace per lane per row
  static constexpr int WP = (W <= 2) ? 2 : (W <= 4) ? 4 : 8;       // padded to a vector store
  static constexpr int ROW_WORDS = 32 * WP;
};

constexpr int WARPS_PER_BLOCK = 4;

// dynamic shared memory: [prof PROF_INTS ints][rowoff WARPS*256 u16][ref codes]
template <int K, bool TRIM = false, bool JOB = false>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) realign_kernel(RealignParams p) {
  static_assert(!(TRIM && JOB), "one mode at a time");
  static_assert(K >= 2 && K <= 16, "columns per lane");
  using TL = TraceLayout<K>;
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t ref_bar;
  int32_t* s_prof = reinterpret_cast<int32_t*>(smem);
  uint16_t* s_rowoff = reinterpret_cast<uint16_t*>(smem + PROF_INTS * 4);
  uint8_t* s_ref = smem + PROF_INTS * 4 + WARPS_PER_BLOCK * MAX_READ * 2;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  // ---- stage the scoring profile and (if it fits) the whole reference once per block
  if (p.ref_in_smem) {
    if (tid == 0) {
      mbar_init(&ref_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&ref_bar, (uint32_t)p.ref_bytes);
      // bulk copies are limited only by the tx counter (2^20-1 bytes); chunk anyway
      for (int o = 0; o < p.ref_bytes; o += 32768) {
        int n = min(32768, p.ref_bytes - o);
        bulk_g2s(s_ref + o, p.ref_codes + o, (uint32_t)n, &ref_bar);
      }
    }
  }
  for (int i = tid; i < PROF_INTS; i += blockDim.x) s_prof[i] = p.prof[i] * KEY_MUL;   // pre-multiplied: sub lands in the key's value field
  if (p.ref_in_smem) mbar_wait(&ref_bar, 0);
  __syncthreads();

  uint16_t* my_rowoff = s_rowoff + warp * MAX_READ;
  const uint32_t prof_base = smem_u32(s_prof);
  const int64_t gwarp = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  uint32_t* trace = p.scratch + gwarp * p.scratch_words_per_warp;

  // per-lane column constants.  Scores are kept in "diagonal key" form Sd = S*2048 + (DIAG<<9).
  // candP[j] turns the Sd of column c into its best_gap_col candidate key: value S + GEP*c, marker COL, index c
  int candP[K];
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int c = lane * K + j;
    candP[j] = (GEP * c) * KEY_MUL + ((MARK_COL - MARK_DIAG) << 9) + (KEY_IDX_MASK - c);
  }
  const int candL2 = candP[0] - 2 * (GEP * KEY_MUL - 1);      // columns lane*K-2 and lane*K-1 (held by the left lane)
  const int candL1 = candP[0] - 1 * (GEP * KEY_MUL - 1);
  const int gcLane = -((GOP - GEP) + GEP * (lane * K)) * KEY_MUL;   // key(G_c) = P - (800 + 200*c)*2048
  int value_mask = ~KEY_LOW_MASK;
  asm volatile("" : "+r"(value_mask));                            // keep it in a register

  const int n_list = p.n_list_ptr ? *p.n_list_ptr : p.n_list;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(p.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_list) break;
    const int job = JOB ? p.list[item] : 0;
    const int jr = JOB ? p.job_read[job] : 0;
    const int rd = JOB ? (jr & 0x7fffffff) : TRIM ? item : p.list[item];
    const int ws = p.win_start[JOB ? job : rd];
    const int len1 = p.win_len[JOB ? job : rd];
    if (JOB && !(len1 > p.job_wl_lo && len1 <= p.job_wl_hi)) continue;       // a sibling instantiation's job (warp-uniform)
    const int64_t o0 = TRIM ? 0 : p.off[rd];
    const int L = TRIM ? p.shared_rows : (int)(p.off[rd + 1] - o0);
    const int strand = (TRIM || JOB) ? 0 : (p.rc[rd] ? 1 : 0);
    const int jlo = JOB ? ws - (jr < 0 ? p.strand_stride : 0) : 0;            // the stretch's first column in strand coordinates
    const bool masked_left = JOB && jlo > 0;
    // TRIM: the lane / register that hold the last column, its running first maximum over the rows
    const int lcl = TRIM ? (len1 - 1) / K : 0, lcj = TRIM ? (len1 - 1) - lcl * K : 0;
    int lc_best = INT_MIN, lc_row = 0;

    // ---- read -> per-row profile offsets (pop_s2c_in_a + find_sm_depth), lanes in parallel
    __syncwarp();
    for (int r = lane; r < L; r += 32) {
      int code = base_code(p.bases[o0 + r]);
      my_rowoff[r] = (uint16_t)(prof_row_index(strand, sm_depth(r, L), code) * 4);
    }
    // ---- reference window -> per-column code*4 (pop_s1c_in_a)
    int code4[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
      int c = lane * K + j;
      int code = 4;
      if (c < len1) code = p.ref_in_smem ? s_ref[ws + c] : p.ref_codes[ws + c];
      code4[j] = code * 4;
    }
    __syncwarp();

    // ---- row 0 (mia.c:769-785): plain substitution scores, no trace
    int Sp[K], R[K];
    {
      uint32_t pa = prof_base + my_rowoff[0];
#pragma unroll
      for (int j = 0; j < K; j++) {
        Sp[j] = lds_s32(pa + code4[j]) + DIAG_BITS;          // profile is pre-multiplied by 2048
        R[j] = NEG_KEY;
      }
    }
    if (TRIM && lane == lcl) {
#pragma unroll
      for (int j = 0; j < K; j++) if (j == lcj) lc_best = Sp[j];
    }

    for (int r = 1; r < L; r++) {
      const uint32_t pa = prof_base + my_rowoff[r];
      const int N = p.sg5 ? -(GOP + GEP * (r + 1)) : 0;                       // mia.c:877-880
      const int keyN = N * KEY_MUL;                                           // START: marker 0, loses every tie
      const int NdKey = keyN + DIAG_BITS;                                     // the cell's value if it starts here
      const int grK = -((GOP - GEP) + GEP * r) * KEY_MUL;                     // key(G_r) = R - (800 + 200*r)*2048
      const int rowK = (GEP * (r - 1)) * KEY_MUL + ((MARK_ROW - MARK_DIAG) << 9) + (KEY_IDX_MASK - (r - 1));

      // candidate keys of row r-1: column k becomes a best_gap_col candidate at column k+2
      int cand[K];
      int l2 = __shfl_up_sync(0xffffffffu, Sp[K - 2], 1);
      int l1 = __shfl_up_sync(0xffffffffu, Sp[K - 1], 1);
      cand[0] = lane ? l2 + candL2 : NEG_KEY;
      if (K > 1) cand[1] = lane ? l1 + candL1 : NEG_KEY;
#pragma unroll
      for (int j = 2; j < K; j++) cand[j] = Sp[j - 2] + candP[j - 2];
      int D = lane ? l1 : NdKey;         // column 0: S = sub + N, trace 0 (mia.c:805-822)
      if (JOB && masked_left && lane == 0) D = NEG_KEY;     // masked left neighbour: nothing to continue, the cell starts new (S = N)

      // lane total, then exclusive warp max-scan = best_gap_col state entering this lane
      int t = cand[0];
#pragma unroll
      for (int j = 1; j < K; j++) t = max(t, cand[j]);
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t = max(t, v);
      }
      int P = __shfl_up_sync(0xffffffffu, t, 1);
      if (lane == 0) P = NEG_KEY;

      uint32_t tw[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        P = max(P, cand[j]);
        const int subK = lds_s32(pa + code4[j]);
        const int kGc = P + (gcLane - GEP * KEY_MUL * j);
        const int kGr = R[j] + grK;
        const int best3 = max(max(D, kGc), kGr);        // DIAG > COL > ROW on ties (mia.c:922-948)
        const bool pS = keyN > best3;                   // strictly better than all three (mia.c:910-915)
        tw[j] = (uint32_t)(pS ? 0 : best3);             // low 16 bits: marker + jump target
        const int cont = and_or(best3 + subK, value_mask, DIAG_BITS);
        R[j] = max(R[j], D + rowK);                     // row r-1 joins best_gap_row[c-1] for row r+1
        D = Sp[j];
        Sp[j] = pS ? NdKey : cont;                      // start-new does NOT add the substitution score
      }
      if (lane == 0) R[0] = NEG_KEY;                    // there is no column -1
      if (TRIM && lane == lcl) {                        // mia.c:1345-1352: strict '>' in row order
        int v = INT_MIN;
#pragma unroll
        for (int j = 0; j < K; j++) if (j == lcj) v = Sp[j];
        if (v > lc_best) { lc_best = v; lc_row = r; }
      }

      // ---- one vector store of this lane's K trace words (row r stored at r-1)
      uint32_t* trow = trace + (int64_t)(r - 1) * TL::ROW_WORDS + lane * TL::WP;
      uint32_t wv[TL::WP];
#pragma unroll
      for (int w = 0; w < TL::WP; w++) wv[w] = (w < TL::W) ? __byte_perm(tw[2 * w], (2 * w + 1 < K) ? tw[(2 * w + 1 < K) ? 2 * w + 1 : 0] : 0u, 0x5410) : 0u;
      if (TL::WP == 2) {
        *reinterpret_cast<uint2*>(trow) = make_uint2(wv[0], wv[1]);
      } else {
#pragma unroll
        for (int w = 0; w < TL::WP; w += 4) *reinterpret_cast<uint4*>(trow + w) = make_uint4(wv[w], wv[w + 1], wv[w + 2], wv[w + 3]);
      }
    }

    // ---- max_sg_score (mia.c:1278-1302): first maximum of the last row
    int best = INT_MIN;
#pragma unroll
    for (int j = 0; j < K; j++) {
      int c = lane * K + j;
      int key = (c < len1) ? (Sp[j] - DIAG_BITS) + (KEY_IDX_MASK - c) : INT_MIN;
      best = max(best, key);
    }
    best = __reduce_max_sync(0xffffffffu, best);
    int score = best >> KEY_SHIFT;
    int aec = KEY_IDX_MASK - (best & KEY_IDX_MASK);
    int aer = L - 1;
    if (TRIM) {
      score = (__shfl_sync(0xffffffffu, lc_best, lcl) - DIAG_BITS) >> KEY_SHIFT;
      aer = __shfl_sync(0xffffffffu, lc_row, lcl);
      aec = len1 - 1;
    }

    // ---- find_align_begin + populate_pwaln_to_begin, executed uniformly by the warp
    __syncwarp();   // trace stores of all lanes visible (same warp, global memory)
    int row = aer, col = aec, nrun = 0, curM = 0, ncols = 0;
    uint16_t* my_runs = (p.runs && !(JOB && p.job_score)) ? p.runs + (int64_t)rd * MAX_RUNS : nullptr;
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(trace);
    auto push = [&](int type, int len) {
      if (my_runs && nrun < MAX_RUNS && lane == 0) my_runs[nrun] = (uint16_t)((type << 14) | len);
      nrun++;
      ncols += len;
    };
    // Lanes fetch the next 32 cells of the DIAGONAL through (row,col) in one round trip; the walk then
    // consumes the leading run of plain diagonal moves at once and handles the first other cell.
    for (;;) {
      const int tr_ = row - lane, tc_ = col - lane;
      uint32_t w16 = MARK_START << 9;                        // off-matrix: never consumed
      if (tr_ > 0 && tc_ > 0) {
        const int tl = tc_ / K, tj = tc_ - tl * K;
        w16 = __ldcg(t16 + ((int64_t)(tr_ - 1) * TL::ROW_WORDS + tl * TL::WP) * 2 + tj);
      }
      const int mk_l = (w16 >> 9) & 3;
      const int idx_l = KEY_IDX_MASK - (int)(w16 & KEY_IDX_MASK);
      const bool diag_l = tr_ > 0 && tc_ > 0 && (mk_l == MARK_DIAG || (mk_l != MARK_START && idx_l == 0));   // idx 0: trace == 0 reads as diagonal (H2)
      const unsigned nd = __ballot_sync(0xffffffffu, !diag_l);
      const int n = nd ? __ffs(nd) - 1 : 32;                 // leading diagonal moves
      curM += n; row -= n; col -= n;
      if (n == 32) continue;
      if (row <= 0 || col <= 0) break;                       // row 0 / column 0: trace 0 == -row or == col (mia.c:617-618)
      const int mk = __shfl_sync(0xffffffffu, mk_l, n), idx = __shfl_sync(0xffffffffu, idx_l, n);
      if (mk == MARK_START) break;
      curM++;
      if (mk == MARK_ROW) {                                  // mia.c:1466-1476
        push(MIAGPU_RUN_M, curM); curM = 0;
        push(MIAGPU_RUN_I, row - 1 - idx);
        row = idx; col--;
      } else {                                               // mia.c:1477-1487
        push(MIAGPU_RUN_M, curM); curM = 0;
        push(MIAGPU_RUN_D, col - 1 - idx);
        col = idx; row--;
      }
    }
    push(MIAGPU_RUN_M, curM + 1);
    if (lane == 0) {
      uint8_t st = MIAGPU_ST_OK;
      if (nrun > MAX_RUNS) { st |= MIAGPU_ST_RUNS_OVERFLOW; nrun = -1; }
      if (ncols > 2 * MAX_READ) st |= MIAGPU_ST_STR_OVERFLOW;
      if (my_runs)
        for (int a = 0, b = nrun - 1; a < b; a++, b--) {     // runs were produced 3'->5'
          uint16_t x = my_runs[a]; my_runs[a] = my_runs[b]; my_runs[b] = x;
        }
      if (p.cells_done) atomicAdd(p.cells_done, (unsigned long long)L * (unsigned long long)len1);
      if (JOB && p.job_score) {
        p.job_score[job] = score;
        p.job_abc[job] = col + jlo; p.job_aec[job] = aec + jlo; p.job_abr[job] = row;
        p.job_status[job] = (nrun == 1 && st == MIAGPU_ST_OK) ? MIAGPU_ST_OK : 0x40;
      } else if (JOB) {
        // sg_align's coordinates (mia.c:1568-1610); runs go out in forward-reference orientation (strip.cuh does the same)
        const int s = jr < 0 ? 1 : 0;
        const int abc = col + jlo, aes = aec + jlo;
        int start = abc, end = aes;
        if (s == 1) {                                        // c2rcc, mia.c:26-30; revcom_PWAF reverses the columns
          start = p.seq_len - (aes % p.seq_len) - 1;
          end = p.seq_len - (abc % p.seq_len) - 1;
          if (my_runs && nrun > 0)
            for (int a = 0, b = nrun - 1; a < b; a++, b--) { uint16_t x = my_runs[a]; my_runs[a] = my_runs[b]; my_runs[b] = x; }   // undo the reversal above
        }
        int as = start, ae = end;
        if (as > ae) ae = p.seq_len + as;                    // mia.c:1600-1604
        if (end > p.seq_len) end -= p.seq_len;               // mia.c:1606-1610
        if (score != p.score[rd]) st |= MIAGPU_ST_UNSUPPORTED;   // the 16-bit job's score is exact: anything else is a bug, not a result
        p.as_out[rd] = as; p.ae_out[rd] = ae; p.start[rd] = start; p.end[rd] = end;
        p.abr[rd] = s == 1 ? 0 : row;
        p.n_runs[rd] = nrun;
        p.status[rd] = st;
      } else {
        p.score[rd] = score;
        p.as_out[rd] = TRIM ? col : col + ws;       // mia_main.c:250-255
        p.ae_out[rd] = TRIM ? aec : aec + ws;
        p.abr[rd] = row;
        if (TRIM) p.aer_out[rd] = aer;
        if (p.n_runs) p.n_runs[rd] = nrun;
        p.status[rd] = st;
      }
    }
  }
}

}  // namespace miagpu
