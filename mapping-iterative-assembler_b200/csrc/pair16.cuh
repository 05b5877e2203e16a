// pair16.cuh -- windowed PSSM semi-global DP in packed 16-bit SIMD (s16x2), TWO READS PER WARP,
// with an in-band diagonal traceback (sm_100a).
//
// Same contract as realign.cuh (reiterate_assembly's per-read body, mia_main.c:178-257:
// dyn_prog mia.c:740-981, max_sg_score 1278-1302, find_align_begin 612-637,
// populate_pwaln_to_begin 1440-1497), for the common case: short reads whose path is a plain
// diagonal near the expected one.  Everything else is handed, read by read, to the 32-bit
// kernel of realign.cuh through its work list -- never approximated.
//
// 1. Two reads of equal length share a group of G lanes: every 32-bit register holds the same DP
//    cell of read A (low half) and read B (high half), so one VIADDMNMX.S16x2 / VIMNMX3.S16x2 /
//    VIADD.16x2 does two cells.  Lane l of the group owns columns [l*K, l*K+K) as in realign.cuh.
//    G = 16: a warp carries two such pairs (four reads of equal length), which halves the
//    per-row cost of the cross-lane scan, the neighbour shuffles and the loop itself per cell.
//
// 2. Row frame.  Scores are kept as  V(r,c) = S(r,c) + GEP*r - OFF.  In this frame
//      * the start-new candidate N_r = -(GOP + GEP*(r+1)) (mia.c:877-880) is the CONSTANT
//        NCMP = -(GOP+2*GEP) - OFF when compared in row r-1's frame, and a start-new cell is the
//        constant NCMP + GEP;
//      * the row-gap candidate max_j S[j][c-1] - P(r-j-1) (mia.c:856-868) is  max_j V(j,c-1) - GOP:
//        a plain running maximum, no per-row decay;
//      * the column-gap candidate keeps its GEP decay per column (mia.c:838-850):
//        Q(c) = max(Q(c-1) - GEP, V(r-1,c-2) - (GOP+GEP)), evaluated as a per-lane chain T plus a
//        5-step cross-lane max-scan with decay GEP*K per lane.
//    A gap candidate below NCMP can never be chosen nor change the start-new test, so the scan
//    clamps from below before subtracting its decay: no 16-bit operation wraps (checked lane
//    for lane against the oracle by tests/model/pair16_model.c).  OFF and the longest read the
//    frame can hold follow from the matrices' extreme entries (pair16_limits in miagpu.cu).
//
// 3. Only scores are computed: no arg-max indices, no trace words.  The start-new rule "S = N,
//    substitution score NOT added" is a predicated move of the profile address (the cell reads
//    the constant GEP instead of its substitution score), not a select on the result.
//
// 4. Traceback.  Cells within +-P16_BAND diagonals of the expected one (window start = as - 50)
//    are stored (16 bit per cell, a few lanes per row) in a per-warp scratch that stays in L2.
//    From the first maximum of the last row the warp checks, 32 rows at a time, that every cell
//    up to row 0 / column 0 satisfies  V(r,c) - sub(r,c) == V(r-1,c-1)  and  V(r-1,c-1) >= N_r:
//    exactly the condition under which dyn_prog stores trace 0 there.  If it holds the alignment
//    is one M run; if not (gap, start-new cell, jump stored as 0, path outside the band) the read
//    is appended to the 32-bit kernel's list.
#pragma once
#include "common.cuh"
#include "realign.cuh"

namespace miagpu {

constexpr int P16_BAND = 16;
constexpr int P16_DIAG0 = REALIGN_BUFFER;
constexpr int P16_MAXL = 144;                       // rows the shared row-offset arrays hold
constexpr int PROF16_N = 2 * NMAT * 5 * PROF_ROW_INTS;   // int16 entries (sub + GEP); entry PROF16_N holds GEP
constexpr int P16_NKB = 4;                          // width classes K = 4, 5, 6, 8 (128 / 160 / 192 / 256 columns)

__host__ __device__ inline int p16_class(int len1) { return len1 <= 128 ? 0 : len1 <= 160 ? 1 : len1 <= 192 ? 2 : len1 <= 256 ? 3 : -1; }
__host__ __device__ inline int bucket32_of(int len1) {
  return len1 <= 64 ? 0 : len1 <= 128 ? 1 : len1 <= 160 ? 2 : len1 <= 192 ? 3 : len1 <= 224 ? 4 : len1 <= 256 ? 5 : len1 <= 320 ? 6 : len1 <= 384 ? 7 : len1 <= 512 ? 8 : 9;
}

struct Pair16Params {
  const uint8_t* bases;
  const int64_t* off;
  const uint8_t* rc;
  const int32_t* win_start;
  const int32_t* win_len;
  const int32_t* pairs;          // [n_items][32/G][2] read ids; -1 = empty slot (a work item's first slot is never empty)
  const int32_t* n_items;        // device counter (layout kernel wrote it)
  int32_t* counter;
  const uint8_t* ref_codes;
  int32_t ref_bytes;
  int32_t ref_in_smem;
  const int16_t* prof16;
  int32_t off16;
  int32_t* score;
  int32_t* as_out;
  int32_t* ae_out;
  int32_t* abr;
  int32_t* n_runs;
  uint16_t* runs;
  uint8_t* status;
  int32_t* lists;                // 32-bit kernel work lists [NBUCKET][n]
  int32_t* list_counts;          // their fill counters
  int64_t n_reads;
  int32_t* n_fallback;           // statistics
  uint32_t* scratch;
  int64_t scratch_words_per_warp;
};

// Band scratch of one warp: plane 0 = W0/4 uint4s [row][lane] (the lane's first W0 columns), plane 1 = KR
// words [row][lane] (the remaining K % 4 columns).  Only the lanes whose columns intersect the band are ever written, so the
// footprint that lives in L2 is ~(2*BAND+K)*4 B per row although the address range is the full matrix.
template <int K>
struct BandLayout {
  static constexpr int REM = K % 4;
  static constexpr int KR = REM == 0 ? 0 : REM == 1 ? 1 : REM == 2 ? 2 : 4;   // words of the second plane per lane
  static constexpr int W0 = REM == 0 ? K : K - REM;                           // words of the first plane (uint4s)
  static constexpr int WORDS_PER_ROW = 32 * (W0 + KR);
};

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_s16(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// max(best, ncmp) per half; where a half of `best` is below ncmp (start-new) the matching profile
// address is replaced by the address of the constant GEP.  The setp.eq pattern is the one ptxas
// folds into VIMNMX.S16x2 with two predicate outputs; the moves stay predicated (FMA-pipe IMAD.MOV).
__device__ __forceinline__ uint32_t vmax_start(uint32_t best, uint32_t ncmp, uint32_t& addr_lo, uint32_t& addr_hi, uint32_t addr_gep) {
  uint32_t r;
  asm("{.reg .pred pu, pv;\n\t"
      ".reg .s16 h0, h1, h2, h3;\n\t"
      "max.s16x2 %0, %3, %4;\n\t"
      "mov.b32 {h0, h1}, %0;\n\t"
      "mov.b32 {h2, h3}, %3;\n\t"
      "setp.eq.s16 pv, h0, h2;\n\t"
      "setp.eq.s16 pu, h1, h3;\n\t"
      "@!pv mov.b32 %1, %5;\n\t"
      "@!pu mov.b32 %2, %5;}\n\t"
      : "=r"(r), "+r"(addr_lo), "+r"(addr_hi)
      : "r"(best), "r"(ncmp), "r"(addr_gep));
  return r;
}
__device__ __forceinline__ uint32_t mad16(uint32_t hi, uint32_t lo) {   // hi * 65536 + lo on the FMA pipe
  uint32_t r;
  asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(r) : "r"(hi), "r"(lo));
  return r;
}

// dynamic shared memory: [prof16 (PROF16_N + 8) int16][rowoff WARPS*(32/G)*2*P16_MAXL u16][ref codes]
template <int K, int G>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) pair16_kernel(Pair16Params p) {
  static_assert(K >= 4 && K <= 16 && (G == 16 || G == 32), "columns per lane / lanes per pair");
  using BL = BandLayout<K>;
  constexpr int NP = 32 / G;                         // pairs per warp
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(8) uint64_t ref_bar;
  int16_t* s_prof = reinterpret_cast<int16_t*>(smem);
  constexpr int PROF_BYTES = (PROF16_N + 8) * 2;
  uint16_t* s_rowoff = reinterpret_cast<uint16_t*>(smem + PROF_BYTES);
  uint8_t* s_ref = smem + PROF_BYTES + WARPS_PER_BLOCK * NP * 2 * P16_MAXL * 2;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int sub = lane & (G - 1);                    // lane within the pair's group
  const int hw = lane / G;                           // which pair of the warp
  const unsigned gmask = G == 32 ? 0xffffffffu : (0xffffu << (hw * 16));

  if (p.ref_in_smem) {
    if (tid == 0) {
      mbar_init(&ref_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&ref_bar, (uint32_t)p.ref_bytes);
      for (int o = 0; o < p.ref_bytes; o += 32768) bulk_g2s(s_ref + o, p.ref_codes + o, (uint32_t)min(32768, p.ref_bytes - o), &ref_bar);
    }
  }
  for (int i = tid; i < PROF16_N + 8; i += blockDim.x) s_prof[i] = p.prof16[i];
  if (p.ref_in_smem) mbar_wait(&ref_bar, 0);
  __syncthreads();

  uint16_t* rowA = s_rowoff + ((warp * NP + hw) * 2 + 0) * P16_MAXL;
  uint16_t* rowB = s_rowoff + ((warp * NP + hw) * 2 + 1) * P16_MAXL;
  const uint32_t prof_base = smem_u32(s_prof);
  const uint32_t addr_gep = prof_base + PROF16_N * 2;
  const int64_t gwarp = (int64_t)blockIdx.x * WARPS_PER_BLOCK + warp;
  uint32_t* band = p.scratch + gwarp * p.scratch_words_per_warp;

  const int OFF = p.off16;
  const int NCMP = -(GOP + 2 * GEP) - OFF;
  uint32_t NCMP2 = pack2(NCMP);
  const uint32_t SEED2 = pack2(-OFF - GEP);
  constexpr uint32_t M_OPEN = ((uint32_t)(-(GOP + GEP)) & 0xffffu) * 0x10001u;     // -(GOP+GEP) in both halves
  constexpr uint32_t M_GEP = ((uint32_t)(-GEP) & 0xffffu) * 0x10001u;
  constexpr uint32_t M_GOP = ((uint32_t)(-GOP) & 0xffffu) * 0x10001u;
  constexpr uint32_t SENT2 = ((uint32_t)(-32768 + GEP * K + GEP) & 0xffffu) * 0x10001u;
  constexpr uint32_t CLK2 = ((uint32_t)(-32768 + GEP * K) & 0xffffu) * 0x10001u;
  constexpr uint32_t NEG2 = 0x80008000u;
  const int n_items = *p.n_items;

  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(p.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    int rdA = p.pairs[2 * (item * NP + hw)];
    int rdB = p.pairs[2 * (item * NP + hw) + 1];
    const bool hasA = rdA >= 0;                       // an empty pair slot recomputes the warp's first pair and writes nothing
    if (!hasA) rdA = p.pairs[2 * (item * NP)];
    const bool hasB = hasA && rdB >= 0;
    if (rdB < 0 || !hasA) rdB = rdA;
    const int64_t oA = p.off[rdA], oB = p.off[rdB];
    const int L = (int)(p.off[rdA + 1] - oA);         // the same for every read of the work item
    const int wsA = p.win_start[rdA], wsB = p.win_start[rdB];
    const int lenA = p.win_len[rdA], lenB = p.win_len[rdB];
    const int sA = p.rc[rdA] ? 1 : 0, sB = p.rc[rdB] ? 1 : 0;

    __syncwarp();
    for (int r = sub; r < L; r += G) {
      const int d = sm_depth(r, L);
      rowA[r] = (uint16_t)(prof_row_index(sA, d, base_code(p.bases[oA + r])) * 2);
      rowB[r] = (uint16_t)(prof_row_index(sB, d, base_code(p.bases[oB + r])) * 2);
    }
    uint32_t cA[K], cB[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
      const int c = sub * K + j;
      int a = 4, b = 4;
      if (c < lenA) a = p.ref_in_smem ? s_ref[wsA + c] : p.ref_codes[wsA + c];
      if (c < lenB) b = p.ref_in_smem ? s_ref[wsB + c] : p.ref_codes[wsB + c];
      cA[j] = a * 2;
      cB[j] = b * 2;
    }
    __syncwarp();

    // band bookkeeping: row r keeps the lanes that own a column c with c - r in [DIAG0 - BAND, DIAG0 + BAND]:
    // sub*K + K-1 >= r + DIAG0 - BAND  and  sub*K <= r + DIAG0 + BAND
    uint4* plane0 = reinterpret_cast<uint4*>(band);
    uint32_t* plane1 = band + (size_t)L * 32 * BL::W0;
    auto store_row = [&](int r, const uint32_t* W) {
      const int u = sub * K + (K - 1) - (P16_DIAG0 - P16_BAND) - r;
      if ((unsigned)u <= (unsigned)(2 * P16_BAND + K - 1)) {
        const int e = r * 32 + lane;
#pragma unroll
        for (int q = 0; q < BL::W0 / 4; q++) plane0[e * (BL::W0 / 4) + q] = make_uint4(W[4 * q], W[4 * q + 1], W[4 * q + 2], W[4 * q + 3]);
        constexpr int B1 = BL::W0;                    // first column of the second plane
        if (BL::KR == 1) plane1[e] = W[B1 < K ? B1 : 0];
        if (BL::KR == 2) *reinterpret_cast<uint2*>(plane1 + 2 * e) = make_uint2(W[B1 < K ? B1 : 0], W[B1 + 1 < K ? B1 + 1 : 0]);
        if (BL::KR == 4) *reinterpret_cast<uint4*>(plane1 + 4 * e) = make_uint4(W[B1 < K ? B1 : 0], W[B1 + 1 < K ? B1 + 1 : 0], W[B1 + 2 < K ? B1 + 2 : 0], W[B1 + 3 < K ? B1 + 3 : 0]);
      }
    };

    // ---- row 0 (mia.c:769-785): V = sub - OFF
    uint32_t W[K], Rg[K];
    {
      const uint32_t pa = prof_base + rowA[0], pb = prof_base + rowB[0];
#pragma unroll
      for (int j = 0; j < K; j++) {
        W[j] = __vadd2(mad16(lds_u16(pb + cB[j]), lds_u16(pa + cA[j])), SEED2);
        Rg[j] = NEG2;
      }
    }
    store_row(0, W);

    for (int r = 1; r < L; r++) {
      const uint32_t pa = prof_base + rowA[r], pb = prof_base + rowB[r];
      const uint32_t l2 = __shfl_up_sync(0xffffffffu, W[K - 2], 1, G);
      const uint32_t l1 = __shfl_up_sync(0xffffffffu, W[K - 1], 1, G);
      // per-lane chain of column-gap candidates (incoming prefix taken as -inf)
      uint32_t T[K];
      T[0] = sub ? __vadd2(l2, M_OPEN) : SENT2;
      T[1] = __viaddmax_s16x2(T[0], M_GEP, sub ? __vadd2(l1, M_OPEN) : SENT2);
#pragma unroll
      for (int j = 2; j < K; j++) T[j] = __viaddmax_s16x2(T[j - 1], M_GEP, __vadd2(W[j - 2], M_OPEN));
      // inclusive cross-lane scan of the lane totals, decay GEP*K per lane, clamped so nothing wraps
      uint32_t X = T[K - 1];
#pragma unroll
      for (int d = 1; d < G; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, X, d, G);
        const uint32_t cl = ((uint32_t)(-32768 + GEP * K * d) & 0xffffu) * 0x10001u;
        const uint32_t dec = ((uint32_t)(-GEP * K * d) & 0xffffu) * 0x10001u;
        X = __viaddmax_s16x2(__vmaxs2(y, cl), dec, X);    // lanes < d get their own X back from the shuffle: max(X, max(X,cl)-dec) = X
      }
      uint32_t qin = __shfl_up_sync(0xffffffffu, X, 1, G);
      if (sub == 0) qin = SENT2;
      qin = __vmaxs2(qin, CLK2);

      uint32_t D = sub ? l1 : NCMP2;            // column 0: S = sub + N, never start-new (mia.c:805-822)
#pragma unroll
      for (int j = 0; j < K; j++) {
        const uint32_t mj = ((uint32_t)(-GEP * (j + 1)) & 0xffffu) * 0x10001u;
        const uint32_t Q = __viaddmax_s16x2(qin, mj, T[j]);
        const uint32_t best = __vimax3_s16x2(D, Q, Rg[j]);
        Rg[j] = __viaddmax_s16x2(D, M_GOP, Rg[j]);          // row r-1 joins the row-gap candidates of column c-1
        uint32_t aA = pa + cA[j], aB = pb + cB[j];
        const uint32_t bp = vmax_start(best, NCMP2, aA, aB, addr_gep);
        D = W[j];
        W[j] = __vadd2(bp, mad16(lds_u16(aB), lds_u16(aA)));
      }
      store_row(r, W);
    }

    // ---- max_sg_score + in-band diagonal traceback, one read (half) at a time, each group for its own pair
    __syncwarp();
#pragma unroll 1
    for (int h = 0; h < 2; h++) {
      const bool live = h ? hasB : hasA;              // uniform within the group
      const int rd = h ? rdB : rdA;
      const int ws = h ? wsB : wsA;
      const int len1 = h ? lenB : lenA;
      const uint16_t* rowX = h ? rowB : rowA;
      int best = INT_MIN;
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int c = sub * K + j;
        const int v = h ? ((int)W[j] >> 16) : (int)(short)(W[j] & 0xffffu);
        const int key = (c < len1) ? v * 512 + (KEY_IDX_MASK - c) : INT_MIN;
        best = max(best, key);
      }
      best = __reduce_max_sync(gmask, best);
      const int aec = KEY_IDX_MASK - (best & KEY_IDX_MASK);
      const int score = (best >> 9) + OFF - GEP * (L - 1);
      const int nsteps = min(L - 1, aec);
      const int dg = aec - (L - 1);
      bool ok = dg >= P16_DIAG0 - P16_BAND && dg <= P16_DIAG0 + P16_BAND;
      auto cell = [&](int r, int c) -> int {
        const int lc = c / K, j = c - lc * K;
        const int e = r * 32 + hw * G + lc;
        const uint32_t w = j < BL::W0 ? __ldcg(band + BL::W0 * e + j) : __ldcg(plane1 + BL::KR * e + (j - BL::W0));
        return h ? ((int)w >> 16) : (int)(short)(w & 0xffffu);
      };
      for (int t0 = 0; ok && t0 < nsteps; t0 += G) {
        const int t = t0 + sub;
        bool good = true;
        if (t < nsteps) {
          const int r = L - 1 - t, c = aec - t;
          const int v = cell(r, c), dv = cell(r - 1, c - 1);
          const int code = p.ref_in_smem ? s_ref[ws + c] : p.ref_codes[ws + c];
          const int sb = lds_s16(prof_base + rowX[r] + code * 2);       // sub + GEP; V(r) = V(r-1) + sub + GEP on a diagonal move
          good = (v - sb == dv) && (dv >= NCMP);
        }
        ok = __all_sync(gmask, good);
      }
      if (sub == 0 && live) {
        if (ok) {
          p.score[rd] = score;
          p.as_out[rd] = aec - nsteps + ws;
          p.ae_out[rd] = aec + ws;
          p.abr[rd] = L - 1 - nsteps;
          p.n_runs[rd] = 1;
          p.runs[(int64_t)rd * MAX_RUNS] = (uint16_t)((MIAGPU_RUN_M << 14) | (nsteps + 1));
          p.status[rd] = MIAGPU_ST_OK;
        } else {
          const int b = bucket32_of(len1);
          const int slot = atomicAdd(p.list_counts + b, 1);
          p.lists[(int64_t)b * p.n_reads + slot] = rd;
          atomicAdd(p.n_fallback, 1);
        }
      }
    }
  }
}

}  // namespace miagpu
