// sweep16.cuh -- pass 1 WITHOUT the k-mer filter: both whole strands of a read in packed 16-bit SIMD (sm_100a).
//
// sg_align (mia.c:1500-1610) runs dyn_prog over the read against the whole wrapped reference, forward strand and
// reverse-complement strand, both with the forward matrix (H5), and keeps the better strand.  The two matrices have the
// same rows (same read, same substitution profile) and the same width, so they ride in the two halves of one
// register: low half = forward strand, high half = reverse-complement strand -- the lane frame, the 25-entry
// per-row table, the FMA-pipe adds and the folded pure-diagonal verdict are those of pair16.cuh (read 1. - 4.
// there), with K = 16 columns per lane and 16 lanes per read: a warp carries two reads of equal length.
//
// The strands are 16,825 columns (1 Mb + 256 for the nuclear case), not a window: the sweep goes CHUNK by chunk of
// 256 columns, all L rows per chunk, and hands the next chunk, per row, exactly what a lane's left neighbour hands
// over inside a chunk -- the last two cells of the row, the verdict carry of the last cell, and the running
// column-gap maximum (best_gap_col as a value: mia.c:838-850) -- through a per-read ring in shared memory.  In the
// lane frame a hand-over is the constant shift -GEP*K whether the neighbour is a lane or the previous chunk.
//
// Per strand the kernel keeps the first maximum of the last row over all chunks (max_sg_score, mia.c:1278-1302)
// and the verdict at that cell; it writes them as the two "jobs" of the read, and p1_merge_kernel (pass1.cuh)
// picks the strand, applies sg_align's coordinates, or hands the read to the general kernel (strip.cuh) when the
// winner's path is not one plain diagonal.  Reads longer than the 16-bit frame holds never come here.
#pragma once
#include "common.cuh"
#include "pair16.cuh"

namespace miagpu {

constexpr int SW_K = 16;                            // columns per lane
constexpr int SW_G = 16;                            // lanes per read
constexpr int SW_CW = SW_K * SW_G;                  // chunk width
constexpr int SW_CLASS = 7;                         // the pair class whose frame limits apply (K = 16)

struct Sweep16Params {
  const uint8_t* bases;
  const int64_t* off;
  const int32_t* pairs;          // [n_items][2] reads of equal length; -1 = empty second slot
  const int32_t* n_items;
  int32_t* counter;
  const uint8_t* ref_fw;         // codes of the wrapped forward strand
  const uint8_t* ref_rc;         // codes of the wrapped reverse-complement strand
  int32_t len1;
  const int16_t* prof16;
  uint32_t gep2;
  // per job (2 * read + strand), the layout p1_merge_kernel reads
  int32_t* jscore; int32_t* jabc; int32_t* jaec; int32_t* jabr;
  uint8_t* jstatus;
};

// dynamic shared memory: [prof16][rowoff WARPS*2*P16_MAXL u16][tab WARPS*2*2*P16_TAB_WORDS u32][ring WARPS*2*2*P16_MAXL*4 u32]
__host__ __device__ constexpr int sw_smem() {
  return (PROF16_N + 8) * 2 + WARPS_PER_BLOCK * 2 * P16_MAXL * 2 + WARPS_PER_BLOCK * 2 * 2 * P16_TAB_WORDS * 4 +
         WARPS_PER_BLOCK * 2 * 2 * P16_MAXL * 16;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4) sweep16_kernel(Sweep16Params p) {
  constexpr int K = SW_K, G = SW_G;
  constexpr int NE = (25 + G - 1) / G;
  extern __shared__ __align__(16) uint8_t smem[];
  int16_t* s_prof = reinterpret_cast<int16_t*>(smem);
  constexpr int PROF_BYTES = (PROF16_N + 8) * 2;
  constexpr int ROWOFF_BYTES = WARPS_PER_BLOCK * 2 * P16_MAXL * 2;
  constexpr int TAB_BYTES = WARPS_PER_BLOCK * 2 * 2 * P16_TAB_WORDS * 4;
  uint16_t* s_rowoff = reinterpret_cast<uint16_t*>(smem + PROF_BYTES);
  uint32_t* s_tab = reinterpret_cast<uint32_t*>(smem + PROF_BYTES + ROWOFF_BYTES);
  uint4* s_ring = reinterpret_cast<uint4*>(smem + PROF_BYTES + ROWOFF_BYTES + TAB_BYTES);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & (G - 1), hw = lane / G;
  const unsigned gmask = 0xffffu << (hw * 16);
  for (int i = tid; i < PROF16_N + 8; i += blockDim.x) s_prof[i] = p.prof16[i];
  __syncthreads();

  uint16_t* row = s_rowoff + (warp * 2 + hw) * P16_MAXL;
  uint32_t* tab = s_tab + (warp * 2 + hw) * 2 * P16_TAB_WORDS;
  uint4* ring = s_ring + (size_t)(warp * 2 + hw) * 2 * P16_MAXL;      // two buffers of P16_MAXL rows: {l2, l1, acc, scan total}
  const uint32_t prof_base = smem_u32(s_prof);
  const uint32_t tab_addr = smem_u32(tab);

  constexpr int OFF = p16_off(K);
  constexpr int CONV = GEP * K;
  constexpr int SENT = -32768 + GOP;
  const int n_items = *p.n_items;
  const uint32_t gep2 = p.gep2;
  const int len1 = p.len1;

  uint32_t eoa[NE], eob[NE];
#pragma unroll
  for (int t = 0; t < NE; t++) {
    const int e = min(sub + G * t, 24);
    eoa[t] = (e / 5) * 2;
    eob[t] = (e % 5) * 2;
  }
  // both halves use the read's own row: entry (a, b) = { sub(r, a), sub(r, b) << 16 }
  auto build_table = [&](int r, uint32_t* dst) {
    const uint32_t pr = prof_base + row[r];
#pragma unroll
    for (int t = 0; t < NE; t++) {
      const uint2 v = make_uint2((uint32_t)lds_s16(pr + eoa[t]), lds_u16(pr + eob[t]) << 16);
      if (sub + G * t < 25) reinterpret_cast<uint2*>(dst)[sub + G * t] = v;
    }
  };

  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(p.counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    int rd = p.pairs[2 * item + hw];
    const bool live = rd >= 0;                       // an empty slot recomputes the item's first read and writes nothing
    if (!live) rd = p.pairs[2 * item];
    const int64_t o = p.off[rd];
    const int L = (int)(p.off[rd + 1] - o);          // the same for both reads of the item
    __syncwarp();
    for (int r = sub; r < L; r += G) row[r] = (uint16_t)(prof_row_index(0, sm_depth(r, L), base_code(p.bases[o + r])) * 2);
    __syncwarp();

    int bestv[2] = {INT_MIN, INT_MIN}, bestc[2] = {0, 0};
    bool bestbad[2] = {false, false};
    int cur = 0;                                     // ring buffer the current chunk WRITES; it reads the other one
    for (int c0 = 0; c0 < len1; c0 += SW_CW, cur ^= 1) {
      const bool first = c0 == 0;
      uint4* rin = ring + (cur ^ 1) * P16_MAXL;
      uint4* rout = ring + cur * P16_MAXL;
      // lane masks of the read's first lane: in the first chunk it has no left neighbour (one LOP3 instead of a select)
      const bool edge = sub == 0 && first;
      const uint32_t keep = edge ? 0u : 0xffffffffu;
      const uint32_t sentm = edge ? B2(SENT) : 0u;
      const uint32_t ncmp0m = edge ? B2(-(GOP + 3 * GEP) - OFF) : 0u;
      uint32_t comb[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int c = c0 + sub * K + j;
        int a = 4, b = 4;
        if (c < len1) { a = __ldg(p.ref_fw + c); b = __ldg(p.ref_rc + c); }
        comb[j] = tab_addr + (uint32_t)(a * 5 + b) * 8;
      }
      __syncwarp();
      build_table(0, tab);
      __syncwarp();
      // ---- row 0 (mia.c:769-785)
      uint32_t W[K], Rg[K], acc[K];
      if (L > 1) build_table(1, tab + P16_TAB_WORDS);
#pragma unroll
      for (int j = 0; j < K; j++) {
        const uint2 e = lds_entry<0>(comb[j]);
        W[j] = B2(GEP * j - OFF) + e.x + e.y;
        Rg[j] = B2(-32768);
        acc[j] = 0;
      }
      if (sub == G - 1) rout[0] = make_uint4(W[K - 2], W[K - 1], 0u, 0u);
      __syncwarp();

      auto dp_row = [&](int r, auto par, const uint32_t (&W)[K], const uint32_t (&acc)[K], uint32_t (&Wn)[K], uint32_t (&accn)[K]) {
        constexpr int PAR = decltype(par)::value;
        if (r + 1 < L) build_table(r + 1, tab + (PAR ^ 1) * P16_TAB_WORDS);
        uint32_t l2 = __shfl_up_sync(0xffffffffu, W[K - 2], 1, G);
        uint32_t l1 = __shfl_up_sync(0xffffffffu, W[K - 1], 1, G);
        uint32_t ain = __shfl_up_sync(0xffffffffu, acc[K - 1], 1, G);
        uint32_t qin = 0;                                            // biased -infinity
        if (sub == 0 && !first) {                                    // the previous chunk's last lane is this lane's left neighbour
          const uint4 a = rin[r - 1], b = rin[r];
          l2 = a.x; l1 = a.y; ain = a.z; qin = b.w;
        }
        l2 = and_or(__vadd2(l2, K2(-CONV)), keep, sentm);
        const uint32_t l1c = __vadd2(l1, K2(-CONV));
        l1 = and_or(l1c, keep, sentm);
        ain &= keep;
        // lane total of the column-gap candidates E = {l2, l1, W[0..K-3]}
        uint32_t X = __vimax3_u16x2(l2, l1, W[0]);
#pragma unroll
        for (int j = 1; j + 1 < K - 2; j += 2) X = __vimax3_u16x2(X, W[j], W[j + 1]);
        if ((K - 3) & 1) X = __vmaxu2(X, W[K - 3]);
        X = __vadd2(X, K2(-GOP));
        // what the previous chunk's scan carried up to its last lane, one lane hop away (0 = -infinity stays 0)
        const uint32_t qc = sub == 0 ? __vadd2(__vmaxu2(qin, B2(-32768 + CONV)), K2(-CONV)) : 0u;
        X = __vmaxu2(X, qc);
#pragma unroll
        for (int d = 1; d < G; d <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, X, d, G);
          X = __viaddmax_u16x2(__vmaxu2(y, B2(-32768 + CONV * d)), K2(-CONV * d), X);
        }
        uint32_t q = __shfl_up_sync(0xffffffffu, X, 1, G);
        q = __vadd2(__vmaxu2(q, B2(-32768 + CONV)), K2(-CONV));
        if (sub == 0) q = qc;
        uint32_t Q[K];
        Q[0] = __viaddmax_u16x2(l2, K2(-GOP), q);
        Q[1] = __viaddmax_u16x2(l1, K2(-GOP), Q[0]);
#pragma unroll
        for (int j = 2; j < K; j++) Q[j] = __viaddmax_u16x2(W[j - 2], K2(-GOP), Q[j - 1]);
#pragma unroll
        for (int j = 0; j < K; j++) {
          const uint32_t ncmpj = B2(-(GOP + 3 * GEP) - OFF) + K2(GEP * j);
          uint32_t D, ad;
          if (j > 0) { D = W[j - 1]; ad = acc[j - 1]; }
          else { D = and_or(l1c, keep, ncmp0m); ad = ain; }          // matrix column 0: S = sub + N, never start-new (mia.c:805-822)
          const uint32_t best = __vimax3_u16x2(D, Q[j], Rg[j]);
          Rg[j] = __viaddmax_u16x2(D, K2(-GOP), Rg[j]);
          const uint2 e = lds_entry<PAR * P16_TAB_WORDS * 4>(comb[j]);
          uint32_t bp;
          Wn[j] = cell_pair(best, ncmpj, e.x, e.y, gep2, bp);
          accn[j] = ad | (bp ^ D);
        }
        if (sub == G - 1) {                                          // hand-over to the next chunk: this row's cells, this row's scan total
          rout[r] = make_uint4(Wn[K - 2], Wn[K - 1], accn[K - 1], X);
        }
        __syncwarp();
      };
      {
        uint32_t W1[K], acc1[K];
        int r = 1;
        for (; r + 1 < L; r += 2) {
          dp_row(r, std::integral_constant<int, 1>{}, W, acc, W1, acc1);
          dp_row(r + 1, std::integral_constant<int, 0>{}, W1, acc1, W, acc);
        }
        if (r < L) {
          dp_row(r, std::integral_constant<int, 1>{}, W, acc, W1, acc1);
#pragma unroll
          for (int j = 0; j < K; j++) { W[j] = W1[j]; acc[j] = acc1[j]; }
        }
      }
      // ---- this chunk's part of max_sg_score: first maximum of the last row, strict '>' across chunks
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int best = INT_MIN;
#pragma unroll
        for (int j = 0; j < K; j++) {
          const int cl = sub * K + j;
          const int v = (int)(h ? (W[j] >> 16) : (W[j] & 0xffffu)) - 32768 - GEP * j;
          const int key = (c0 + cl < len1) ? v * 512 + (KEY_IDX_MASK - cl) : INT_MIN;
          best = max(best, key);
        }
        best = __reduce_max_sync(gmask, best);
        const int cl = KEY_IDX_MASK - (best & KEY_IDX_MASK);
        const int v = best >> 9;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < K; j++)
          if (sub * K + j == cl) bad = (h ? (acc[j] >> 16) : (acc[j] & 0xffffu)) != 0;
        bad = __any_sync(gmask, bad);
        if (best != INT_MIN && v > bestv[h]) { bestv[h] = v; bestc[h] = c0 + cl; bestbad[h] = bad; }
      }
      __syncwarp();
    }
    if (sub == 0 && live) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int64_t j = 2 * (int64_t)rd + h;
        const int aec = bestc[h];
        const int nsteps = min(L - 1, aec);
        p.jscore[j] = bestv[h] + OFF - GEP * (L - 1);
        p.jaec[j] = aec;
        p.jabc[j] = aec - nsteps;
        p.jabr[j] = L - 1 - nsteps;
        p.jstatus[j] = bestbad[h] ? P16_ST_GENERAL : MIAGPU_ST_OK;
      }
    }
  }
}

// every read becomes two "jobs" (forward strand, reverse-complement strand) of the sweep; reads the 16-bit frame does
// not hold go to the general kernel.  new_kmer_filter returns 1 without a filter (kmer.c:251-255).
struct SweepPrepParams {
  int64_t n;
  const int64_t* off;
  int lmax;
  uint8_t* kind; uint8_t* route; int32_t* jfirst; uint16_t* jcount; uint8_t* jkind; int32_t* hits;
  int32_t* general_list; int32_t* meta;
};
__global__ void sweep_prep_kernel(SweepPrepParams p, int p1_ngeneral) {
  const int64_t rd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rd >= p.n) return;
  const int L = (int)(p.off[rd + 1] - p.off[rd]);
  p.hits[rd] = 1;
  if (L >= 1 && L <= p.lmax) {
    p.kind[rd] = (uint8_t)(16 + SW_CLASS);
    p.route[rd] = 1;
    p.jfirst[rd] = (int32_t)(2 * rd);
    p.jcount[rd] = 0x0101;
    p.jkind[2 * rd] = p.jkind[2 * rd + 1] = (uint8_t)(16 + SW_CLASS);
    atomicAdd(&p.meta[META_HIST + SW_CLASS * (P16_MAXL + 1) + L], 1);
    atomicAdd(&p.meta[META_PREADS + SW_CLASS], 1);
  } else {
    p.kind[rd] = 0;
    p.route[rd] = 2;
    p.jcount[rd] = 0;
    p.general_list[atomicAdd(p.meta + p1_ngeneral, 1)] = (int32_t)rd;
  }
}

}  // namespace miagpu
