// scorecut.cuh -- device side of an iteration's score cut (a12): find_fsdb_score_cut (fsdb.c:269-383) and the
// per-read test of cull_maln_from_fsdb (mia.c:418-479) over scores that are already in HBM.
//
// The regression's two double-precision chains (scorecut.hpp explains why they can be taken block-wise without
// changing a bit) are split like this:
//   cut_stats_kernel   integer sums (sum len, sum score, count) and the best score per read length   [per DP chunk]
//   -- host: xbar, ybar, the 257-entry tables (len - xbar), (len - xbar)^2 -- the same doubles the reference forms
//   cut_approx_kernel  plain block sums of both chains (predict the binade of the running sum at every block start)
//   cut_exact_kernel   per block: T = sum rint(a_i / ulp), A = sum |rint(a_i / ulp)|, tie / range flags
//   -- host: chain_stitch (in order; unproven blocks read by read), slope / intercept, threshold per length
//   cut_flags_kernel   below = score < threshold[len]; sticky |= below (H10); entry flags for the accumulation
// Every addend is formed with __dmul_rn / __dsub_rn / __dadd_rn: IEEE operations that nvcc never contracts into an
// FMA, so a_i, a_i / ulp and the magic-constant rounding are bit for bit what the host code computes.
#pragma once
#include <limits.h>

#include "common.cuh"

namespace miagpu {

constexpr int CUT_BLOCK = 512;                       // = CHAIN_BLOCK of scorecut.hpp
constexpr int CUT_THREADS = 256;
constexpr int CUT_PER_THREAD = CUT_BLOCK / CUT_THREADS;

struct CutStatsDev {
  long long sx, sy, cnt, bad;                        // bad = smallest index whose seq_len is outside [0, MAX_READ] (LLONG_MAX = none)
  int best[MAX_READ + 1];
  int pad;                                           // reads newly dropped by cut_flags_kernel (reset by cut_init_kernel)
  long long sxx, sxy;                                // sum len^2, sum len * score over the reads the fit uses (sharded rounds: where a rank's chain starts)
};
struct CutTables {                                   // host-made, the same doubles the reference forms per read
  double ybar;
  double dx[MAX_READ + 1];
  double dx2[MAX_READ + 1];
};
struct CutBlockDev {                                 // chain 0 = ssxy, chain 1 = ssxx
  double approx[2], T[2], A[2];
  int e[2], ok[2];
  double run[2];                                     // approximate running sums at the block start (cut_prefix_kernel)
};

__global__ void cut_init_kernel(CutStatsDev* st) {
  const int t = threadIdx.x;
  if (t == 0) { st->sx = 0; st->sy = 0; st->cnt = 0; st->bad = LLONG_MAX; st->pad = 0; st->sxx = 0; st->sxy = 0; }
  for (int l = t; l <= MAX_READ; l += blockDim.x) st->best[l] = INT_MIN;
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  return v;
}

// reads [lo, hi): used = unique_best && score >= FIRST_ROUND_SCORE_CUTOFF (fsdb.c:283-293)
__global__ void __launch_bounds__(CUT_THREADS) cut_stats_kernel(int64_t lo, int64_t hi, const int32_t* __restrict__ seq_len,
                                                                const int32_t* __restrict__ score, const uint8_t* __restrict__ unique_best,
                                                                CutStatsDev* st) {
  __shared__ int s_best[MAX_READ + 1];
  __shared__ long long s_sum[5];
  for (int l = threadIdx.x; l <= MAX_READ; l += blockDim.x) s_best[l] = INT_MIN;
  if (threadIdx.x < 5) s_sum[threadIdx.x] = 0;
  __syncthreads();
  long long sx = 0, sy = 0, cnt = 0, sxx = 0, sxy = 0;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = seq_len[i];
    if (l < 0 || l > MAX_READ) { atomicMin(&st->bad, (long long)i); continue; }
    const int sc = score[i];
    if ((!unique_best || unique_best[i]) && sc >= FIRST_ROUND_SCORE_CUTOFF) {
      sx += l; sy += sc; cnt++; sxx += (long long)l * l; sxy += (long long)l * sc;
      if (sc > s_best[l]) atomicMax(&s_best[l], sc);
    }
  }
  sx = warp_sum_ll(sx); sy = warp_sum_ll(sy); cnt = warp_sum_ll(cnt); sxx = warp_sum_ll(sxx); sxy = warp_sum_ll(sxy);
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd((unsigned long long*)&s_sum[0], (unsigned long long)sx);
    atomicAdd((unsigned long long*)&s_sum[1], (unsigned long long)sy);
    atomicAdd((unsigned long long*)&s_sum[2], (unsigned long long)cnt);
    atomicAdd((unsigned long long*)&s_sum[3], (unsigned long long)sxx);
    atomicAdd((unsigned long long*)&s_sum[4], (unsigned long long)sxy);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_sum[2]) {
    atomicAdd((unsigned long long*)&st->sx, (unsigned long long)s_sum[0]);
    atomicAdd((unsigned long long*)&st->sy, (unsigned long long)s_sum[1]);
    atomicAdd((unsigned long long*)&st->cnt, (unsigned long long)s_sum[2]);
    atomicAdd((unsigned long long*)&st->sxx, (unsigned long long)s_sum[3]);
    atomicAdd((unsigned long long*)&st->sxy, (unsigned long long)s_sum[4]);
  }
  for (int l = threadIdx.x; l <= MAX_READ; l += blockDim.x)
    if (s_best[l] != INT_MIN) atomicMax(&st->best[l], s_best[l]);
}

// Where the chains' per-read inputs come from: the resident arrays of one GPU, or (sharded rounds)
// the all-gathered keys of every rank: key = score << 9 | seq_len for a read the fit uses, CUT_KEY_UNUSED otherwise.
// The gathered space is world * stride words, stride a multiple of CUT_BLOCK; words at local index >= n_max
// (padding and the rank's header) count as unused reads: a 0.0 addend leaves a rounded chain unchanged.
constexpr uint32_t CUT_KEY_UNUSED = 0xffffffffu;
constexpr int SHARD_HDR_WORDS = 8;                   // tail of a rank's stride: sum len, sum score, count (int64 each), 2 spare
constexpr int SHARD_PF_SLOTS = 96;                   // chain blocks (per rank) whose keys travel with the block records
constexpr int SHARD_PF_FIRST = 48;                   // ... of which the first ones come to the host with the records (the rest only when a rank used them)

struct CutSrc {
  const int32_t* seq_len; const int32_t* score; const uint8_t* unique_best;       // KEYS = false
  const uint32_t* keys; int64_t n_max, stride;                                    // KEYS = true
};

__device__ __forceinline__ uint32_t cut_key(int l, int sc, bool unique) {
  return (unique && sc >= FIRST_ROUND_SCORE_CUTOFF) ? ((uint32_t)sc << 9) | (uint32_t)min(max(l, 0), MAX_READ) : CUT_KEY_UNUSED;
}

// the two addends of read i: (len - xbar) * (score - ybar) and (len - xbar)^2, 0 for a read the fit does not use
template <bool KEYS>
__device__ __forceinline__ void cut_addends(int64_t i, int64_t local, const CutSrc& s, const CutTables* __restrict__ t, double& axy, double& axx) {
  axy = 0.0; axx = 0.0;
  if (KEYS) {
    if (local >= s.n_max) return;
    const uint32_t key = s.keys[i];
    if (key == CUT_KEY_UNUSED) return;
    const int l = key & 511, sc = (int)(key >> 9);
    axy = __dmul_rn(t->dx[l], __dsub_rn((double)sc, t->ybar));
    axx = t->dx2[l];
  } else {
    const int sc = s.score[i];
    if ((!s.unique_best || s.unique_best[i]) && sc >= FIRST_ROUND_SCORE_CUTOFF) {
      const int l = min(max(s.seq_len[i], 0), MAX_READ);
      axy = __dmul_rn(t->dx[l], __dsub_rn((double)sc, t->ybar));
      axx = t->dx2[l];
    }
  }
}

// block-wide sum of two doubles, result valid in every thread
__device__ __forceinline__ void block_sum2(double& a, double& b, double* s_red) {
  a = warp_sum_d(a); b = warp_sum_d(b);
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { s_red[2 * w] = a; s_red[2 * w + 1] = b; }
  __syncthreads();
  a = 0; b = 0;
#pragma unroll
  for (int k = 0; k < CUT_THREADS / 32; k++) { a += s_red[2 * k]; b += s_red[2 * k + 1]; }
}

template <bool KEYS>
__global__ void __launch_bounds__(CUT_THREADS) cut_approx_kernel(int64_t n, CutSrc src, const CutTables* __restrict__ t, CutBlockDev* blk) {
  __shared__ double s_red[2 * CUT_THREADS / 32];
  const int64_t i0 = (int64_t)blockIdx.x * CUT_BLOCK;
  const int64_t l0 = KEYS ? i0 % src.stride : 0;                    // a block never straddles two ranks' strides
  double s0 = 0, s1 = 0;
#pragma unroll
  for (int k = 0; k < CUT_PER_THREAD; k++) {
    const int64_t i = i0 + k * CUT_THREADS + threadIdx.x;
    if (i < n) {
      double axy, axx;
      cut_addends<KEYS>(i, l0 + k * CUT_THREADS + threadIdx.x, src, t, axy, axx);
      s0 += axy; s1 += axx;
    }
  }
  block_sum2(s0, s1, s_red);
  if (threadIdx.x == 0) { blk[blockIdx.x].approx[0] = s0; blk[blockIdx.x].approx[1] = s1; }
}

// exclusive prefix of the blocks' plain sums: one block, a contiguous run of chain blocks per thread
constexpr int CUT_PREFIX_THREADS = 1024;
__global__ void __launch_bounds__(CUT_PREFIX_THREADS) cut_prefix_kernel(int64_t nb, CutBlockDev* blk, const double* __restrict__ start = nullptr) {
  __shared__ double s_tot[2][CUT_PREFIX_THREADS];
  const int t = threadIdx.x;
  const int64_t per = (nb + CUT_PREFIX_THREADS - 1) / CUT_PREFIX_THREADS;
  const int64_t lo = min(nb, per * t), hi = min(nb, lo + per);
  double a0 = 0, a1 = 0;
  for (int64_t j = lo; j < hi; j++) { a0 += blk[j].approx[0]; a1 += blk[j].approx[1]; }
  s_tot[0][t] = a0; s_tot[1][t] = a1;
  __syncthreads();
  for (int d = 1; d < CUT_PREFIX_THREADS; d <<= 1) {       // Hillis-Steele inclusive scan
    const double v0 = t >= d ? s_tot[0][t - d] : 0.0, v1 = t >= d ? s_tot[1][t - d] : 0.0;
    __syncthreads();
    s_tot[0][t] += v0; s_tot[1][t] += v1;
    __syncthreads();
  }
  double r0 = s_tot[0][t] - a0, r1 = s_tot[1][t] - a1;
  if (start) { r0 += start[0]; r1 += start[1]; }        // sharded rounds: where this rank's part of the chains begins (predicted)
  for (int64_t j = lo; j < hi; j++) {
    blk[j].run[0] = r0; blk[j].run[1] = r1;
    r0 += blk[j].approx[0]; r1 += blk[j].approx[1];
  }
}

// KEYS: blocks the stitch will probably have to add read by read (no proof, or the predicted running sum too close
// to the edge of its binade for the block's increments) also copy their keys to one of SHARD_PF_SLOTS slots
// (pf_ids[0] = number of such blocks, pf_ids[1 + slot] = block), so that the host has them without another round trip.
template <bool KEYS>
__global__ void __launch_bounds__(CUT_THREADS) cut_exact_kernel(int64_t n, CutSrc src, const CutTables* __restrict__ t, CutBlockDev* blk,
                                                                uint32_t* pf_keys, int32_t* pf_ids) {
  __shared__ double s_red[2 * CUT_THREADS / 32];
  __shared__ int s_bad[2];
  __shared__ int s_slot;
  const int b = blockIdx.x;
  // approximate running sums at the block start (any order: it only predicts the binade, chain_stitch verifies it)
  const double run0 = blk[b].run[0], run1 = blk[b].run[1];
  if (threadIdx.x < 2) s_bad[threadIdx.x] = 0;
  __syncthreads();
  const double run[2] = {run0, run1};
  bool valid[2];
  double inv[2];
  int e[2];
#pragma unroll
  for (int ch = 0; ch < 2; ch++) {
    const double s = run[ch];
    valid[ch] = s > 0 && isfinite(s);
    int e2 = 0;
    const double f = valid[ch] ? frexp(s, &e2) : 0.75;
    if (f < 0.5 + 1e-6 || f > 1 - 1e-6) valid[ch] = false;          // too close to a binade boundary to predict
    e[ch] = e2 - 1;
    if (e[ch] < -900 || e[ch] > 900) valid[ch] = false;
    inv[ch] = valid[ch] ? ldexp(1.0, 52 - e[ch]) : 1.0;             // 1 / ulp
  }
  constexpr double MAGIC = 6755399441055744.0;                      // 1.5 * 2^52
  constexpr double LIM = 1125899906842624.0;                        // 2^50
  const int64_t i0 = (int64_t)b * CUT_BLOCK;
  const int64_t l0 = KEYS ? i0 % src.stride : 0;
  double T[2] = {0, 0}, A[2] = {0, 0};
  bool bad[2] = {false, false};
#pragma unroll
  for (int k = 0; k < CUT_PER_THREAD; k++) {
    const int64_t i = i0 + k * CUT_THREADS + threadIdx.x;
    if (i < n) {
      double a[2];
      cut_addends<KEYS>(i, l0 + k * CUT_THREADS + threadIdx.x, src, t, a[0], a[1]);
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        const double x = __dmul_rn(a[ch], inv[ch]);
        const double m = __dsub_rn(__dadd_rn(x, MAGIC), MAGIC);
        bad[ch] |= !(fabs(x) < LIM) | (fabs(__dsub_rn(x, m)) == 0.5);       // exact ties round by the parity of the running sum
        T[ch] += m; A[ch] += fabs(m);
      }
    }
  }
  if (bad[0]) s_bad[0] = 1;                                         // benign race: every writer stores 1 (after block_sum2's barriers)
  if (bad[1]) s_bad[1] = 1;
  block_sum2(T[0], A[0], s_red);
  block_sum2(T[1], A[1], s_red);
  __syncthreads();
  if (threadIdx.x < 2) {
    const int ch = threadIdx.x;
    const double Tc = ch ? T[1] : T[0], Ac = ch ? A[1] : A[0];
    blk[b].T[ch] = Tc;
    blk[b].A[ch] = Ac;
    blk[b].e[ch] = ch ? e[1] : e[0];
    blk[b].ok[ch] = (ch ? valid[1] : valid[0]) && !s_bad[ch] && Ac < 4503599627370496.0;   // 2^52: all partial integer sums exact
  }
  if (pf_keys) {
    if (threadIdx.x == 0) {
      bool suspect = false;
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        const double N0 = run[ch] * inv[ch];                          // ~ the running sum in ulps, [2^52, 2^53) when the prediction holds
        const bool ok = valid[ch] && !s_bad[ch] && A[ch] < 4503599627370496.0;
        suspect |= !ok || !(N0 * (1 - 4e-6) - A[ch] >= 4503599627370497.0) || !(N0 * (1 + 4e-6) + A[ch] <= 9007199254740990.0);
      }
      int slot = -1;
      if (suspect) { slot = atomicAdd(&pf_ids[0], 1); if (slot >= SHARD_PF_SLOTS) slot = -1; else pf_ids[1 + slot] = b; }
      s_slot = slot;
    }
    __syncthreads();
    const int slot = s_slot;
    if (slot >= 0)
      for (int k = threadIdx.x; k < CUT_BLOCK; k += CUT_THREADS) {
        const int64_t i = i0 + k;
        uint32_t key = CUT_KEY_UNUSED;
        if (KEYS) { if (i < n && l0 + k < src.n_max) key = src.keys[i]; }
        else if (i < n) key = cut_key(src.seq_len[i], src.score[i], !src.unique_best || src.unique_best[i]);
        pf_keys[(size_t)slot * CUT_BLOCK + k] = key;
      }
  }
}

// ---- sharded rounds, protocol of round 2: every rank evaluates ITS OWN part of the two chains.
//   all-reduce MAX of [insert maxima | best score per length | world x SHARD_HDR2 header words]: a rank writes its integer sums
//       into its own header row (30-bit pieces, everything else zero), so the MAX is a gather
//   shard_prep_kernel: the sums of all ranks -> xbar, ybar and the tables (the same doubles the host forms), and where this
//       rank's part of either chain starts: sum over the ranks before it of  sum(len - xbar)^2  and  sum(len - xbar)(score - ybar)
//       expanded in the ranks' integer sums -- a PREDICTION of the running sums (it only has to name the binade; the stitch checks)
//   cut_approx / cut_prefix / cut_exact over the rank's own reads -> block records + the keys of the blocks the stitch may have to
//       add read by read -> all-gather (about 50 B per 512 reads instead of 4 B per read)
//   host: chain_stitch_blocks over all ranks' records in rank order
constexpr int SHARD_HDR2 = 16;                       // header words per rank: 5 sums x 2 pieces of 30 bits, local read count
__device__ __forceinline__ void put60(int32_t* dst, long long v) { dst[0] = (int32_t)(v & 0x3fffffff); dst[1] = (int32_t)((v >> 30) & 0x3fffffff); }
__device__ __forceinline__ long long get60(const int32_t* src) { return (long long)src[0] | ((long long)src[1] << 30); }
__global__ void shard_hdr2_kernel(const CutStatsDev* st, int32_t* max_best, int32_t* hdr, int world, int rank, int n_local,
                                  const int32_t* n_slots_local = nullptr) {
  const int t = threadIdx.x;
  for (int i = t; i < world * SHARD_HDR2; i += blockDim.x) hdr[i] = 0;
  for (int l = t; l <= MAX_READ; l += blockDim.x) max_best[l] = max(st->best[l], 0);     // (no read of that length: INT_MIN -> 0; scores that count are >= 2000)
  __syncthreads();
  if (t == 0) {
    int32_t* h = hdr + rank * SHARD_HDR2;
    put60(h, st->sx); put60(h + 2, st->sy); put60(h + 4, st->cnt); put60(h + 6, st->sxx); put60(h + 8, st->sxy);
    h[10] = n_local;
    h[11] = n_slots_local ? *n_slots_local : 0;          // pointer state: AlnSeq slots this rank's reads take this round (slots.cuh)
  }
}
struct ShardPrep { double start[2]; long long sx, sy, cnt; };
__global__ void shard_prep_kernel(int world, int rank, const int32_t* __restrict__ hdr, const int32_t* __restrict__ max_best, CutStatsDev* st,
                                  CutTables* tab, ShardPrep* out, int32_t* pf_ids) {
  __shared__ double s_xbar, s_ybar;
  const int t = threadIdx.x;
  if (t == 0) {
    long long sx = 0, sy = 0, cnt = 0;
    for (int r = 0; r < world; r++) { const int32_t* h = hdr + r * SHARD_HDR2; sx += get60(h); sy += get60(h + 2); cnt += get60(h + 4); }
    st->sx = sx; st->sy = sy; st->cnt = cnt;
    out->sx = sx; out->sy = sy; out->cnt = cnt;
    const double xbar = cnt > 0 ? __ddiv_rn((double)sx, (double)cnt) : 0.0, ybar = cnt > 0 ? __ddiv_rn((double)sy, (double)cnt) : 0.0;
    s_xbar = xbar; s_ybar = ybar;
    tab->ybar = ybar;
    double a = 0, b = 0;                             // predicted running sums where this rank's reads begin
    for (int r = 0; r < rank; r++) {
      const int32_t* h = hdr + r * SHARD_HDR2;
      const double x = (double)get60(h), y = (double)get60(h + 2), c = (double)get60(h + 4), xx = (double)get60(h + 6), xy = (double)get60(h + 8);
      a += xy - xbar * y - ybar * x + c * xbar * ybar;
      b += xx - 2.0 * xbar * x + c * xbar * xbar;
    }
    out->start[0] = a; out->start[1] = b;
    pf_ids[0] = 0;
  }
  __syncthreads();
  for (int l = t; l <= MAX_READ; l += blockDim.x) {
    const double dx = __dsub_rn((double)l, s_xbar);
    tab->dx[l] = dx; tab->dx2[l] = __dmul_rn(dx, dx);
    st->best[l] = max_best[l] > 0 ? max_best[l] : INT_MIN;
  }
}
// what travels: per block T, A (2 chains), e, ok
struct ShardBlockRec { double T[2], A[2]; int e[2], ok[2]; };
__global__ void shard_records_kernel(int64_t nb, const CutBlockDev* __restrict__ blk, ShardBlockRec* out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  ShardBlockRec r;
  for (int ch = 0; ch < 2; ch++) { r.T[ch] = blk[b].T[ch]; r.A[ch] = blk[b].A[ch]; r.e[ch] = blk[b].e[ch]; r.ok[ch] = blk[b].ok[ch]; }
  out[b] = r;
}

// ---- sharded rounds (round-1 protocol, kept for reference): what a rank contributes to the all-gather / all-reduce(MAX), and the merge of what came back
__global__ void shard_pack_kernel(int64_t n, int64_t stride, const int32_t* __restrict__ seq_len, const int32_t* __restrict__ score,
                                  const uint8_t* __restrict__ unique_best, uint32_t* send) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= stride - SHARD_HDR_WORDS) return;
  send[i] = i < n ? cut_key(seq_len[i], score[i], !unique_best || unique_best[i]) : CUT_KEY_UNUSED;
}
// after cut_stats_kernel: integer sums into the stride's tail, per-length maxima behind the insert maxima (MAX-reduced together)
__global__ void shard_hdr_kernel(const CutStatsDev* st, uint32_t* send, int64_t stride, int32_t* max_best) {
  const int t = threadIdx.x;
  if (t == 0) {
    long long* h = (long long*)(send + stride - SHARD_HDR_WORDS);    // 8-byte aligned: stride and SHARD_HDR_WORDS are even
    h[0] = st->sx; h[1] = st->sy; h[2] = st->cnt; h[3] = 0;
  }
  for (int l = t; l <= MAX_READ; l += blockDim.x) max_best[l] = st->best[l];
}
__global__ void shard_merge_kernel(int world, int64_t stride, const uint32_t* recv, const int32_t* max_best, CutStatsDev* st, int32_t* pf_ids) {
  const int t = threadIdx.x;
  if (t == 0) {
    long long sx = 0, sy = 0, cnt = 0;
    for (int r = 0; r < world; r++) {
      const long long* h = (const long long*)(recv + (int64_t)(r + 1) * stride - SHARD_HDR_WORDS);
      sx += h[0]; sy += h[1]; cnt += h[2];
    }
    st->sx = sx; st->sy = sy; st->cnt = cnt;
    pf_ids[0] = 0;
  }
  for (int l = t; l <= MAX_READ; l += blockDim.x) st->best[l] = max_best[l];
}

// below = score < threshold(len) (mia.c:452-470); sticky |= below (H10); the natural entries (2i, 2i+1) of the read
// take the sticky flag (nullable)
// a read that is not unique_best is not tested (mia.c:466) and keeps its flag; newly (nullable): dropped by this round
__global__ void cut_flags_kernel(int64_t n, const int32_t* __restrict__ seq_len, const int32_t* __restrict__ score,
                                 const double* __restrict__ thr, uint8_t* sticky, miagpu_entry* entries, CutStatsDev* st,
                                 const uint8_t* __restrict__ unique = nullptr, uint8_t* newly = nullptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = seq_len[i];
  if (l < 0 || l > MAX_READ) { atomicMin(&st->bad, (long long)i); if (newly) newly[i] = 0; return; }
  const bool tested = !unique || unique[i];
  const uint8_t old = sticky[i];
  const uint8_t s = old | (uint8_t)(tested && (double)score[i] < thr[l]);
  if (newly) newly[i] = s & !old;
  sticky[i] = s;
  if (entries) { entries[2 * i].dropped = s; entries[2 * i + 1].dropped = s; }
  // how many reads this round's cut dropped (st->pad): the next round decides by it whether to accumulate before the cut is known
  const bool nw = s & !old;
  const unsigned act = __activemask(), m = __ballot_sync(act, nw);
  if (nw && (threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&st->pad, __popc(m));
}

}  // namespace miagpu
