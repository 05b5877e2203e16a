// pass1.cuh -- pass 1 with the k-mer filter on: seeding and the windowed fast path (sm_100a).
//
// new_kmer_filter (kmer.c:239-331) unmasks, for every k-mer hit (read offset p, reference position q), the columns
// [q-p-10, q+(L-p)+10] (one less on the reverse strand) of that strand's matrix; sg_align (mia.c:1500-1610) then
// runs dyn_prog over both whole strands, where masked cells are HIM.  For a read whose hits on a strand all lie on
// neighbouring diagonals the unmasked columns form ONE stretch of about L+20 columns, and the masked DP over the
// whole strand equals a DP over that stretch alone (pair16.cuh, JOB kernels, explains the one cell rule that
// differs).  So:
//   p1_seed_kernel   warp per read: k-mer lookups on both strands -> hits, lowest / highest hit diagonal per strand;
//                    a strand with hits becomes a job (2*read + strand) when its stretch is one piece, at most 256
//                    columns wide and the read fits the 16-bit frame; a read whose strands with hits are all jobs
//                    goes to the pair kernels, a read without hits is skipped (mia_main.c:781), everything else
//                    (several stretches, saturated strand, long read) goes to the general kernel (strip.cuh)
//   pair16_kernel<K, G, true>   the jobs, two per register
//   p1_merge_kernel  strand pick (forward only if strictly better, mia.c:1549-1554), sg_align's coordinates
//                    (c2rcc, as / ae / start / end fix-ups, mia.c:1568-1610), or hand-over to the general kernel
//                    when a job's alignment was not one plain diagonal
#pragma once
#include "common.cuh"
#include "pair16.cuh"
#include "strip.cuh"

namespace miagpu {

// p1 meta block (int32 words) -- laid out like the realign meta block where pair_layout / pair_scatter read it
constexpr int P1_NGENERAL = 138;     // reads handed to the general kernel (list fill counter)
constexpr int P1_NFAST = 139;        // reads finished by the pair kernels
constexpr int P1_NSKIPPED = 140;     // reads without a k-mer hit
constexpr int P1_WORK = 141;         // work-fetch counter of the general kernel

struct SeedRange { int hits, dmin, dmax; };

// new_kmer_filter's hit loop for one strand (kmer.c:275-327) without the mask: hit count and diagonal range
__device__ __forceinline__ SeedRange seed_range(const KmerTable& kt, const int k, const uint8_t* __restrict__ read, const int L) {
  const int lane = threadIdx.x & 31;
  int hits = 0, dmin = INT_MAX, dmax = INT_MIN;
  for (int p = lane; p + k <= L; p += 32) {
    uint32_t inx = 0;
    bool ok = true;
    for (int i = 0; i < k; i++) {
      const int c = kmer_code(read[p + i]);
      if (c < 0) { ok = false; break; }
      inx = (inx << 2) | (uint32_t)c;
    }
    if (!ok) continue;
    const int b = (int)(inx >> kt.bucket_shift);
    const int e1 = __ldg(kt.bucket_start + b + 1);
    for (int e = __ldg(kt.bucket_start + b); e < e1; e++) {
      if (__ldg(kt.kmer + e) != inx) continue;
      hits++;
      const int d = __ldg(kt.pos + e) - p;
      dmin = min(dmin, d); dmax = max(dmax, d);
    }
  }
  SeedRange r;
  r.hits = __reduce_add_sync(0xffffffffu, hits);
  r.dmin = __reduce_min_sync(0xffffffffu, dmin);
  r.dmax = __reduce_max_sync(0xffffffffu, dmax);
  return r;
}

struct P1SeedParams {
  const uint8_t* bases;
  const int64_t* off;
  int64_t n;
  int32_t k, len1, strand_stride;
  KmerTable kt[2];
  PairLmax lm;
  // per job (2n)
  uint8_t* jkind;                // 16 + pair class, 0 = no job
  int32_t* jws;                  // first column of the stretch, as an index into the concatenated codes
  int32_t* jwl;                  // columns
  // per read
  int32_t* hits;
  uint8_t* route;                // 0 skipped, 1 pair kernels, 2 general kernel
  int32_t* general_list;
  int32_t* score; int32_t* n_runs; uint8_t* status;
  int32_t* meta;
};

__global__ void __launch_bounds__(256) p1_seed_kernel(P1SeedParams p) {
  __shared__ int s_hist[P16_KEYS];
  __shared__ int s_preads[P16_NKB];
  __shared__ unsigned long long s_pcells[P16_NKB];
  __shared__ int s_counts[3];
  for (int i = threadIdx.x; i < P16_KEYS; i += blockDim.x) s_hist[i] = 0;
  if (threadIdx.x < P16_NKB) { s_preads[threadIdx.x] = 0; s_pcells[threadIdx.x] = 0; }
  if (threadIdx.x < 3) s_counts[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t rd = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rd < p.n; rd += warps) {
    const int64_t o = p.off[rd];
    const int L = (int)(p.off[rd + 1] - o);
    if (L <= 0 || L > MAX_READ) {                                        // the general kernel reports it
      if (lane == 0) { p.route[rd] = 2; p.jkind[2 * rd] = 0; p.jkind[2 * rd + 1] = 0; p.general_list[atomicAdd(p.meta + P1_NGENERAL, 1)] = (int32_t)rd; }
      continue;
    }
    SeedRange sr[2] = {{0, 0, 0}, {0, 0, 0}};
    if (L >= p.k) {
      sr[0] = seed_range(p.kt[0], p.k, p.bases + o, L);
      sr[1] = seed_range(p.kt[1], p.k, p.bases + o, L);
    }
    if (lane != 0) continue;
    const int total = sr[0].hits + sr[1].hits;
    p.hits[rd] = total;
    int kb[2] = {-1, -1}, lo[2] = {0, 0}, wl[2] = {0, 0};
    bool fast = total > 0;
    for (int s = 0; s < 2; s++) {
      if (!sr[s].hits) continue;
      // every hit unmasks [d - 10, d + L + 10 - s] (kmer.c:294, 319): one stretch when the diagonals are close enough
      const int span = L + 2 * ALIGN_MASK_BUFFER - s;                    // last column of a hit's stretch minus its first
      int a = sr[s].dmin - ALIGN_MASK_BUFFER, z = sr[s].dmax + span - ALIGN_MASK_BUFFER;
      if (a < 0) a = 0;
      if (z >= p.len1) z = p.len1 - 1;
      const bool one_piece = sr[s].dmax - sr[s].dmin <= span + 1;
      lo[s] = a; wl[s] = z - a + 1;
      kb[s] = p16_class(wl[s]);
      if (sr[s].hits >= KMER_SATURATE || !one_piece || z < a || kb[s] < 0 || L > p.lm.v[kb[s] < 0 ? 0 : kb[s]]) fast = false;
    }
    if (!total) {
      p.route[rd] = 0;
      p.status[rd] = MIAGPU_ST_SKIPPED; p.n_runs[rd] = 0; p.score[rd] = INT_MIN;      // mia_main.c:781: not aligned at all
      p.jkind[2 * rd] = 0; p.jkind[2 * rd + 1] = 0;
      atomicAdd(&s_counts[0], 1);
    } else if (!fast) {
      p.route[rd] = 2;
      p.jkind[2 * rd] = 0; p.jkind[2 * rd + 1] = 0;
      p.general_list[atomicAdd(p.meta + P1_NGENERAL, 1)] = (int32_t)rd;
    } else {
      p.route[rd] = 1;
      for (int s = 0; s < 2; s++) {
        const int64_t j = 2 * rd + s;
        if (!sr[s].hits) { p.jkind[j] = 0; continue; }
        p.jkind[j] = (uint8_t)(16 + kb[s]);
        p.jws[j] = s * p.strand_stride + lo[s];
        p.jwl[j] = wl[s];
        atomicAdd(&s_hist[kb[s] * (P16_MAXL + 1) + L], 1);
        atomicAdd(&s_preads[kb[s]], 1);
        atomicAdd(&s_pcells[kb[s]], (unsigned long long)L * wl[s]);
      }
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P16_KEYS; k += blockDim.x)
    if (s_hist[k]) atomicAdd(&p.meta[META_HIST + k], s_hist[k]);
  if (threadIdx.x < P16_NKB && s_preads[threadIdx.x]) {
    atomicAdd(&p.meta[META_PREADS + threadIdx.x], s_preads[threadIdx.x]);
    atomicAdd(reinterpret_cast<unsigned long long*>(p.meta + META_PCELLS) + threadIdx.x, s_pcells[threadIdx.x]);
  }
  if (threadIdx.x == 0 && s_counts[0]) atomicAdd(&p.meta[P1_NSKIPPED], s_counts[0]);
}

struct P1MergeParams {
  int64_t n;
  const int64_t* off;
  int32_t seq_len;
  const uint8_t* route;
  const uint8_t* jkind; const uint8_t* jstatus;
  const int32_t* jscore; const int32_t* jabc; const int32_t* jaec; const int32_t* jabr;
  int32_t* general_list;
  int32_t* meta;
  int32_t *score, *fw_score, *rc_score, *as_out, *ae_out, *start, *end, *abr, *n_runs;
  uint8_t* rc_out;
  uint16_t* runs;
  uint8_t* status;
};

__global__ void p1_merge_kernel(P1MergeParams p) {
  const int64_t rd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rd >= p.n || p.route[rd] != 1) return;
  int best[2] = {HIM, HIM};                         // a strand without a hit is masked everywhere: max_sg_score finds HIM in column 0
  bool general = false;
  for (int s = 0; s < 2; s++) {
    const int64_t j = 2 * rd + s;
    if (!p.jkind[j]) continue;
    if (p.jstatus[j] != MIAGPU_ST_OK) general = true;
    else best[s] = p.jscore[j];
  }
  if (general) {
    p.general_list[atomicAdd(p.meta + P1_NGENERAL, 1)] = (int32_t)rd;
    return;
  }
  atomicAdd(p.meta + P1_NFAST, 1);
  const int s = !(best[0] > best[1]) ? 1 : 0;       // forward only if strictly better (mia.c:1549-1554)
  const int64_t j = 2 * rd + s;
  const int L = (int)(p.off[rd + 1] - p.off[rd]);
  const int abc = p.jabc[j], aec = p.jaec[j], abr = p.jabr[j];
  int start = abc, end = aec;
  if (s == 1) {                                     // c2rcc, mia.c:26-30
    start = p.seq_len - (aec % p.seq_len) - 1;
    end = p.seq_len - (abc % p.seq_len) - 1;
  }
  int as = start, ae = end;
  if (as > ae) ae = p.seq_len + as;                 // mia.c:1600-1604
  if (end > p.seq_len) end -= p.seq_len;            // mia.c:1606-1610
  p.as_out[rd] = as; p.ae_out[rd] = ae; p.start[rd] = start; p.end[rd] = end;
  p.rc_out[rd] = (uint8_t)s;
  p.fw_score[rd] = best[0]; p.rc_score[rd] = best[1];
  p.abr[rd] = s == 1 ? 0 : abr;                     // row of the run's first base in the STORED orientation
  p.score[rd] = best[s];
  p.n_runs[rd] = 1;
  p.runs[rd * MAX_RUNS] = (uint16_t)((MIAGPU_RUN_M << 14) | (L - abr));
  p.status[rd] = MIAGPU_ST_OK;
}

}  // namespace miagpu
