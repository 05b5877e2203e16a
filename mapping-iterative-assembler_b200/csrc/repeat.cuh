// repeat.cuh -- the repeat filter (-u / -U): sort_fsdb / sort_fsdb_qscore (fsdb.c:13-88, 90-180, 240-252)
// followed by set_uniq_in_fsdb (fsdb.c:440-508), for read sets where the host's qsort + pointer chase is the serial tail.
//
// fs_comp orders by strand (reverse first), then forward reads by (as ascending, ae descending, key4 descending) and reverse
// reads by (ae descending, as ascending, key4 descending); key4 = score (-u) or qual_sum (-U).  glibc's qsort is a merge
// sort while its buffer fits, i.e. stable: reads that compare equal keep their FSDB order, and the FIRST of them becomes
// the unique one.  Here the comparison is one 64-bit key
//     [strand: 1][k1: 21][k2: 21][k3: 21]      k1, k2 = as / (2^21-1 - ae) (forward) or (2^21-1 - ae) / as (reverse),
//                                               k3 = 2^20-1 - key4
// sorted by a stable LSD radix sort (cub::DeviceRadixSort::SortPairs).  set_uniq_in_fsdb compares every read with the
// head of the current group of identical (strand, as, ae); with tolerance 0 identical triples are adjacent after the
// sort, so the head a group head sees is simply its predecessor: one independent test per read.  With a tolerance (-C)
// the grouping is a greedy scan and runs on the host over the sorted order.
#pragma once
#include "common.cuh"

namespace miagpu {

constexpr int RF_COORD_BITS = 21;
constexpr int RF_COORD_MAX = (1 << RF_COORD_BITS) - 1;
constexpr int RF_KEY4_HALF = 1 << 20;

__global__ void rf_key_kernel(int64_t n, const uint8_t* __restrict__ rc, const int32_t* __restrict__ as, const int32_t* __restrict__ ae,
                              const int32_t* __restrict__ key4, uint64_t* keys, int32_t* idx, int* bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int a = as[i], e = ae[i], k = key4[i];
  if (a < 0 || a > RF_COORD_MAX || e < 0 || e > RF_COORD_MAX || k < -RF_KEY4_HALF || k >= RF_KEY4_HALF) atomicExch(bad, 1);
  const bool r = rc[i] != 0;
  const uint64_t k1 = r ? (uint64_t)(RF_COORD_MAX - e) : (uint64_t)a;
  const uint64_t k2 = r ? (uint64_t)a : (uint64_t)(RF_COORD_MAX - e);
  const uint64_t k3 = (uint64_t)(RF_KEY4_HALF - 1 - k);
  keys[i] = ((uint64_t)(r ? 0 : 1) << 63) | ((k1 & RF_COORD_MAX) << 42) | ((k2 & RF_COORD_MAX) << 21) | (k3 & RF_COORD_MAX);
  idx[i] = (int32_t)i;
}

// tolerance 0: unique_best of the read at sorted position k from its predecessor (see the header)
__global__ void rf_unique_kernel(int64_t n, const int32_t* __restrict__ order, const uint8_t* __restrict__ rc, const int32_t* __restrict__ as,
                                 const int32_t* __restrict__ ae, const uint8_t* __restrict__ trimmed, int just_outer_coords,
                                 uint8_t* unique) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int f = order[k];
  if (k == 0) { unique[f] = 1; return; }
  const int g = order[k - 1];
  const bool frc = rc[f] != 0, grc = rc[g] != 0;
  uint8_t u;
  if (frc == grc && as[f] == as[g] && ae[f] == ae[g]) u = 0;
  else if (just_outer_coords) u = 1;
  else if (!frc) u = (as[f] == as[g]) ? (trimmed && trimmed[f] ? 1 : 0) : 1;
  else u = (ae[f] == ae[g]) ? (trimmed && trimmed[f] ? 1 : 0) : 1;
  unique[f] = u;
}

__global__ void rf_widen_kernel(int64_t n, const int32_t* __restrict__ order, int64_t* out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = order[k];
}

}  // namespace miagpu
