// common.cuh -- shared device/host definitions for libmiagpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/miagpu.h"

namespace miagpu {

// ---- constants of the reference (params.h); static so the kernels fold them
constexpr int GOP = 1000;            // params.h:26
constexpr int GEP = 200;             // params.h:27
constexpr int PSSM_DEPTH = 15;       // params.h:22
constexpr int NMAT = 2 * PSSM_DEPTH + 1;
constexpr int MAX_READ = MIAGPU_MAX_READ;
constexpr int REALIGN_BUFFER = 50;   // params.h:34
constexpr int MAX_RUNS = MIAGPU_MAX_RUNS;
constexpr int HIM = -1073741824;     // INT_MIN/2, mia.c:751
constexpr int FIRST_ROUND_SCORE_CUTOFF = 2000;

// ---- kernel-side PSSM layout: prof[strand][depth][read_base][ref_code padded to 8]
// One "profile row" (the 5 scores a read base can get against A,C,G,T,other at a
// given depth) is 8 consecutive ints = 32 B, so a DP row needs ONE uniform
// offset and every lane adds its column's ref_code*4.
constexpr int PROF_ROW_INTS = 8;
constexpr int PROF_INTS = 2 * NMAT * 5 * PROF_ROW_INTS;       // 2480 ints = 9920 B
__host__ __device__ inline int prof_row_index(int strand, int depth, int read_code) {
  return ((strand * NMAT + depth) * 5 + read_code) * PROF_ROW_INTS;
}

// ---- packed arg-max keys (windowed kernel).
// key = value * 2048 + (marker << 9) + (511 - index);  max() over keys picks the
// larger value and, on equal values, the SMALLER index -- which is the
// reference's strict-'>' "earliest candidate wins" rule (mia.c:839-843, 857-861).
// value needs |value| < 2^20; index < 512.
constexpr int KEY_SHIFT = 11;
constexpr int KEY_MUL = 1 << KEY_SHIFT;
constexpr int KEY_IDX_MASK = 511;
constexpr int MARK_DIAG = 0, MARK_START = 1, MARK_COL = 2, MARK_ROW = 3;
constexpr int NEG_VALUE = -1000000;                  // "-infinity" that still packs
constexpr int NEG_KEY = NEG_VALUE * KEY_MUL;         // -2,048,000,000 > INT_MIN
constexpr int PSSM_ABS_LIMIT = 2000;                 // keeps 256*|x| + 200*511 below 2^20

__host__ __device__ inline int base_code(uint8_t b) {   // mia.c:1054-1082
  return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : 4;
}

__host__ __device__ inline int sm_depth(int row, int len) {   // pssm.c:36-46
  if (row < PSSM_DEPTH) return row;
  int from_end = len - (row + 1);
  if (from_end < PSSM_DEPTH) return 2 * PSSM_DEPTH - from_end;
  return PSSM_DEPTH;
}

// error plumbing (C ABI: 1 = ok, 0 = failure + message)
void set_error(const char* fmt, ...);
#define MIAGPU_CUDA(call)                                                              \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      miagpu::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return 0;                                                                        \
    }                                                                                  \
  } while (0)

}  // namespace miagpu
