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
constexpr int FLAT_MATCH = 200;      // params.h:28
constexpr int FLAT_MISMATCH = -600;  // params.h:29
constexpr int N_SCORE_FLAT = -100;   // params.h:30 N_SCORE
constexpr int NR_SCORE_FLAT = -10;   // params.h:31 NR_SCORE
constexpr int TRIM_SCORE_CUT = 1000; // params.h:32

// ---- kernel-side PSSM layout: prof[strand][depth][read_base][ref_code padded to 8]
// One "profile row" (the 5 scores a read base can get against A,C,G,T,other at a
// given depth) is 8 consecutive ints = 32 B, so a DP row needs ONE uniform
// offset and every lane adds its column's ref_code*4.
constexpr int PROF_ROW_INTS = 8;
constexpr int PROF_INTS = 2 * NMAT * 5 * PROF_ROW_INTS;       // 2480 ints = 9920 B (plain); a second copy * 2048 follows
__host__ __device__ inline int prof_row_index(int strand, int depth, int read_code) {
  return ((strand * NMAT + depth) * 5 + read_code) * PROF_ROW_INTS;
}

// ---- packed keys (windowed kernel).
// key = value * 2048 + (marker << 9) + (511 - index).  max() over keys picks the larger value;
// on equal values the larger marker, then the SMALLER index.  That single ordering carries
// every tie rule of dyn_prog (mia.c:838-965):
//   * candidates for best_gap_col / best_gap_row share one marker, so "earliest index wins"
//     is the reference's strict-'>' replacement (839-843, 857-861);
//   * between move types DIAG(3) > COL(2) > ROW(1) reproduces "diag if >= both, else col if >= row";
//   * START(0) with index bits 0 loses every tie, i.e. start-new needs strictly-greater (910-915).
// Scores themselves live in "diagonal key" form Sd = S*2048 + (DIAG<<9) so that no conversion is
// needed between a cell's result and the next row's operands.  |value| < 2^20, index < 512.
constexpr int KEY_SHIFT = 11;
constexpr int KEY_MUL = 1 << KEY_SHIFT;
constexpr int KEY_LOW_MASK = KEY_MUL - 1;
constexpr int KEY_IDX_MASK = 511;
constexpr int MARK_START = 0, MARK_ROW = 1, MARK_COL = 2, MARK_DIAG = 3;
constexpr int DIAG_BITS = MARK_DIAG << 9;
constexpr int NEG_VALUE = -900000;                   // "-infinity" that still packs after -(800+200*511)
constexpr int NEG_KEY = NEG_VALUE * KEY_MUL;         // -1,843,200,000
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
