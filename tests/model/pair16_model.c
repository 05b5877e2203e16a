/* pair16_model.c -- CPU model of the 16-bit "V-frame" arithmetic used by the
 * SIMD (s16x2) realign kernel, csrc/pair16.cuh.  TEST INFRASTRUCTURE: it exists
 * so that the frame algebra (row frame V = S + 200 r - OFF, constant start floor,
 * decay-free row-gap chain, clamped cross-lane scan, sentinels, range limits) can be
 * checked against the oracle on the CPU, lane for lane, before it runs on a GPU.
 * Every 16-bit operation goes through w16(), which records any wrap-around: the
 * model must be wrap-free (except the discarded sum of a start-new cell).
 *
 * Follows dyn_prog (mia.c:740-981) for the unmasked, sg5 = 1 case and
 * max_sg_score (mia.c:1278-1302); the traceback part only recognises the pure
 * diagonal path (find_align_begin mia.c:612-637 with trace == 0 all the way).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GOP 1000
#define GEP 200
#define DEPTH 15

static int g_wrap;
static int w16(int x) {
  if (x < -32768 || x > 32767) { g_wrap++; x = ((x + 32768) & 0xffff) - 32768; }
  return x;
}
static int code_of(char b) { return b == 'A' ? 0 : b == 'C' ? 1 : b == 'G' ? 2 : b == 'T' ? 3 : 4; }
static int sm_depth(int row, int len) {
  if (row < DEPTH) return row;
  int fe = len - (row + 1);
  if (fe < DEPTH) return 2 * DEPTH - fe;
  return DEPTH;
}
static int imax(int a, int b) { return a > b ? a : b; }

/* host-side constants, same formulas as miagpu.cu:pair16_limits() */
void p16_limits(const int* sm, int K, int* off16, int* lmax) {
  int mn = 0, mx = 0;
  for (int i = 0; i < 775; i++) { if (sm[i] < mn) mn = sm[i]; if (sm[i] > mx) mx = sm[i]; }
  int a = 32768 - 2 * (GOP + GEP) - GEP * (K - 1) + mn;      /* lowest intermediate: LOW - 1200 - 200(K-1) >= -32768 */
  int b = 32768 - (GOP + 2 * GEP) - GEP * K - GEP - 1;       /* sentinel  -32768 + 200K + 200  <  Ncmp = -1400 - OFF */
  int off = (a < b ? a : b) - 32;
  int inc = mx + GEP > 0 ? mx + GEP : 0;
  int lm = inc ? 1 + (32767 + off - mx) / inc : 256;
  if (lm > 256) lm = 256;
  *off16 = off;
  *lmax = lm;
}

/* out: [0] score [1] aec [2] pure diagonal (1) / needs the 32-bit kernel (0) [3] row_stop [4] col_stop [5] wraps */
int p16_model(const char* ref, int len1, const char* read, int L, const int* sm, int K, int G, int band, int diag0, int* out) {
  int OFF, LMAX;
  p16_limits(sm, K, &OFF, &LMAX);
  if (L > LMAX || len1 > G * K || L < 1) return 0;
  g_wrap = 0;
  const int NC = G * K;
  const int NCMP = -(GOP + 2 * GEP) - OFF;       /* N_r in the frame of row r-1: constant */
  const int SENT = -32768 + GEP * K + GEP;
  int* V = (int*)malloc(sizeof(int) * (size_t)L * NC);
  int* W = V;                                    /* row r-1 */
  int* R = (int*)malloc(sizeof(int) * NC);
  int* T = (int*)malloc(sizeof(int) * NC);
  int* X = (int*)malloc(sizeof(int) * G);
  int* Y = (int*)malloc(sizeof(int) * G);
  int rcode[256], ccode[1024];
  for (int r = 0; r < L; r++) rcode[r] = code_of(read[r]);
  for (int c = 0; c < NC; c++) ccode[c] = c < len1 ? code_of(ref[c]) : 4;
#define SUBP(r, c) (sm[(sm_depth(r, L) * 5 + ccode[c]) * 5 + rcode[r]] + GEP)
  for (int c = 0; c < NC; c++) { V[c] = w16(w16(-OFF - GEP) + SUBP(0, c)); R[c] = -32768; }
  for (int r = 1; r < L; r++) {
    int* Wn = V + (size_t)r * NC;
    W = V + (size_t)(r - 1) * NC;
    for (int l = 0; l < G; l++) {                /* local chains */
      for (int j = 0; j < K; j++) {
        int c = l * K + j;
        int cand = c >= 2 ? w16(W[c - 2] - (GOP + GEP)) : SENT;
        T[c] = j == 0 ? cand : imax(w16(T[c - 1] - GEP), cand);
      }
      X[l] = T[l * K + K - 1];
    }
    for (int d = 1; d < G; d <<= 1) {            /* clamped inclusive scan, decay 200 K per lane */
      memcpy(Y, X, sizeof(int) * G);
      for (int l = d; l < G; l++) {
        int y = imax(Y[l - d], -32768 + GEP * K * d);
        X[l] = imax(X[l], w16(y - GEP * K * d));
      }
    }
    for (int l = 0; l < G; l++) {
      int qin = l ? X[l - 1] : SENT;
      qin = imax(qin, -32768 + GEP * K);
      for (int j = 0; j < K; j++) {
        int c = l * K + j;
        int D = c >= 1 ? W[c - 1] : NCMP;
        int Q = imax(T[c], w16(qin - GEP * (j + 1)));
        int best = imax(imax(D, Q), R[c]);
        R[c] = imax(R[c], w16(D - GOP));
        int start = best < NCMP;
        int bp = imax(best, NCMP);
        int add = start ? GEP : SUBP(r, c);
        Wn[c] = w16(bp + add);
      }
    }
  }
  /* max_sg_score: first maximum of the last row */
  const int* last = V + (size_t)(L - 1) * NC;
  int aec = 0;
  for (int c = 1; c < len1; c++) if (last[c] > last[aec]) aec = c;
  out[0] = last[aec] + OFF - GEP * (L - 1);
  out[1] = aec;
  /* pure-diagonal verification inside the stored band */
  int d = aec - (L - 1), pure = 1, row = L - 1, col = aec;
  if (d < diag0 - band || d > diag0 + band) pure = 0;
  while (pure && row > 0 && col > 0) {
    int v = V[(size_t)row * NC + col], dv = V[(size_t)(row - 1) * NC + col - 1];
    if (v - SUBP(row, col) == dv && dv >= NCMP) { row--; col--; } else pure = 0;
  }
  out[2] = pure; out[3] = row; out[4] = col; out[5] = g_wrap;
  free(V); free(R); free(T); free(X); free(Y);
  return 1;
}
