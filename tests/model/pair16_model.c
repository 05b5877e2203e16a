/* pair16_model.c -- CPU model of the 16-bit "lane frame" arithmetic used by the
 * SIMD (u16x2) realign kernel, csrc/pair16.cuh.  TEST INFRASTRUCTURE: it exists
 * so that the frame algebra can be checked against the oracle on the CPU, lane
 * for lane, before it runs on a GPU.  Every 16-bit operation goes through w16(),
 * which records any wrap-around: the model must be wrap-free.
 *
 * Lane frame.  Lane l of a group owns columns c = l*K + j, j in [0,K).  A cell is held as
 *     V(r,c) = S(r,c) + GEP*r + GEP*j - OFF
 * and a value of another lane is converted by -GEP*K per lane of distance, which extends
 * the lane's frame linearly to j < 0.  In that frame all three candidates of dyn_prog
 * (mia.c:838-871) lose their per-step decay and share one offset:
 *     diagonal    V(r-1,c-1)                         + 2*GEP
 *     column gap  max_{k<=c-2} V(r-1,k) - GOP        + 2*GEP     (a plain running maximum)
 *     row gap     max_{i<=r-2} V(i,c-1) - GOP        + 2*GEP     (a plain running maximum)
 * and the start-new candidate N_r (mia.c:877-880) is the per-column constant
 *     NCMP_j + 2*GEP,   NCMP_j = -(GOP + 3*GEP) - OFF + GEP*j.
 * The table a cell reads holds sub + 2*GEP (start-new cell: 2*GEP).
 *
 * Pure-diagonal test.  bad(r,c) = max(D, Gc, Gr, NCMP_j) != D is OR-ed down the diagonal
 * (acc(r,c) = acc(r-1,c-1) | bad(r,c), rows and columns >= 1): acc == 0 at the end cell
 * iff dyn_prog stored trace 0 in every cell find_align_begin (mia.c:612-637) visits.
 *
 * Follows dyn_prog (mia.c:740-981) for the unmasked, sg5 = 1 case and
 * max_sg_score (mia.c:1278-1302).
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
  /* lowest real intermediate: (lowest cell -OFF-GOP-GEP+mn) converted to the next lane (-GEP*K), minus GOP;
   * OFF is a compile-time constant of the kernel, taken for the most negative entry set_pssm accepts (-2000) */
  int off = 32768 - 2 * GOP - GEP - 2000 - GEP * K - 32;
  if (mn < -2000) { *off16 = off; *lmax = 0; return; }
  int inc = mx + GEP > 0 ? mx + GEP : 0;
  /* highest cell: L*mx + GEP*(L-1) + GEP*(K-1) - OFF <= 32767 */
  int lm = inc ? (32767 + off - GEP * (K - 2)) / inc : 256;
  if (lm > 256) lm = 256;
  *off16 = off;
  *lmax = lm;
}

/* out: [0] score [1] aec [2] pure diagonal (1) / needs the 32-bit kernel (0) [3] row_stop [4] col_stop [5] wraps */
int p16_model(const char* ref, int len1, const char* read, int L, const int* sm, int K, int G, int* out) {
  int OFF, LMAX;
  p16_limits(sm, K, &OFF, &LMAX);
  if (L > LMAX || len1 > G * K || L < 1) return 0;
  g_wrap = 0;
  const int NC = G * K;
  const int NCMP0 = -(GOP + 3 * GEP) - OFF;
  const int SENT = -32768 + GOP;                 /* "-infinity" that survives one -GOP */
  const int CONV = GEP * K;                      /* frame shift per lane of distance */
  int* W = (int*)malloc(sizeof(int) * NC);       /* row r-1, lane frames */
  int* Wn = (int*)malloc(sizeof(int) * NC);
  int* R = (int*)malloc(sizeof(int) * NC);
  int* A = (int*)malloc(sizeof(int) * NC);       /* acc of row r-1 */
  int* An = (int*)malloc(sizeof(int) * NC);
  int* X = (int*)malloc(sizeof(int) * G);
  int* Y = (int*)malloc(sizeof(int) * G);
  int rcode[256], ccode[1024];
  for (int r = 0; r < L; r++) rcode[r] = code_of(read[r]);
  for (int c = 0; c < NC; c++) ccode[c] = c < len1 ? code_of(ref[c]) : 4;
#define SUB2(r, c) (sm[(sm_depth(r, L) * 5 + ccode[c]) * 5 + rcode[r]] + 2 * GEP)
  for (int c = 0; c < NC; c++) {
    int j = c % K;
    W[c] = w16(w16(GEP * j - 2 * GEP - OFF) + SUB2(0, c));
    R[c] = -32768;
    A[c] = 0;
  }
  for (int r = 1; r < L; r++) {
    /* lane totals of the column-gap candidates: E[0] = l2c, E[1] = l1c, E[j] = W[j-2] */
    for (int l = 0; l < G; l++) {
      int m = -32768;
      for (int j = 0; j < K; j++) {
        int c = l * K + j;
        int e = j >= 2 ? W[c - 2] : (l ? w16(W[c - 2] - CONV) : SENT);
        m = imax(m, e);
      }
      X[l] = w16(m - GOP);
    }
    for (int d = 1; d < G; d <<= 1) {            /* clamped inclusive scan, decay GEP*K per lane */
      memcpy(Y, X, sizeof(int) * G);
      for (int l = d; l < G; l++) {
        int y = imax(Y[l - d], -32768 + CONV * d);
        X[l] = imax(X[l], w16(y - CONV * d));
      }
    }
    for (int l = 0; l < G; l++) {
      int q = -32768;
      if (l) q = w16(imax(X[l - 1], -32768 + CONV) - CONV);
      int acc_in = l ? A[l * K - 1] : 0;
      for (int j = 0; j < K; j++) {
        int c = l * K + j;
        const int NCMPj = NCMP0 + GEP * j;
        int e = j >= 2 ? W[c - 2] : (l ? w16(W[c - 2] - CONV) : SENT);
        q = imax(q, w16(e - GOP));
        int D = j >= 1 ? W[c - 1] : (l ? w16(W[c - 1] - CONV) : NCMPj);
        int best = imax(imax(D, q), R[c]);
        R[c] = imax(R[c], w16(D - GOP));
        int start = best < NCMPj;
        int bp = imax(best, NCMPj);
        int raw = w16(bp + SUB2(r, c));          /* the add is done for start-new cells too: it must not wrap either */
        Wn[c] = start ? w16(NCMPj + 2 * GEP) : raw;
        An[c] = (j >= 1 ? A[c - 1] : acc_in) | (bp != D);
        if (c == 0) An[c] = 0;
      }
    }
    int* t = W; W = Wn; Wn = t;
    t = A; A = An; An = t;
  }
  /* max_sg_score: first maximum of the last row, S = V - GEP*(L-1) - GEP*j + OFF */
  int aec = 0, best = -1000000000;
  for (int c = 0; c < len1; c++) {
    int s = W[c] - GEP * (c % K);
    if (s > best) { best = s; aec = c; }
  }
  out[0] = best + OFF - GEP * (L - 1);
  out[1] = aec;
  int pure = A[aec] == 0;
  int steps = aec < L - 1 ? aec : L - 1;
  out[2] = pure; out[3] = L - 1 - steps; out[4] = aec - steps; out[5] = g_wrap;
  free(W); free(Wn); free(R); free(A); free(An); free(X); free(Y);
  return 1;
}
