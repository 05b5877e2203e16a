/* mia_oracle.c -- CPU restatement of MIA's hot path.  TEST INFRASTRUCTURE ONLY.
 * See mia_oracle.h for the contract and the parity status (PINNED against
 * oracle/_ref by tests/test_oracle_vs_ref.py).  Citations are
 * /root/reference/src/<file>:<line>.
 *
 * This is a restatement, not a transcription: the DP is written in the
 * "running arg-max of S + GEP*index" form the CUDA kernels use, which is
 * algebraically the reference's "compare new gap option with previous best"
 * form (mia.c:838-868):
 *     S[k] - P(1) > S[b] - P(c-b-1)   <=>   S[k] + GEP*k > S[b] + GEP*b ,  k = c-2
 * with P(n) = GOP + GEP*n.  The test-suite proves the equivalence empirically.
 */
#define _POSIX_C_SOURCE 200809L
#include "mia_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <ctype.h>

#define SM(sm,d,i,j) ((sm)[((d)*5 + (i))*5 + (j)])

/* ------------------------------------------------------------------ a1 */
void orc_flat_pssm( int* sm ) {                       /* pssm.c:96-126 */
  for ( int d = 0; d < ORC_NMAT; d++ )
    for ( int i = 0; i < 5; i++ )
      for ( int j = 0; j < 5; j++ ) {
        int v;
        if ( i == 4 ) v = -10;                        /* NR_SCORE: ref non-ACGT */
        else if ( j == 4 ) v = -100;                  /* N_SCORE: read non-ACGT */
        else v = ( i == j ) ? 200 : -600;             /* FLAT_MATCH / FLAT_MISMATCH */
        SM(sm,d,i,j) = v;
      }
}

/* io.c:408-503.  31 blocks "# Matrix for position..." + 4 lines of 4 ints +
 * blank line; block 15 must say MIDDLE. */
int orc_parse_pssm( const char* text, int* sm ) {
  const char* p = text;
  for ( int d = 0; d < ORC_NMAT; d++ ) {
    const char* eol = strchr( p, '\n' );
    size_t n = eol ? (size_t)(eol - p) : strlen( p );
    char head[128];
    if ( n >= sizeof(head) ) n = sizeof(head) - 1;
    memcpy( head, p, n ); head[n] = '\0';
    if ( d == ORC_PSSM_DEPTH ) { if ( !strstr( head, "# Matrix for position: MIDDLE" ) ) return 0; }
    else if ( !strstr( head, "# Matrix for position" ) ) return 0;
    if ( !eol ) return 0;
    p = eol + 1;
    for ( int i = 0; i < 4; i++ ) {
      int v[4];
      if ( sscanf( p, "%d\t%d\t%d\t%d", &v[0], &v[1], &v[2], &v[3] ) != 4 ) return 0;
      for ( int j = 0; j < 4; j++ ) SM(sm,d,i,j) = v[j];
      SM(sm,d,i,4) = -100;
      eol = strchr( p, '\n' );
      if ( !eol ) return 0;
      p = eol + 1;
    }
    for ( int j = 0; j < 5; j++ ) SM(sm,d,4,j) = -10;
    eol = strchr( p, '\n' );                          /* blank separator */
    p = eol ? eol + 1 : p + strlen( p );
  }
  return 1;
}

void orc_revcom_pssm( const int* in, int* out ) {     /* pssm.c:53-93 */
  for ( int d = 0; d < ORC_NMAT; d++ ) {
    int rd = ORC_NMAT - 1 - d;
    for ( int i = 0; i < 5; i++ )
      for ( int j = 0; j < 5; j++ ) {
        int ci = ( i < 4 ) ? 3 - i : 4;
        int cj = ( j < 4 ) ? 3 - j : 4;
        SM(out,rd,i,j) = SM(in,d,ci,cj);
      }
  }
}

int orc_sm_depth( int row, int len ) {                /* pssm.c:36-46 */
  if ( row < ORC_PSSM_DEPTH ) return row;
  int from_end = len - ( row + 1 );
  if ( from_end < ORC_PSSM_DEPTH ) return 2*ORC_PSSM_DEPTH - from_end;
  return ORC_PSSM_DEPTH;
}

int orc_base_code( char b ) {                         /* mia.c:1054-1082, 1243-1268 */
  switch ( b ) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; }
  return 4;
}

char orc_revcom_char( char b ) {                      /* map_align.c:418-431 */
  static const char* from = "ABCDGHKMNRSTUVWXY";
  static const char* to   = "TVGHCDMKNYSAABWXR";
  if ( b == '-' ) return '-';
  int lower = ( b >= 'a' && b <= 'z' );
  char u = lower ? (char)( b - 32 ) : b;
  const char* q = ( u >= 'A' && u <= 'Z' ) ? strchr( from, u ) : NULL;
  if ( !q || !*q ) return 'N';
  return lower ? (char)( to[q - from] + 32 ) : to[q - from];
}

/* --------------------------------------------------------------- a5-a7 */
static int *g_S = NULL, *g_T = NULL;
static size_t g_cells = 0;

/* pop_hpl_and_hps, mia.c:1193-1234: per position the length and the start of its homopolymer (raw characters compared) */
void orc_hp_runs( const char* seq, int len, int* hpl, int* hps ) {
  int start = 0;
  for ( int i = 0; i < len; i++ ) {
    if ( i > 0 && seq[i] != seq[i-1] ) {
      for ( int j = start; j < i; j++ ) hpl[j] = i - start;
      start = i;
    }
    hps[i] = start;
  }
  for ( int j = start; j < len; j++ ) hpl[j] = len - start;
}
/* hp_discount_penalty, mia.c:1096-1134: int penalty = GEP * gap_len, then `penalty += GOP * f` in double, truncated on the
   way back into the int; the first length is not used */
int orc_hp_penalty( int gap_len, int hplen2 ) {
  static const double f[11] = { 0.10, 1.0, 0.5, 0.33, 0.25, 0.2, 0.17, 0.14, 0.13, 0.11, 0.10 };
  int penalty = ORC_GEP * gap_len;
  if ( hplen2 == 1 ) { penalty += ORC_GOP; return penalty; }
  double fac = ( hplen2 >= 2 && hplen2 <= 10 ) ? f[hplen2] : 0.10;
  penalty = (int)( (double)penalty + ORC_GOP * fac );
  return penalty;
}

int orc_align( const char* seq1, int len1, const char* seq2, int len2,
               const unsigned char* mask, const int* sm, int sg5,
               int* out5, char* ref_gapped, char* read_gapped,
               int* score_mat, int* trace_mat ) {
  return orc_align_hp( seq1, len1, seq2, len2, mask, sm, sg5, 0, out5, ref_gapped, read_gapped, score_mat, trace_mat );
}

/* hp != 0: mia -h, the two homopolymer-discounted gap candidates of mia.c:882-905 and their place in the cascade (910-964) */
int orc_align_hp( const char* seq1, int len1, const char* seq2, int len2,
                  const unsigned char* mask, const int* sm, int sg5, int hp,
                  int* out5, char* ref_gapped, char* read_gapped,
                  int* score_mat, int* trace_mat ) {
  if ( len1 <= 0 || len2 <= 0 || len2 > ORC_MAX_READ ) return 0;
  int *hpcl = NULL, *hpcs = NULL, hprl[ORC_MAX_READ], hprs[ORC_MAX_READ];
  if ( hp ) {
    hpcl = (int*)malloc( sizeof(int) * len1 ); hpcs = (int*)malloc( sizeof(int) * len1 );
    orc_hp_runs( seq1, len1, hpcl, hpcs );
    orc_hp_runs( seq2, len2, hprl, hprs );
  }
  size_t need = (size_t)len1 * len2;
  if ( need > g_cells ) {
    free( g_S ); free( g_T );
    g_S = (int*)malloc( need * sizeof(int) );
    g_T = (int*)malloc( need * sizeof(int) );
    g_cells = need;
    if ( !g_S || !g_T ) { g_cells = 0; return 0; }
  }
  int* S = g_S; int* T = g_T;
  unsigned char* c1 = (unsigned char*)malloc( len1 );
  /* per column: running arg-max over rows of S[j][c] + GEP*j  (best_gap_row) */
  int* RI = (int*)malloc( sizeof(int) * len1 );
  for ( int c = 0; c < len1; c++ ) c1[c] = (unsigned char)orc_base_code( seq1[c] );
#define M(c) ( mask == NULL || mask[c] )
  /* row 0: mia.c:769-785 */
  {
    int b = orc_base_code( seq2[0] );
    for ( int c = 0; c < len1; c++ ) {
      S[c] = M(c) ? SM(sm,0,c1[c],b) : ORC_HIM;
      T[c] = 0;
      RI[c] = 0;                 /* best_gap_row[c] = 0 */
    }
  }
  for ( int r = 1; r < len2; r++ ) {
    int d = orc_sm_depth( r, len2 );
    int b = orc_base_code( seq2[r] );
    int* Sr = S + (size_t)r * len1; int* Tr = T + (size_t)r * len1;
    const int* Sp = Sr - len1;                        /* row r-1 */
    int start_new = sg5 ? -( ORC_GOP + ORC_GEP * ( r + 1 ) ) : 0;   /* mia.c:877-880 */
    /* col 0: mia.c:805-822 */
    Sr[0] = M(0) ? SM(sm,d,c1[0],b) + start_new : ORC_HIM;
    Tr[0] = 0;
    int bgc = 0;                                      /* mia.c:825 */
    for ( int c = 1; c < len1; c++ ) {
      if ( !M(c) ) { Sr[c] = ORC_HIM; Tr[c] = 0; continue; }  /* mia.c:967-970 */
      int sub = SM(sm,d,c1[c],b);
      long long gc = ORC_HIM, gr = ORC_HIM;
      if ( c >= 2 ) {                                 /* mia.c:838-850 */
        if ( (long long)Sp[c-2] + (long long)ORC_GEP*(c-2) > (long long)Sp[bgc] + (long long)ORC_GEP*bgc ) bgc = c - 2;
        gc = (long long)Sp[bgc] - ( ORC_GOP + (long long)ORC_GEP * ( c - bgc - 1 ) );
      }
      if ( r >= 2 ) {                                 /* mia.c:856-868 */
        int br = RI[c-1];
        const int* Sq = S + (size_t)( r - 2 ) * len1;
        if ( (long long)Sq[c-1] + (long long)ORC_GEP*(r-2) > (long long)S[(size_t)br*len1 + c-1] + (long long)ORC_GEP*br ) { br = r - 2; RI[c-1] = br; }
        gr = (long long)S[(size_t)br*len1 + c-1] - ( ORC_GOP + (long long)ORC_GEP * ( r - br - 1 ) );
      }
      long long dg = Sp[c-1];
      long long hc = ORC_HIM, hr = ORC_HIM;           /* mia.c:882-905 */
      if ( hp && seq1[c] == seq2[r] ) {
        if ( hprs[r] == r && hpcs[c] != c && hpcs[c] > 0 )
          hc = (long long)Sp[hpcs[c]-1] - orc_hp_penalty( c - hpcs[c], hprl[r] );
        if ( hpcs[c] == c && hprs[r] != r && hprs[r] > 0 )
          hr = (long long)S[(size_t)( hprs[r] - 1 ) * len1 + c-1] - orc_hp_penalty( c - hpcs[c], hprl[r] );   /* the COLUMN distance, as written: 0 here */
      }
      if ( start_new > dg && start_new > gc && start_new > gr && start_new > hc && start_new > hr ) { Sr[c] = start_new; Tr[c] = c; }   /* mia.c:910-918 */
      else if ( dg >= gc && dg >= gr && dg >= hc && dg >= hr ) { Sr[c] = (int)( sub + dg ); Tr[c] = 0; }   /* 922-929 */
      else if ( gc >= gr && gc >= hc && gc >= hr ) { Sr[c] = (int)( sub + gc ); Tr[c] = bgc; }             /* 933-939 */
      else if ( gr >= hc && gr >= hr ) { Sr[c] = (int)( sub + gr ); Tr[c] = -RI[c-1]; }                    /* 942-948 */
      else if ( hc >= hr ) { Sr[c] = (int)( sub + hc ); Tr[c] = hpcs[c] - 1; }                             /* 950-955 */
      else                 { Sr[c] = (int)( sub + hr ); Tr[c] = -( hprs[r] - 1 ); }                        /* 956-961 */
    }
    /* mia.c:975-979 writes the sg3 penalty to column len1, outside the matrix: no effect */
  }
#undef M
  /* a6: first maximum of the last row, mia.c:1278-1302 */
  int best = INT_MIN, aec = 0, aer = len2 - 1;
  { const int* Sl = S + (size_t)aer * len1;
    for ( int c = 0; c < len1; c++ ) if ( Sl[c] > best ) { best = Sl[c]; aec = c; } }
  /* a7: mia.c:612-637 + 1440-1497 */
  char ras[2*ORC_MAX_READ + 1], fas[2*ORC_MAX_READ + 1];
  int ri = 2*ORC_MAX_READ, fi = 2*ORC_MAX_READ;
  ras[ri] = fas[fi] = '\0';
  int row = aer, col = aec, ok = 1;
  for (;;) {
    int t = T[(size_t)row*len1 + col];
    if ( t == col || t == -row ) break;
    if ( ri < 2 ) { ok = 0; break; }                 /* would overflow the reference's 512-char buffers */
    ras[--ri] = seq1[col]; fas[--fi] = seq2[row];
    if ( t == 0 ) { row--; col--; }
    else if ( t < 0 ) { int nr = -t; row--; col--; while ( row > nr && ri > 1 ) { fas[--fi] = seq2[row--]; ras[--ri] = '-'; } }
    else { int nc = t; row--; col--; while ( col > nc && ri > 1 ) { fas[--fi] = '-'; ras[--ri] = seq1[col--]; } }
  }
  ras[--ri] = seq1[col]; fas[--fi] = seq2[row];
  out5[0] = best; out5[1] = row; out5[2] = col; out5[3] = aer; out5[4] = aec;
  if ( ref_gapped )  strcpy( ref_gapped, ras + ri );
  if ( read_gapped ) strcpy( read_gapped, fas + fi );
  if ( score_mat ) memcpy( score_mat, S, need * sizeof(int) );
  if ( trace_mat ) memcpy( trace_mat, T, need * sizeof(int) );
  free( c1 ); free( RI ); free( hpcl ); free( hpcs );
  return ok;
}

/* ------------------------------------------------------------- a2, a3 */
struct orc_kmer {
  int k;
  long long n;                 /* stored (kmer,pos) pairs, sorted by kmer then pos */
  unsigned long long* e;       /* kmer << 32 | pos */
};

static int kmer_index( const char* s, int k, unsigned long long* inx ) {   /* kmer.c:18-48 */
  unsigned long long v = 0;
  for ( int i = 0; i < k; i++ ) {
    int code;
    switch ( toupper( (unsigned char)s[i] ) ) {
      case 'A': code = 0; break; case 'C': code = 1; break;
      case 'G': code = 2; break; case 'T': code = 3; break;
      default: return 0;
    }
    v = ( v << 2 ) | (unsigned)code;
  }
  *inx = v;
  return 1;
}

static int cmp_u64( const void* a, const void* b ) {
  unsigned long long x = *(const unsigned long long*)a, y = *(const unsigned long long*)b;
  return ( x > y ) - ( x < y );
}

orc_kmer* orc_kmer_build( const char* seq, long long len, int k, int soft_mask ) {  /* kmer.c:153-168 */
  orc_kmer* t = (orc_kmer*)calloc( 1, sizeof(orc_kmer) );
  t->k = k;
  long long cap = len > 0 ? len : 1;
  unsigned long long* all = (unsigned long long*)malloc( sizeof(unsigned long long) * cap );
  long long n = 0;
  for ( long long i = 0; i + k <= len; i++ ) {
    if ( soft_mask ) {                               /* all_upper, kmer.c:140-148 */
      int ok = 1;
      for ( int j = 0; j < k; j++ ) if ( islower( (unsigned char)seq[i+j] ) ) { ok = 0; break; }
      if ( !ok ) continue;
    }
    unsigned long long inx;
    if ( kmer_index( seq + i, k, &inx ) ) all[n++] = ( inx << 32 ) | (unsigned long long)i;
  }
  qsort( all, n, sizeof(unsigned long long), cmp_u64 );
  /* keep the first MAX_KMER_POS=128 positions of each k-mer (add_kmer, kmer.c:63-85) */
  long long m = 0, run = 0;
  for ( long long i = 0; i < n; i++ ) {
    if ( i > 0 && ( all[i] >> 32 ) == ( all[i-1] >> 32 ) ) run++; else run = 0;
    if ( run < 128 ) all[m++] = all[i];
  }
  t->n = m; t->e = all;
  return t;
}

void orc_kmer_free( orc_kmer* t ) { if ( t ) { free( t->e ); free( t ); } }

static long long kmer_lower( const orc_kmer* t, unsigned long long inx ) {
  long long lo = 0, hi = t->n;
  unsigned long long key = inx << 32;
  while ( lo < hi ) { long long mid = ( lo + hi ) / 2; if ( t->e[mid] < key ) lo = mid + 1; else hi = mid; }
  return lo;
}

int orc_kmer_lookup( const orc_kmer* t, long long inx, unsigned int* out ) {
  long long i = kmer_lower( t, (unsigned long long)inx );
  int n = 0;
  while ( i < t->n && (long long)( t->e[i] >> 32 ) == inx ) { if ( out ) out[n] = (unsigned int)( t->e[i] & 0xffffffffu ); n++; i++; }
  return n;
}

static void unmask( unsigned char* m, int lo, int hi, int len1 ) {
  if ( lo < 0 ) lo = 0;
  if ( hi >= len1 ) hi = len1 - 1;
  if ( hi >= lo ) memset( m + lo, 1, hi - lo + 1 );
}

unsigned orc_kmer_filter( const orc_kmer* f, const orc_kmer* r, int k,
                          const char* read, int L, int len1,
                          unsigned char* mask_f, unsigned char* mask_r ) {   /* kmer.c:239-331 */
  unsigned nf = 0, nr = 0;
  if ( k < 0 ) { memset( mask_f, 1, len1 ); return 1; }   /* rc mask untouched: kmer.c:251-255 */
  memset( mask_f, 0, len1 ); memset( mask_r, 0, len1 );
  if ( L < k ) return 0;
  for ( int p = 0; p + k <= L; p++ ) {
    unsigned long long inx;
    if ( !kmer_index( read + p, k, &inx ) ) continue;
    for ( int strand = 0; strand < 2; strand++ ) {
      const orc_kmer* t = strand ? r : f;
      unsigned char* m = strand ? mask_r : mask_f;
      long long i = kmer_lower( t, inx ), j = i;
      while ( j < t->n && ( t->e[j] >> 32 ) == inx ) j++;
      if ( j == i ) continue;
      if ( strand ) { nr += (unsigned)( j - i ); if ( nr >= 128 ) memset( m, 1, len1 ); }   /* KMER_SATURATE */
      else          { nf += (unsigned)( j - i ); if ( nf >= 128 ) memset( m, 1, len1 ); }
      for ( ; i < j; i++ ) {
        int q = (int)( t->e[i] & 0xffffffffu );
        /* fwd upper bound q+(L-p)+10, rc upper bound q+L-p-1+10: kmer.c:294 vs 319 */
        unmask( m, q - p - 10, q + ( L - p ) + 10 - ( strand ? 1 : 0 ), len1 );
      }
    }
  }
  return nf + nr;
}

/* ------------------------------------------------------------- context */
struct orc_ctx {
  char *seq, *rcseq;           /* upper-cased, wrapped, NUL-terminated */
  int seq_len, wrap_len, circular, k, distant_ref, hp;
  orc_kmer *fk, *rk;
  int smf[ORC_PSSM_INTS], smr[ORC_PSSM_INTS];
};

orc_ctx* orc_ctx_new( const char* seq, int seq_len, int circular, int with_rc,
                      int k, int soft_mask, const int* sm_fwd, int distant_ref ) {
  orc_ctx* c = (orc_ctx*)calloc( 1, sizeof(orc_ctx) );
  int wrap = circular ? ( seq_len < ORC_MAX_READ ? seq_len : ORC_MAX_READ ) : 0;   /* mia.c:657-689 */
  c->seq_len = seq_len; c->wrap_len = seq_len + wrap; c->circular = circular;
  c->k = k; c->distant_ref = distant_ref;
  c->seq = (char*)malloc( c->wrap_len + 1 );
  memcpy( c->seq, seq, seq_len ); memcpy( c->seq + seq_len, seq, wrap ); c->seq[c->wrap_len] = '\0';
  if ( with_rc ) {                                     /* io.c:388-398 then wrap */
    c->rcseq = (char*)malloc( c->wrap_len + 1 );
    for ( int i = 0; i < seq_len; i++ ) c->rcseq[i] = orc_revcom_char( seq[seq_len - 1 - i] );
    memcpy( c->rcseq + seq_len, c->rcseq, wrap ); c->rcseq[c->wrap_len] = '\0';
  }
  if ( k > 0 ) {                                       /* mia_main.c:659-672: BEFORE upper-casing */
    c->fk = orc_kmer_build( c->seq, c->wrap_len, k, soft_mask );
    if ( with_rc ) c->rk = orc_kmer_build( c->rcseq, c->wrap_len, k, soft_mask );
  }
  for ( int i = 0; i < c->wrap_len; i++ ) {            /* mia.c:642-648 */
    c->seq[i] = (char)toupper( (unsigned char)c->seq[i] );
    if ( with_rc ) c->rcseq[i] = (char)toupper( (unsigned char)c->rcseq[i] );
  }
  memcpy( c->smf, sm_fwd, sizeof(c->smf) );
  orc_revcom_pssm( c->smf, c->smr );
  return c;
}

void orc_ctx_set_hp( orc_ctx* c, int hp ) { c->hp = hp; }    /* mia -h: init_alignment( ..., hp_special ), mia_main.c:683-690 */

void orc_ctx_free( orc_ctx* c ) {
  if ( !c ) return;
  free( c->seq ); free( c->rcseq ); orc_kmer_free( c->fk ); orc_kmer_free( c->rk ); free( c );
}
int orc_ctx_wrap_len( const orc_ctx* c ) { return c->wrap_len; }
const char* orc_ctx_seq( const orc_ctx* c ) { return c->seq; }
const char* orc_ctx_rcseq( const orc_ctx* c ) { return c->rcseq; }

static void revcom_str( char* s ) {                   /* revcom_PWAF, map_align.c:512-534 */
  int n = (int)strlen( s );
  for ( int i = 0; i < n / 2; i++ ) {
    char a = s[i], b = s[n-1-i];
    s[i] = orc_revcom_char( b ); s[n-1-i] = orc_revcom_char( a );
  }
  if ( n & 1 ) s[n/2] = orc_revcom_char( s[n/2] );
}

/* split_pwaln, mia.c:1376-1438: cut both strings after the ref base seq_len-1 */
static void split_strings( char* f_ref, char* f_frag, char* b_ref, char* b_frag, int start, int wrap_point ) {
  int ref_pos = start, aln = 0;
  while ( ref_pos < wrap_point ) { if ( f_ref[aln] != '-' ) ref_pos++; aln++; }
  strcpy( b_ref, f_ref + aln ); strcpy( b_frag, f_frag + aln );
  f_ref[aln] = '\0'; f_frag[aln] = '\0';
}

/* ------------------------------------------------------------------ a8 */
int orc_pass1( const orc_ctx* c, const char* read, int L, int* out,
               char* f_ref, char* f_frag, char* b_ref, char* b_frag,
               unsigned char* mask_f, unsigned char* mask_r ) {
  int len1 = c->circular ? c->wrap_len : c->seq_len;  /* mia_main.c:721-728 */
  unsigned char *mf = mask_f, *mr = mask_r;
  for ( int i = 0; i < 18; i++ ) out[i] = 0;
  f_ref[0] = f_frag[0] = b_ref[0] = b_frag[0] = '\0';
  if ( !mf ) mf = (unsigned char*)malloc( len1 );
  if ( !mr ) mr = (unsigned char*)malloc( len1 );
  if ( c->k > 0 ) out[0] = (int)orc_kmer_filter( c->fk, c->rk, c->k, read, L, len1, mf, mr );
  else { memset( mf, 1, len1 ); memset( mr, 1, len1 ); out[0] = 1; }
  if ( out[0] ) {
    int of[5], orc[5];
    char fr[ORC_ALN_STR], ff[ORC_ALN_STR], rr[ORC_ALN_STR], rf[ORC_ALN_STR];
    /* both strands with the FORWARD matrix, sg5 = 1: mia_main.c:788-789, mia.c:1535-1542 */
    orc_align_hp( c->seq,   len1, read, L, mf, c->smf, 1, c->hp, of,  fr, ff, NULL, NULL );   /* hp arrays over the whole strands: mia_main.c:735-739 */
    orc_align_hp( c->rcseq, len1, read, L, mr, c->smf, 1, c->hp, orc, rr, rf, NULL, NULL );
    int rc = !( of[0] > orc[0] );                     /* tie -> rc, mia.c:1549-1554 */
    const int* b = rc ? orc : of;
    strcpy( f_ref, rc ? rr : fr ); strcpy( f_frag, rc ? rf : ff );
    int start = b[2], end = b[4], as, ae;
    if ( rc ) {                                       /* mia.c:1576-1595, c2rcc mia.c:26-30 */
      revcom_str( f_ref ); revcom_str( f_frag );
      start = c->seq_len - ( b[4] % c->seq_len ) - 1;
      end   = c->seq_len - ( b[2] % c->seq_len ) - 1;
    }
    as = start; ae = end;
    if ( as > ae ) ae = c->seq_len + as;              /* mia.c:1600-1604 */
    if ( end > c->seq_len ) end -= c->seq_len;        /* mia.c:1606-1610 */
    out[2] = b[0]; out[3] = rc; out[4] = as; out[5] = ae;
    out[7] = of[0]; out[8] = orc[0];
    out[14] = b[1]; out[15] = b[2]; out[16] = b[3]; out[17] = b[4];
    out[9] = start; out[10] = end;
    if ( b[0] >= 2000 || c->distant_ref ) {           /* FIRST_ROUND_SCORE_CUTOFF, mia.c:1614 */
      out[1] = 1;
      out[6] = ( b[0] > 2000 );                       /* strand_known, mia.c:1653 */
      if ( start > end ) {                            /* mia.c:1619-1634 */
        split_strings( f_ref, f_frag, b_ref, b_frag, start, c->seq_len );
        out[11] = 1; out[12] = 0; out[13] = end;
        out[10] = c->seq_len - 1;
      }
    }
  }
  if ( !mask_f ) free( mf );
  if ( !mask_r ) free( mr );
  return 1;
}

/* ------------------------------------------------------------------ a9 */
int orc_realign( const orc_ctx* c, const char* read, int L, int rc,
                 int as, int ae, int* out, char* ref_gapped, char* read_gapped ) {
  int ref_start = ( as - 50 < 0 ) ? 0 : as - 50;                       /* mia_main.c:191-196 */
  int ref_end = ( ae + 50 + 1 > c->wrap_len ) ? c->wrap_len : ae + 50; /* 197-203 */
  if ( ref_start + L > ref_end ) { ref_start = 0; ref_end = c->wrap_len; }   /* 209-212 */
  int o[5];
  if ( !orc_align_hp( c->seq + ref_start, ref_end - ref_start, read, L, NULL,
                      rc ? c->smr : c->smf, 1, c->hp, o, ref_gapped, read_gapped, NULL, NULL ) ) return 0;   /* hp arrays over the window: mia_main.c:221-224 */
  out[0] = o[0]; out[1] = o[2] + ref_start; out[2] = o[4] + ref_start;       /* 250-257 */
  out[3] = o[1]; out[4] = o[2]; out[5] = o[3]; out[6] = o[4]; out[7] = ref_start;
  return 1;
}

/* ------------------------------------------------------------ a10-a13 */
/* AlnSeq slots are PERSISTENT objects, exactly like maln->AlnSeqArray[k]
 * (map_alignment.c:38-65): a round re-uses slot k for the k-th merged segment
 * and copies every field except `dropped` (map_align.c:885-893, H10), and a
 * read's back_asp pointer is never cleared by reiterate_assembly
 * (mia_main.c:265-276) -- so a read that was wrap-split once keeps pointing at
 * its old back slot ("stale back"), whatever that slot holds now.  pop_smp,
 * cull and the consensus all follow those pointers; so do we. */
typedef struct {
  int start, end, score, revcom, dropped, segment;
  char seq[ORC_ALN_STR], smp[ORC_ALN_STR];
  char* ins[ORC_ALN_STR];
} orc_slot;

struct orc_asm {
  int seq_len, wrap_len, n, cap;
  orc_slot** s;
  int* gaps;                   /* wrap_len+1 (+slack) */
  int ne, ecap;                /* culled entry list = culled_maln->AlnSeqArray */
  int* e;
};

orc_asm* orc_asm_new( void ) { return (orc_asm*)calloc( 1, sizeof(orc_asm) ); }

void orc_asm_free( orc_asm* a ) {
  if ( !a ) return;
  for ( int i = 0; i < a->cap; i++ ) if ( a->s[i] ) { for ( int j = 0; j < ORC_ALN_STR; j++ ) free( a->s[i]->ins[j] ); free( a->s[i] ); }
  free( a->s ); free( a->gaps ); free( a->e ); free( a );
}

/* mia_main.c:43-106: new reference, gaps zeroed, inserts of the previous
 * round's slots freed, num_aln_seqs = 0 */
void orc_asm_begin_round( orc_asm* a, int seq_len, int wrap_len ) {
  for ( int i = 0; i < a->n; i++ )
    for ( int j = 0; j < ORC_ALN_STR; j++ ) { free( a->s[i]->ins[j] ); a->s[i]->ins[j] = NULL; }
  a->n = 0; a->ne = 0;
  a->seq_len = seq_len; a->wrap_len = wrap_len;
  free( a->gaps );
  a->gaps = (int*)calloc( wrap_len + 1 + 2*ORC_MAX_READ, sizeof(int) );
}

/* merge_pwaln_into_maln, map_align.c:866-954 */
static int merge_one( orc_asm* a, const char* rg, const char* fg,
                      int start, int end, int revcom, int score, int segment ) {
  if ( a->n == a->cap ) {
    int nc = a->cap ? 2*a->cap : 1024;
    a->s = (orc_slot**)realloc( a->s, sizeof(orc_slot*) * nc );
    for ( int i = a->cap; i < nc; i++ ) a->s[i] = NULL;
    a->cap = nc;
  }
  if ( !a->s[a->n] ) a->s[a->n] = (orc_slot*)calloc( 1, sizeof(orc_slot) );
  orc_slot* s = a->s[a->n];
  s->start = start; s->end = end; s->score = score; s->revcom = revcom; s->segment = segment;
  int n = (int)strlen( fg ), pos = 0, run = 0, this_gaps[ORC_ALN_STR + 1];
  char buf[ORC_ALN_STR];
  this_gaps[0] = 0;
  for ( int i = 0; i < n; i++ ) {
    if ( rg[i] == '-' ) { this_gaps[pos]++; buf[run++] = fg[i]; }
    else {
      free( s->ins[pos] ); s->ins[pos] = NULL;
      if ( run ) { buf[run] = '\0'; s->ins[pos] = strdup( buf ); run = 0; }
      s->seq[pos++] = fg[i];
      this_gaps[pos] = 0;
    }
  }
  s->seq[pos] = '\0';
  for ( int i = 0; i < end - start + 1; i++ )
    if ( this_gaps[i] > a->gaps[start + i] ) a->gaps[start + i] = this_gaps[i];
  return a->n++;
}

/* mia_main.c:259-276 / mia.c:1606-1643.  *back_slot is written only when the
 * alignment is split; the caller decides what "not split" means for its
 * pointer (pass 1 clears it, mia.c:1642; iterations leave it, mia_main.c:273-276) */
int orc_asm_add( orc_asm* a, const char* ref_gapped, const char* read_gapped,
                 int start, int end, int revcom, int score, int* front_slot, int* back_slot ) {
  if ( end > a->seq_len ) end -= a->seq_len;          /* mia_main.c:259-263 */
  if ( start > end ) {                                /* 265-272 */
    char fr[ORC_ALN_STR], ff[ORC_ALN_STR], br[ORC_ALN_STR], bf[ORC_ALN_STR];
    strcpy( fr, ref_gapped ); strcpy( ff, read_gapped );
    split_strings( fr, ff, br, bf, start, a->seq_len );
    *front_slot = merge_one( a, fr, ff, start, a->seq_len - 1, revcom, score, 'f' );
    *back_slot  = merge_one( a, br, bf, 0, end, revcom, score, 'b' );
    return 2;
  }
  *front_slot = merge_one( a, ref_gapped, read_gapped, start, end, revcom, score, 'a' );
  return 1;
}

static int slot_len( const orc_slot* s ) {            /* asp_len, fsdb.c:518-530 */
  int n = s->end - s->start + 1, t = n;
  for ( int i = 0; i < n; i++ ) if ( s->ins[i] ) t += (int)strlen( s->ins[i] );
  return t;
}

void orc_asm_pop_smp( orc_asm* a, long long n_reads, const int* front, const int* back ) {   /* fsdb.c:542-619 */
  for ( long long r = 0; r < n_reads; r++ ) {
    if ( front[r] < 0 ) continue;
    orc_slot* f = a->s[front[r]];
    orc_slot* b = back[r] >= 0 ? a->s[back[r]] : NULL;
    int fl = slot_len( f ), bl = b ? slot_len( b ) : 0, act = 0;
    for ( int seg = 0; seg < 2; seg++ ) {
      orc_slot* s = seg ? b : f;
      if ( !s ) break;
      int n = s->end - s->start + 1, i;
      for ( i = 0; i < n; i++ ) {
        if ( s->ins[i] ) act += (int)strlen( s->ins[i] );
        /* NB the back segment adds the front length on top of a counter that
           was never reset (fsdb.c:591-596) -- reproduced on purpose */
        int dfront = seg ? fl + act : act;
        int dback = fl + bl - act - 1;
        if ( dfront <= ORC_PSSM_DEPTH ) s->smp[i] = (char)( 'A' + dfront );
        else if ( dback < ORC_PSSM_DEPTH ) s->smp[i] = (char)( 'A' + 2*ORC_PSSM_DEPTH - dback );
        else s->smp[i] = (char)( 'A' + ORC_PSSM_DEPTH );
        if ( s->seq[i] != '-' ) act++;
      }
      s->smp[i] = '\0';
    }
  }
}

void orc_score_cut( long long n, const int* seq_len, const int* score,
                    const unsigned char* unique_best, double* slope, double* intercept ) {  /* fsdb.c:269-383 */
  double xbar = 0, ybar = 0, ssxy = 0, ssxx = 0, maxd = 0;
  long long j = 0;
  for ( long long i = 0; i < n; i++ )
    if ( ( !unique_best || unique_best[i] ) && score[i] >= 2000 ) { xbar += seq_len[i]; ybar += score[i]; j++; }
  xbar /= j; ybar /= j;
  for ( long long i = 0; i < n; i++ )
    if ( ( !unique_best || unique_best[i] ) && score[i] >= 2000 ) {
      ssxy += ( seq_len[i] - xbar ) * ( score[i] - ybar );
      ssxx += ( seq_len[i] - xbar ) * ( seq_len[i] - xbar );
    }
  double bf = ssxy / ssxx, ib = ybar - bf * xbar;
  for ( long long i = 0; i < n; i++ )
    if ( ( !unique_best || unique_best[i] ) && score[i] >= 2000 ) {
      double d = ( score[i] - ( ( bf * seq_len[i] ) + ib ) ) / seq_len[i];
      if ( d > maxd ) maxd = d;
    }
  *intercept = ib;
  if ( ( bf - maxd ) > 0 ) *slope = bf - ( maxd * 2.0 );
  else *slope = (double)( bf * ( 80 / 100.0 ) );       /* SCORE_CUTOFF_BUFFER */
}

void orc_asm_cull( orc_asm* a, long long n_reads, const int* front, const int* back,
                   const int* seq_len, const int* score,
                   int hard_cut, int score_cut_set, double s, double n ) {   /* mia.c:418-506 */
  orc_asm_cull_u( a, n_reads, front, back, seq_len, score, NULL, hard_cut, score_cut_set, s, n );
}
void orc_asm_cull_u( orc_asm* a, long long n_reads, const int* front, const int* back,
                     const int* seq_len, const int* score, const unsigned char* unique_best,
                     int hard_cut, int score_cut_set, double s, double n ) {
  orc_asm_cull_d( a, n_reads, front, back, seq_len, score, unique_best, NULL, hard_cut, score_cut_set, s, n );
}

/* find_alignable_len, mia.c:69-91: the read's length minus the N positions of ref[as, min(ae, wrap_len)), at least
 * MIN_ALIGNABLE_LEN = 15 (params.h:38) */
int orc_alignable_len( const char* ref_wrapped, int wrap_len, int seq_len, int as, int ae ) {
  int n = seq_len;
  long long end = ae;
  if ( end > wrap_len ) end = wrap_len;
  for ( long long i = as; i < end; i++ ) if ( ref_wrapped[i] == 'N' ) n--;
  return n < 15 ? 15 : n;
}

/* -D: the threshold of a read takes alignable_len[r] (mia.c:460-463); the regression still runs over seq_len */
void orc_asm_cull_d( orc_asm* a, long long n_reads, const int* front, const int* back,
                     const int* seq_len, const int* score, const unsigned char* unique_best,
                     const int* alignable_len, int hard_cut, int score_cut_set, double s, double n ) {
  double slope, intercept;
  if ( score_cut_set ) { slope = s; intercept = n; }
  else orc_score_cut( n_reads, seq_len, score, unique_best, &slope, &intercept );
  if ( slope <= 0 ) slope = 100.0;
  a->ne = 0;
  for ( long long r = 0; r < n_reads; r++ ) {
    if ( front[r] < 0 ) continue;
    if ( unique_best && !unique_best[r] ) continue;   /* mia.c:466: not in the culled list at all */
    double min_score = hard_cut > 0 ? (double)hard_cut : (double)( intercept + ( slope * ( alignable_len ? alignable_len[r] : seq_len[r] ) ) );
    int drop = ( score[r] < min_score );
    int ids[2] = { front[r], back[r] };
    for ( int q = 0; q < 2; q++ ) {
      if ( ids[q] < 0 ) continue;
      if ( a->ne == a->ecap ) { a->ecap = a->ecap ? 2*a->ecap : 1024; a->e = (int*)realloc( a->e, sizeof(int) * a->ecap ); }
      a->e[a->ne++] = ids[q];
      if ( drop ) a->s[ids[q]]->dropped = 1;          /* only ever set: H10 */
    }
  }
  for ( int i = 0; i < a->seq_len; i++ ) {            /* mia.c:486-504 */
    if ( a->gaps[i] <= 0 ) continue;
    int g = 0;
    for ( int j = 0; j < a->ne; j++ ) {
      const orc_slot* q = a->s[a->e[j]];
      if ( q->start < i && q->end >= i && q->ins[i - q->start] ) {
        int l = (int)strlen( q->ins[i - q->start] );
        if ( l > g ) g = l;
      }
    }
    a->gaps[i] = g;
  }
}

typedef struct { int As, Cs, Gs, Ts, gaps, cov, sA, sC, sG, sT; } orc_bc;

static void add_base( orc_bc* b, char ch, const int* sm, int code ) {   /* map_align.c:229-263 */
  switch ( ch ) { case 'A': b->As++; break; case 'C': b->Cs++; break; case 'G': b->Gs++; break;
                  case 'T': b->Ts++; break; case '-': b->gaps++; break; }
  b->cov++;
  if ( ch == '-' ) return;
  int j = orc_base_code( ch ), d = code - 'A';
  b->sA += SM(sm,d,0,j); b->sC += SM(sm,d,1,j); b->sG += SM(sm,d,2,j); b->sT += SM(sm,d,3,j);
}

static char call_base( const orc_bc* b, int cons_code ) {               /* map_align.c:294-391 */
  if ( b->cov == 0 ) return 'N';
  if ( (double)b->gaps / (double)b->cov >= 0.5 ) return '-';
  int top = b->sA, second = INT_MIN; char base = 'A';
  const int sc[3] = { b->sC, b->sG, b->sT }; const char nm[3] = { 'C', 'G', 'T' };
  for ( int i = 0; i < 3; i++ ) {
    if ( sc[i] >= top ) { second = top; top = sc[i]; base = nm[i]; }   /* >= : later base wins ties */
    else if ( i == 0 || sc[i] >= second ) second = sc[i];
  }
  if ( cons_code == 2 ) return ( top >= 0 || ( top - 2400 ) > second ) ? base : 'N';
  return top >= -399 ? base : 'N';
}

int orc_find_consensus( const int* in, int cons_code ) {
  orc_bc b = { in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], in[8], in[9] };
  return call_base( &b, cons_code );
}

/* mia.c:515-603 over the culled entry list (duplicates included) */
int orc_asm_consensus( const orc_asm* a, const int* smf, const int* smr, int cons_code,
                       char* cons, int* counts ) {
  int cp = 0;
  for ( int p = 0; p < a->seq_len; p++ ) {
    int g = a->gaps[p];
    if ( g > 0 && p > 0 ) {                           /* find_ins_cons, map_align.c:444-510 */
      for ( int j = 0; j < g; j++ ) {
        orc_bc b; memset( &b, 0, sizeof(b) );
        for ( int i = 0; i < a->ne; i++ ) {
          const orc_slot* s = a->s[a->e[i]];
          if ( !( s->start < p && s->end >= p ) ) continue;            /* dropped NOT checked */
          const char* ins = s->ins[p - s->start];
          char ch = ( ins && j < (int)strlen( ins ) ) ? ins[j] : '-';
          add_base( &b, ch, s->revcom ? smr : smf, s->smp[p - s->start] );
        }
        char cb = call_base( &b, cons_code );
        if ( cb != '-' ) cons[cp++] = cb;
      }
    }
    orc_bc b; memset( &b, 0, sizeof(b) );
    for ( int i = 0; i < a->ne; i++ ) {
      const orc_slot* s = a->s[a->e[i]];
      if ( s->start <= p && s->end >= p && !s->dropped )
        add_base( &b, s->seq[p - s->start], s->revcom ? smr : smf, s->smp[p - s->start] );
    }
    if ( counts ) memcpy( counts + 10*p, &b, sizeof(b) );
    char cb = call_base( &b, cons_code );
    if ( cb != '-' ) cons[cp++] = cb;
  }
  cons[cp] = '\0';
  return cp;
}

int orc_asm_num_slots( const orc_asm* a ) { return a->n; }
int orc_asm_num_entries( const orc_asm* a ) { return a->ne; }
int orc_asm_entry( const orc_asm* a, int i ) { return a->e[i]; }
void orc_asm_gaps( const orc_asm* a, int* out ) { memcpy( out, a->gaps, sizeof(int) * ( a->wrap_len + 1 ) ); }

void orc_asm_slot( const orc_asm* a, int i, int* out7, char* seq, char* smp, char* ins ) {
  const orc_slot* s = a->s[i];
  int n = 0; char* p = ins;
  out7[0] = s->start; out7[1] = s->end; out7[2] = s->score; out7[3] = s->revcom;
  out7[4] = s->dropped; out7[5] = s->segment;
  strcpy( seq, s->seq ); strcpy( smp, s->smp );
  for ( int j = 0; j < s->end - s->start + 1; j++ )
    if ( s->ins[j] ) { p += sprintf( p, "%d:%s;", j, s->ins[j] ); n++; }
  *p = '\0';
  out7[6] = n;
}

/* ---- f1: repeat filter (fsdb.c:13-88 fs_comp, 90-180 fs_comp_qscore, 440-508 set_uniq_in_fsdb) */
typedef struct { const unsigned char* rc; const int* as; const int* ae; const int* k4; } orc_rf;
static int orc_rf_comp( const orc_rf* d, long long a, long long b ) {
  const int ra = d->rc[a] != 0, rb = d->rc[b] != 0;
  if ( ra && !rb ) return -1;                        /* reverse strand first */
  if ( !ra && rb ) return 1;
  if ( !ra ) {                                       /* forward: as ascending, ae descending, key4 descending */
    if ( d->as[a] != d->as[b] ) return d->as[a] < d->as[b] ? -1 : 1;
    if ( d->ae[a] != d->ae[b] ) return d->ae[a] < d->ae[b] ? 1 : -1;
  } else {                                           /* reverse: ae descending, as ascending, key4 descending */
    if ( d->ae[a] != d->ae[b] ) return d->ae[a] < d->ae[b] ? 1 : -1;
    if ( d->as[a] != d->as[b] ) return d->as[a] < d->as[b] ? -1 : 1;
  }
  if ( d->k4[a] != d->k4[b] ) return d->k4[a] < d->k4[b] ? 1 : -1;
  return 0;
}
void orc_repeat_filter( long long n, const unsigned char* rc, const int* as, const int* ae, const int* key4,
                        const unsigned char* trimmed, int just_outer_coords, int tolerance,
                        long long* order, unsigned char* unique ) {
  orc_rf d;
  long long i, w, *tmp;
  int curr_rc, curr_as, curr_ae;
  if ( n <= 0 ) return;
  d.rc = rc; d.as = as; d.ae = ae; d.k4 = key4;
  tmp = (long long*)malloc( (size_t)n * sizeof(long long) );
  for ( i = 0; i < n; i++ ) order[i] = i;
  for ( w = 1; w < n; w *= 2 ) {                     /* bottom-up merge sort: stable */
    long long lo;
    for ( lo = 0; lo < n; lo += 2 * w ) {
      long long mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
      long long a = lo, b = mid, k = lo;
      while ( a < mid && b < hi ) tmp[k++] = orc_rf_comp( &d, order[b], order[a] ) < 0 ? order[b++] : order[a++];
      while ( a < mid ) tmp[k++] = order[a++];
      while ( b < hi ) tmp[k++] = order[b++];
    }
    memcpy( order, tmp, (size_t)n * sizeof(long long) );
  }
  free( tmp );
  /* set_uniq_in_fsdb */
  curr_rc = rc[order[0]] != 0; curr_as = as[order[0]]; curr_ae = ae[order[0]];
  unique[order[0]] = 1;
  for ( i = 1; i < n; i++ ) {
    const long long f = order[i];
    const int frc = rc[f] != 0;
    if ( frc == curr_rc && abs( as[f] - curr_as ) <= tolerance && abs( ae[f] - curr_ae ) <= tolerance ) {
      unique[f] = 0;                                 /* curr stays */
      continue;
    }
    if ( just_outer_coords ) unique[f] = 1;
    else if ( !frc ) unique[f] = ( as[f] == curr_as ) ? ( trimmed && trimmed[f] ? 1 : 0 ) : 1;
    else unique[f] = ( ae[f] == curr_ae ) ? ( trimmed && trimmed[f] ? 1 : 0 ) : 1;
    curr_rc = frc; curr_as = as[f]; curr_ae = ae[f];
  }
}

/* ---- f4: trim_frag (mia.c:1318-1368) */
int orc_trim( const char* read, int read_len, const char* adapter, int adapter_len, int* out6 ) {
  int sm[775], o5[5];
  char rg[2*ORC_MAX_READ + 2], fg[2*ORC_MAX_READ + 2];
  if ( read_len <= 0 || adapter_len <= 0 || adapter_len > ORC_MAX_READ ) return 0;
  int* S = (int*)malloc( sizeof(int) * (size_t)read_len * adapter_len );
  int* T = (int*)malloc( sizeof(int) * (size_t)read_len * adapter_len );
  orc_flat_pssm( sm );
  if ( !orc_align( read, read_len, adapter, adapter_len, NULL, sm, 1, o5, rg, fg, S, T ) ) { free( S ); free( T ); return 0; }
  /* the last column, rows in order, strict '>' (mia.c:1345-1352) */
  int col = read_len - 1, best = INT_MIN, aer = 0;
  for ( int r = 0; r < adapter_len; r++ )
    if ( S[(size_t)r * read_len + col] > best ) { best = S[(size_t)r * read_len + col]; aer = r; }
  /* find_align_begin (mia.c:612-637) */
  int row = aer;
  for (;;) {
    int t = T[(size_t)row * read_len + col];
    if ( t == col || t == -row ) break;
    if ( t == 0 ) { row--; col--; }
    else if ( t < 0 ) { row = -t; col--; }
    else { col = t; row--; }
  }
  int trimmed = ( best >= 1000 ) || ( best >= ( aer - row + 1 ) * 200 );   /* TRIM_SCORE_CUT, FLAT_MATCH: params.h:28,32 */
  out6[0] = trimmed; out6[1] = trimmed ? col - 1 : 0; out6[2] = best; out6[3] = row; out6[4] = col; out6[5] = aer;
  free( S ); free( T );
  return 1;
}
