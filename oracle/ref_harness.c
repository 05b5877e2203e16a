/* ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * ctypes-friendly entry points that drive the UNMODIFIED reference
 * (compiled from /root/reference/src where it lies; see oracle/Makefile)
 * so that tests can pin oracle/mia_oracle.c and the CUDA path against the
 * reference's own behaviour.  Nothing in here restates an algorithm: every
 * result comes out of a reference function.  The call sequences mirror
 *   - ccheck.cc:571-603 / mia_main.c:217-235   (refh_align)
 *   - mia_main.c:659-672, 781                  (refh_kmer_*)
 *   - mia_main.c:618-976                       (refh_sess_*)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include "mia.h"
#include <time.h>

/* lives in mia_main.c, compiled with -Dmain=mia_cli_main */
void reiterate_assembly( char* new_ref_seq, int iter_num, MapAlignmentP maln,
                         FSDB fsdb, AlignmentP a, PWAlnFragP front_pwaln,
                         PWAlnFragP back_pwaln, PSSMP ancsubmat,
                         PSSMP rcancsubmat );

/* ------------------------------------------------------------------ PSSM */
static PSSMP pssm_from_flat( const int* sm775 ) {
  PSSMP p = (PSSMP)malloc( sizeof(PSSM) );
  memcpy( p->sm, sm775, sizeof(p->sm) );
  p->depth = PSSM_DEPTH;
  return p;
}

void refh_read_pssm( const char* fn, int* out775 ) {
  PSSMP p = read_pssm( fn );
  memcpy( out775, p->sm, sizeof(p->sm) );
  free( p );
}

void refh_flat_pssm( int* out775 ) {
  PSSMP p = init_flatsubmat();
  memcpy( out775, p->sm, sizeof(p->sm) );
  free( p );
}

void refh_revcom_pssm( const int* in775, int* out775 ) {
  PSSMP p = pssm_from_flat( in775 );
  PSSMP r = revcom_submat( p );
  memcpy( out775, r->sm, sizeof(r->sm) );
  free( p );
  free( r );
}

int refh_find_sm_depth( int row, int len ) { return find_sm_depth( row, len ); }

/* ------------------------------------------------------------- alignment */
static AlignmentP g_al = NULL;
static int g_al_cols = 0;
static int g_hp = 0, g_al_hp = 0;
/* mia -h (hp_special, mia_main.c:424, 497): the alignments refh_align and refh_sess_new make from now on carry it */
void refh_set_hp( int on ) { g_hp = on ? 1 : 0; }

/* out5 = score, abr, abc, aer, aec.  score_mat / trace_mat (nullable) get the
   full len2 x len1 matrices, row-major.  Returns 1, or 0 on alloc failure. */
int refh_align( const char* seq1, int len1, const char* seq2, int len2,
                const unsigned char* mask, const int* sm775, int sg5,
                int* out5, char* ref_gapped, char* read_gapped,
                int* score_mat, int* trace_mat ) {
  PWAlnFrag pw;
  PSSMP sm = pssm_from_flat( sm775 );
  int r, c;
  if ( g_al == NULL || g_al_cols < len1 + 1 || g_al_hp != g_hp ) {
    if ( g_al ) free_alignment( g_al );
    g_al_cols = len1 + 2 * INIT_ALN_SEQ_LEN;
    g_al = init_alignment( INIT_ALN_SEQ_LEN, g_al_cols, 0, g_hp );
    g_al_hp = g_hp;
    if ( g_al == NULL ) return 0;
  }
  g_al->seq1 = seq1;
  g_al->len1 = len1;
  g_al->seq2 = seq2;
  g_al->len2 = len2;
  g_al->submat = sm;
  g_al->sg5 = sg5;
  g_al->sg3 = sg5;
  if ( mask ) memcpy( g_al->align_mask, mask, len1 );
  else        memset( g_al->align_mask, 1, len1 );
  pop_s1c_in_a( g_al );
  pop_s2c_in_a( g_al );
  if ( g_al->hp ) {                                   /* mia_main.c:221-224 */
    pop_hpl_and_hps( g_al->seq2, g_al->len2, g_al->hprl, g_al->hprs );
    pop_hpl_and_hps( g_al->seq1, g_al->len1, g_al->hpcl, g_al->hpcs );
  }
  dyn_prog( g_al );
  out5[0] = max_sg_score( g_al );
  find_align_begin( g_al );
  populate_pwaln_to_begin( g_al, &pw );
  out5[1] = g_al->abr; out5[2] = g_al->abc;
  out5[3] = g_al->aer; out5[4] = g_al->aec;
  strcpy( ref_gapped, pw.ref_seq );
  strcpy( read_gapped, pw.frag_seq );
  if ( score_mat || trace_mat ) {
    for ( r = 0; r < len2; r++ )
      for ( c = 0; c < len1; c++ ) {
        if ( score_mat ) score_mat[r*len1 + c] = g_al->m->mat[r][c].score;
        if ( trace_mat ) trace_mat[r*len1 + c] = g_al->m->mat[r][c].trace;
      }
  }
  free( sm );
  return 1;
}

/* Time `reps` passes of the reference's per-read hot sequence over a batch
   (pop_s1c -> pop_s2c -> dyn_prog -> max_sg_score -> find_align_begin ->
   populate_pwaln_to_begin); reads concatenated, windows given per read.
   Returns seconds of wall time; checksum defeats dead-code elimination. */
double refh_time_realign( const char* ref, int n, const char* reads,
                          const long long* off, const int* win_start,
                          const int* win_len, const int* rc,
                          const int* smf775, const int* smr775,
                          long long* cells_out, long long* checksum ) {
  struct timespec t0, t1;
  PSSMP f = pssm_from_flat( smf775 ), r = pssm_from_flat( smr775 );
  PWAlnFrag pw;
  char rd[INIT_ALN_SEQ_LEN + 1];
  int i, maxw = 0;
  long long cells = 0, ck = 0;
  for ( i = 0; i < n; i++ ) if ( win_len[i] > maxw ) maxw = win_len[i];
  if ( g_al == NULL || g_al_cols < maxw + 1 ) {
    if ( g_al ) free_alignment( g_al );
    g_al_cols = maxw + 2 * INIT_ALN_SEQ_LEN;
    g_al = init_alignment( INIT_ALN_SEQ_LEN, g_al_cols, 0, 0 );
  }
  memset( g_al->align_mask, 1, g_al_cols );
  g_al->sg5 = 1; g_al->sg3 = 1;
  clock_gettime( CLOCK_MONOTONIC, &t0 );
  for ( i = 0; i < n; i++ ) {
    int len2 = (int)(off[i+1] - off[i]);
    memcpy( rd, reads + off[i], len2 );
    rd[len2] = '\0';
    g_al->submat = rc[i] ? r : f;
    g_al->seq2 = rd; g_al->len2 = len2;
    pop_s2c_in_a( g_al );
    g_al->seq1 = ref + win_start[i]; g_al->len1 = win_len[i];
    pop_s1c_in_a( g_al );
    dyn_prog( g_al );
    ck += max_sg_score( g_al );
    find_align_begin( g_al );
    populate_pwaln_to_begin( g_al, &pw );
    ck += g_al->abc + pw.ref_seq[0];
    cells += (long long)len2 * win_len[i];
  }
  clock_gettime( CLOCK_MONOTONIC, &t1 );
  *cells_out = cells; *checksum = ck;
  free( f ); free( r );
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ----------------------------------------------------------------- k-mer */
void* refh_kmer_new( const char* seq, long long len, int k, int soft_mask ) {
  KPL* kpa = init_kpa( k );
  populate_kpa( kpa, seq, (size_t)len, k, soft_mask );
  return kpa;
}

/* number of stored positions for k-mer index inx; positions copied to out */
int refh_kmer_lookup( void* kpa_, long long inx, unsigned int* out ) {
  KPL* kpa = (KPL*)kpa_;
  if ( kpa[inx] == NULL ) return 0;
  memcpy( out, kpa[inx]->positions, kpa[inx]->num_pos * sizeof(unsigned int) );
  return (int)kpa[inx]->num_pos;
}

void refh_kmer_free( void* kpa_, int k ) {
  KPL* kpa = (KPL*)kpa_;
  size_t i, n = (size_t)1 << (2*k);
  for ( i = 0; i < n; i++ ) free( kpa[i] );
  free( kpa );
}

/* Runs new_kmer_filter (kmer.c:239) with masks of length len1 (both strands).
   Returns its return value (total hits; 0 => do not align). */
unsigned int refh_kmer_filter( void* fkpa, void* rkpa, int k,
                               const char* read, int read_len, int len1,
                               unsigned char* mask_f, unsigned char* mask_r ) {
  FragSeq fs;
  Alignment fwa, rca;
  unsigned int hits;
  memset( &fs, 0, sizeof(fs) );
  memcpy( fs.seq, read, read_len );
  fs.seq[read_len] = '\0';
  fs.seq_len = read_len;
  fs.trimmed = 0;
  fwa.align_mask = mask_f; fwa.len1 = len1;
  rca.align_mask = mask_r; rca.len1 = len1;
  hits = new_kmer_filter( &fs, (KPL*)fkpa, (KPL*)rkpa, k, &fwa, &rca );
  return hits;
}

/* --------------------------------------------------------------- session */
typedef struct {
  MapAlignmentP maln, culled;
  FSDB fsdb;
  AlignmentP fw, rc;
  PSSMP anc, rcanc;
  KPL *fkpa, *rkpa;
  int k;
  PWAlnFragP front, back;
  FragSeqP fs;
  int iter;
  char *last_cons, *cons;
  int hard_cut, score_cut_set;
  double slope, intercept;
  int repeat_filt, just_outer_coords;      /* -u (1) or -U (2), -A */
  int next_qual_sum;                       /* FragSeq.qual_sum of the next refh_sess_pass1 read (what read_fastq computes) */
} Sess;

/* mia_main.c:618-757 with the getopt results passed in */
void* refh_sess_new( const char* ref_fasta, int circular, int k, int soft_mask,
                     const int* sm775, int distant_ref, int cons_code ) {
  Sess* s = (Sess*)calloc( 1, sizeof(Sess) );
  int i;
  s->anc = pssm_from_flat( sm775 );
  s->rcanc = revcom_submat( s->anc );
  s->maln = init_map_alignment();
  s->maln->cons_code = cons_code;
  s->maln->distant_ref = distant_ref;
  s->fsdb = init_FSDB();
  if ( read_fasta_ref( s->maln->ref, ref_fasta ) != 1 ) return NULL;
  if ( circular ) add_ref_wrap( s->maln->ref );
  else s->maln->ref->wrap_seq_len = s->maln->ref->seq_len;
  s->maln->ref->gaps = (int*)malloc( (s->maln->ref->wrap_seq_len+1) * sizeof(int) );
  for ( i = 0; i <= s->maln->ref->wrap_seq_len; i++ ) s->maln->ref->gaps[i] = 0;
  s->k = k;
  if ( k > 0 ) {
    s->fkpa = init_kpa( k );
    s->rkpa = init_kpa( k );
    populate_kpa( s->fkpa, s->maln->ref->seq, s->maln->ref->wrap_seq_len, k, soft_mask );
    populate_kpa( s->rkpa, s->maln->ref->rcseq, s->maln->ref->wrap_seq_len, k, soft_mask );
  }
  make_ref_upper( s->maln->ref );
  s->fs = (FragSeqP)calloc( 1, sizeof(FragSeq) );
  s->fw = init_alignment( INIT_ALN_SEQ_LEN, s->maln->ref->wrap_seq_len + 2*INIT_ALN_SEQ_LEN, 0, g_hp );
  s->rc = init_alignment( INIT_ALN_SEQ_LEN, s->maln->ref->wrap_seq_len + 2*INIT_ALN_SEQ_LEN, 1, g_hp );
  s->fw->seq1 = s->maln->ref->seq;
  s->rc->seq1 = s->maln->ref->rcseq;
  s->fw->len1 = circular ? s->maln->ref->wrap_seq_len : s->maln->ref->seq_len;
  s->rc->len1 = s->fw->len1;
  pop_s1c_in_a( s->fw );
  pop_s1c_in_a( s->rc );
  if ( g_hp ) {                                       /* mia_main.c:735-739 */
    pop_hpl_and_hps( s->fw->seq1, s->fw->len1, s->fw->hpcl, s->fw->hpcs );
    pop_hpl_and_hps( s->rc->seq1, s->rc->len1, s->rc->hpcl, s->rc->hpcs );
  }
  s->front = (PWAlnFragP)calloc( 1, sizeof(PWAlnFrag) );
  s->back  = (PWAlnFragP)calloc( 1, sizeof(PWAlnFrag) );
  s->slope = DEF_S; s->intercept = DEF_N;
  s->just_outer_coords = 1;                /* mia_main.c:420 */
  return s;
}

void refh_sess_set_cut( void* s_, int hard_cut, int score_cut_set, double slope, double intercept ) {
  Sess* s = (Sess*)s_;
  s->hard_cut = hard_cut; s->score_cut_set = score_cut_set;
  s->slope = slope; s->intercept = intercept;
}

/* -u (repeat_filt) and -A (just_outer_coords = 0): mia_main.c:500-505 */
void refh_sess_set_repeat( void* s_, int repeat_filt, int just_outer_coords ) {
  Sess* s = (Sess*)s_;
  s->repeat_filt = repeat_filt; s->just_outer_coords = just_outer_coords;
}
static void sess_repeat_filter( Sess* s ) {            /* mia_main.c:827-834, 883-886, 938-941 */
  if ( s->repeat_filt && s->fsdb->num_fss > 0 ) {
    if ( s->repeat_filt == 2 ) sort_fsdb_qscore( s->fsdb );       /* -U: mia_main.c:836-844, 887-890, 942-945 */
    else sort_fsdb( s->fsdb );
    set_uniq_in_fsdb( s->fsdb, s->just_outer_coords, 0 );
  }
}
void refh_sess_set_next_qual_sum( void* s_, int q ) { ((Sess*)s_)->next_qual_sum = q; }
void refh_sess_fs_id( void* s_, long long i, char* id ) { strcpy( id, ((Sess*)s_)->fsdb->fss[i]->id ); }

/* One read through mia_main.c:759-797 (no -T, no -I).
   out[0]=kmer hits (return of new_kmer_filter)  out[1]=added to fsdb (0/1)
   out[2]=score out[3]=rc out[4]=as out[5]=ae out[6]=strand_known
   out[7]=fw best_score out[8]=rc best_score
   out[9]=front start out[10]=front end out[11]=split (0/1)
   out[12]=back start out[13]=back end
   out[14]=best abr out[15]=abc out[16]=aer out[17]=aec
   strings: front ref/frag, back ref/frag (each >= 513 bytes). */
int refh_sess_pass1( void* s_, const char* id, const char* seq, int* out,
                     char* f_ref, char* f_frag, char* b_ref, char* b_frag ) {
  Sess* s = (Sess*)s_;
  FragSeqP fs = s->fs;
  size_t before = s->fsdb->num_fss;
  int ok = 1, i;
  AlignmentP best;
  for ( i = 0; i < 18; i++ ) out[i] = 0;
  f_ref[0] = f_frag[0] = b_ref[0] = b_frag[0] = '\0';
  strncpy( fs->id, id, MAX_ID_LEN ); fs->id[MAX_ID_LEN] = '\0';
  fs->desc[0] = '\0';
  strncpy( fs->seq, seq, INIT_ALN_SEQ_LEN ); fs->seq[INIT_ALN_SEQ_LEN] = '\0';
  fs->seq_len = strlen( fs->seq );
  fs->qual[0] = '\0';
  fs->qual_sum = s->next_qual_sum;
  fs->trimmed = 0;
  out[0] = new_kmer_filter( fs, s->fkpa, s->rkpa, s->k > 0 ? s->k : -1, s->fw, s->rc );
  if ( out[0] ) {
    s->fw->submat = s->anc;
    s->rc->submat = s->anc;
    ok = sg_align( s->maln, fs, s->fsdb, s->fw, s->rc, s->front, s->back );
    out[1] = ( s->fsdb->num_fss > before );
    out[2] = fs->score; out[3] = fs->rc; out[4] = fs->as; out[5] = fs->ae;
    out[6] = out[1] ? fs->strand_known : 0;
    out[7] = s->fw->best_score; out[8] = s->rc->best_score;
    best = ( s->fw->best_score > s->rc->best_score ) ? s->fw : s->rc;
    out[14] = best->abr; out[15] = best->abc; out[16] = best->aer; out[17] = best->aec;
    out[9] = s->front->start; out[10] = s->front->end;
    strcpy( f_ref, s->front->ref_seq ); strcpy( f_frag, s->front->frag_seq );
    if ( out[1] && s->front->segment == 'f' ) {
      out[11] = 1;
      out[12] = s->back->start; out[13] = s->back->end;
      strcpy( b_ref, s->back->ref_seq ); strcpy( b_frag, s->back->frag_seq );
    }
  }
  return ok;
}

/* masks as left by the last refh_sess_pass1 call */
void refh_sess_masks( void* s_, unsigned char* mf, unsigned char* mr ) {
  Sess* s = (Sess*)s_;
  memcpy( mf, s->fw->align_mask, s->fw->len1 );
  memcpy( mr, s->rc->align_mask, s->rc->len1 );
}

/* mia_main.c:812-876 (no -u/-U/-C) */
void refh_sess_end_pass1( void* s_ ) {
  Sess* s = (Sess*)s_;
  pop_smp_from_FSDB( s->fsdb, PSSM_DEPTH );
  s->iter = 1;
  s->culled = init_culled_map_alignment( s->maln );
  sess_repeat_filter( s );
  cull_maln_from_fsdb( s->culled, s->fsdb, s->hard_cut, s->score_cut_set, s->slope, s->intercept );
  s->culled->fpsm = s->anc;
  s->culled->rpsm = s->rcanc;
  sort_aln_frags( s->culled );
  s->fw->submat = s->anc;
  s->fw->sg5 = 1;
  s->fw->sg3 = 1;
  s->last_cons = (char*)malloc( s->maln->ref->seq_len + 1 );
  strncpy( s->last_cons, s->maln->ref->seq, s->maln->ref->seq_len );
  s->last_cons[s->maln->ref->seq_len] = '\0';
  memset( s->fw->align_mask, 1, s->fw->len1 );
  clean_FSDB( s->fsdb );
  s->cons = NULL;
}

/* One round of mia_main.c:878-900 (first call) or 918-955 (later calls),
   followed by consensus_assembly_string (913 / 963).  If sort==0 the
   sort_aln_frags step is skipped so that culled AlnSeq order == FSDB order
   (the consensus does not depend on that order).  Returns the consensus;
   *converged = strcmp(new, last)==0. */
const char* refh_sess_iterate( void* s_, int sort, int* converged ) {
  Sess* s = (Sess*)s_;
  char* ref_for_round;
  if ( s->cons == NULL ) {
    ref_for_round = s->last_cons;            /* "iteration 1" */
  } else {
    s->iter++;
    free( s->last_cons );
    s->last_cons = s->cons;
    ref_for_round = s->cons;
  }
  reiterate_assembly( ref_for_round, s->iter, s->maln, s->fsdb, s->fw,
                      s->front, s->back, s->anc, s->rcanc );
  pop_smp_from_FSDB( s->fsdb, PSSM_DEPTH );
  sess_repeat_filter( s );
  cull_maln_from_fsdb( s->culled, s->fsdb, s->hard_cut, s->score_cut_set, s->slope, s->intercept );
  s->culled->fpsm = s->anc;
  s->culled->rpsm = s->rcanc;
  if ( sort ) sort_aln_frags( s->culled );
  s->cons = consensus_assembly_string( s->culled );
  *converged = ( strcmp( s->cons, s->last_cons ) == 0 );
  return s->cons;
}

int refh_sess_iter_num( void* s_ ) { return ((Sess*)s_)->iter; }
int refh_sess_write_ma( void* s_, char* fn ) { return write_ma( fn, ((Sess*)s_)->culled ); }
int refh_sess_ref_len( void* s_ ) { return ((Sess*)s_)->maln->ref->seq_len; }
int refh_sess_wrap_len( void* s_ ) { return ((Sess*)s_)->maln->ref->wrap_seq_len; }
const char* refh_sess_ref_seq( void* s_ ) { return ((Sess*)s_)->maln->ref->seq; }
void refh_sess_gaps( void* s_, int* out ) {
  Sess* s = (Sess*)s_;
  memcpy( out, s->maln->ref->gaps, (s->maln->ref->wrap_seq_len+1) * sizeof(int) );
}
long long refh_sess_num_fs( void* s_ ) { return (long long)((Sess*)s_)->fsdb->num_fss; }

/* out: seq_len, score, rc, as, ae, strand_known, unique_best, has_back */
void refh_sess_fs( void* s_, long long i, int* out, char* seq ) {
  FragSeqP fs = ((Sess*)s_)->fsdb->fss[i];
  out[0] = fs->seq_len; out[1] = fs->score; out[2] = fs->rc; out[3] = fs->as;
  out[4] = fs->ae; out[5] = fs->strand_known; out[6] = fs->unique_best;
  out[7] = ( fs->back_asp != NULL );
  if ( seq ) strcpy( seq, fs->seq );
}

int refh_sess_num_aln( void* s_ ) { return ((Sess*)s_)->culled->num_aln_seqs; }

/* out: start, end, score, revcom, dropped, segment(char), n_ins
   seq/smp copied; ins: for each column with an insert "pos:SEQ;" appended */
void refh_sess_aln( void* s_, int i, int* out, char* id, char* seq, char* smp, char* ins ) {
  AlnSeqP a = ((Sess*)s_)->culled->AlnSeqArray[i];
  int j, n = 0, len = a->end - a->start + 1;
  char* p = ins;
  out[0] = a->start; out[1] = a->end; out[2] = a->score; out[3] = a->revcom ? 1 : 0;
  out[4] = a->dropped ? 1 : 0; out[5] = a->segment;
  strcpy( id, a->id ); strcpy( seq, a->seq ); strcpy( smp, a->smp );
  for ( j = 0; j < len; j++ )
    if ( a->ins[j] != NULL ) { p += sprintf( p, "%d:%s;", j, a->ins[j] ); n++; }
  *p = '\0';
  out[6] = n;
}

/* BaseCounts of base column `pos` exactly as consensus_assembly_string
   (mia.c:575-598) accumulates them; out10 = As,Cs,Gs,Ts,gaps,cov,sA,sC,sG,sT.
   Returns the called character (may be '-'). */
int refh_sess_column( void* s_, int pos, int* out10 ) {
  Sess* s = (Sess*)s_;
  MapAlignmentP m = s->culled;
  BaseCounts b;
  int j;
  reset_base_counts( &b );
  for ( j = 0; j < m->num_aln_seqs; j++ ) {
    AlnSeqP a = m->AlnSeqArray[j];
    if ( a->start <= pos && a->end >= pos && !a->dropped )
      add_base( a->seq[pos - a->start], &b, a->revcom ? m->rpsm : m->fpsm, a->smp[pos - a->start] );
  }
  out10[0] = b.As; out10[1] = b.Cs; out10[2] = b.Gs; out10[3] = b.Ts; out10[4] = b.gaps;
  out10[5] = b.cov; out10[6] = b.scoreA; out10[7] = b.scoreC; out10[8] = b.scoreG; out10[9] = b.scoreT;
  return find_consensus( &b, m->cons_code );
}

/* consensus call on a hand-filled BaseCounts (map_align.c:294) */
int refh_find_consensus( const int* in10, int cons_code ) {
  BaseCounts b;
  b.As = in10[0]; b.Cs = in10[1]; b.Gs = in10[2]; b.Ts = in10[3]; b.gaps = in10[4];
  b.cov = in10[5]; b.scoreA = in10[6]; b.scoreC = in10[7]; b.scoreG = in10[8]; b.scoreT = in10[9];
  return find_consensus( &b, cons_code );
}

/* f1: the repeat filter -- sort_fsdb / sort_fsdb_qscore (fsdb.c:240-252) followed by set_uniq_in_fsdb
   (fsdb.c:440-508) on n FragSeqs that carry only the fields those functions read.  order[k] = input
   index of the FragSeq at position k afterwards; unique[i] = its unique_best, by INPUT index. */
void refh_repeat_filter( long long n, const unsigned char* rc, const int* as, const int* ae,
                         const int* score, const int* qual_sum, const unsigned char* trimmed,
                         int use_qscore, int just_outer_coords, int tolerance,
                         long long* order, unsigned char* unique ) {
  long long i;
  FragSeq* all;
  FragSeqDB db;
  if ( n <= 0 ) return;
  all = (FragSeq*)calloc( (size_t)n, sizeof(FragSeq) );
  db.fss = (FragSeqP*)malloc( (size_t)n * sizeof(FragSeqP) );
  db.size = db.num_fss = (size_t)n;
  db.trim_sort = 0;
  for ( i = 0; i < n; i++ ) {
    all[i].rc = rc[i]; all[i].as = as[i]; all[i].ae = ae[i];
    all[i].score = score ? score[i] : 0; all[i].qual_sum = qual_sum ? qual_sum[i] : 0;
    all[i].trimmed = trimmed ? trimmed[i] : 0;
    db.fss[i] = &all[i];
  }
  if ( use_qscore ) sort_fsdb_qscore( &db ); else sort_fsdb( &db );
  set_uniq_in_fsdb( &db, just_outer_coords, (unsigned short)tolerance );
  for ( i = 0; i < n; i++ ) {
    order[i] = (long long)( db.fss[i] - all );
    unique[ order[i] ] = (unsigned char)db.fss[i]->unique_best;
  }
  free( db.fss ); free( all );
}

/* f4: trim_frag (mia.c:1318-1368) as main() sets it up (mia_main.c:692-713): adapter rows x read columns, flat matrix,
   sg5 = 1, sg3 = 0, no homopolymer discount.  out6: trimmed, trim_point, best score of the last column, abr, abc, aer */
void refh_trim( const char* read, const char* adapter, int* out6 ) {
  static AlignmentP a = NULL;
  static PSSMP flat = NULL;
  static char adapt[INIT_ALN_SEQ_LEN + 1];
  FragSeq fs;
  if ( a == NULL ) {
    a = init_alignment( INIT_ALN_SEQ_LEN, INIT_ALN_SEQ_LEN, 0, 0 );
    flat = init_flatsubmat();
    a->submat = flat;
    a->sg5 = 1; a->sg3 = 0;
  }
  strncpy( adapt, adapter, INIT_ALN_SEQ_LEN ); adapt[INIT_ALN_SEQ_LEN] = '\0';
  a->seq2 = adapt;
  a->len2 = strlen( adapt );
  pop_s2c_in_a( a );
  memset( &fs, 0, sizeof(fs) );
  strncpy( fs.seq, read, INIT_ALN_SEQ_LEN );
  fs.seq_len = strlen( fs.seq );
  trim_frag( &fs, adapt, a );
  out6[0] = fs.trimmed; out6[1] = fs.trimmed ? fs.trim_point : 0;
  out6[2] = a->m->mat[a->aer][a->aec].score;
  out6[3] = a->abr; out6[4] = a->abc; out6[5] = a->aer;
}

/* ---------------------------------------------------- FASTA / FASTQ reader
   mia_main.c:746-759: find_input_type, then read_next_seq into ONE reused FragSeq until it
   returns 0.  Every record is dumped as id \t desc \t seq \t qual_sum \n (ids and descs with
   their bytes as they are; tabs inside a description are kept -- the dump is split on the first
   and the last two tabs).  Returns the number of records, -1 if a file cannot be opened. */
long long refh_read_seqs( const char* in_fn, const char* out_fn ) {
  FILE* FF = fopen( in_fn, "r" );
  FILE* OUT = fopen( out_fn, "w" );
  FragSeqP fs = (FragSeqP)calloc( 1, sizeof(FragSeq) );
  long long n = 0;
  int code;
  if ( FF == NULL || OUT == NULL ) return -1;
  code = find_input_type( FF );
  while ( read_next_seq( FF, fs, code ) ) {
    fprintf( OUT, "%s\t%s\t%s\t%d\n", fs->id, fs->desc, fs->seq, code ? fs->qual_sum : 0 );
    n++;
  }
  fclose( FF );
  fclose( OUT );
  free( fs );
  return n;
}
