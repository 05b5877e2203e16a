/* mia_oracle.h -- CPU restatement of MIA's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (libmiagpu.so) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function
 * below against the unmodified reference compiled into oracle/_ref/ (fuzzed
 * inputs plus the reference's own fixtures test/tr1.fna + test/tf.fna), and
 * tests/golden/ holds outputs of that reference for the GPU box, where
 * /root/reference does not exist.  The reference ships no golden vectors of
 * its own.
 *
 * All file:line citations are relative to /root/reference/src.
 */
#ifndef MIA_ORACLE_H
#define MIA_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PSSM_DEPTH   15          /* params.h:22 */
#define ORC_NMAT         31          /* 2*PSSM_DEPTH+1, types.h:155-158 */
#define ORC_PSSM_INTS    (31*5*5)
#define ORC_GOP          1000        /* params.h:26 */
#define ORC_GEP          200         /* params.h:27 */
#define ORC_MAX_READ     256         /* INIT_ALN_SEQ_LEN params.h:71 */
#define ORC_ALN_STR      513         /* 2*INIT_ALN_SEQ_LEN+1, types.h:42-43 */
#define ORC_HIM          (-1073741824) /* INT_MIN/2, mia.c:751 */

/* ---- a1: PSSM (pssm.c:36-126, io.c:408-503); sm is int[31][5][5] flat,
 *      index [depth][ref_base][read_base], bases A,C,G,T,other = 0..4 */
void orc_flat_pssm( int* sm );
int  orc_parse_pssm( const char* text, int* sm );   /* 1 ok, 0 parse error */
void orc_revcom_pssm( const int* in, int* out );
int  orc_sm_depth( int row, int len );
int  orc_base_code( char b );
char orc_revcom_char( char b );                     /* map_align.c:418-431 */

/* ---- a5+a6+a7: dyn_prog + max_sg_score + find_align_begin +
 *      populate_pwaln_to_begin (mia.c:740-981, 1278-1302, 612-637, 1440-1497).
 *      mask may be NULL (= all ones).  out5 = score, abr, abc, aer, aec.
 *      score_mat / trace_mat nullable, len2*len1 row-major. Returns 1/0. */
int orc_align( const char* seq1, int len1, const char* seq2, int len2,
               const unsigned char* mask, const int* sm, int sg5,
               int* out5, char* ref_gapped, char* read_gapped,
               int* score_mat, int* trace_mat );
/* the same with mia -h (hp != 0): the homopolymer-discounted gap candidates of mia.c:882-905, hp_discount_penalty
 * (mia.c:1096-1134) and pop_hpl_and_hps (mia.c:1193-1234) over seq1 and seq2 as given */
int orc_align_hp( const char* seq1, int len1, const char* seq2, int len2,
                  const unsigned char* mask, const int* sm, int sg5, int hp,
                  int* out5, char* ref_gapped, char* read_gapped,
                  int* score_mat, int* trace_mat );
void orc_hp_runs( const char* seq, int len, int* hpl, int* hps );
int  orc_hp_penalty( int gap_len, int hplen2 );

/* ---- a2+a3: k-mer table and filter (kmer.c:18-168, 239-331) */
typedef struct orc_kmer orc_kmer;
orc_kmer* orc_kmer_build( const char* seq, long long len, int k, int soft_mask );
void      orc_kmer_free( orc_kmer* t );
int       orc_kmer_lookup( const orc_kmer* t, long long inx, unsigned int* out );
unsigned  orc_kmer_filter( const orc_kmer* f, const orc_kmer* r, int k,
                           const char* read, int read_len, int len1,
                           unsigned char* mask_f, unsigned char* mask_r );

/* ---- reference context: what mia_main.c:636-733 / 43-78 set up */
typedef struct orc_ctx orc_ctx;
/* seq is the raw reference (case preserved for -M); k<=0 => no k-mer filter */
orc_ctx* orc_ctx_new( const char* seq, int seq_len, int circular, int with_rc,
                      int k, int soft_mask, const int* sm_fwd, int distant_ref );
void     orc_ctx_set_hp( orc_ctx* c, int hp );   /* mia -h for orc_pass1 / orc_realign of this context */
void     orc_ctx_free( orc_ctx* c );
int      orc_ctx_wrap_len( const orc_ctx* c );
const char* orc_ctx_seq( const orc_ctx* c );      /* upper-cased, wrapped */
const char* orc_ctx_rcseq( const orc_ctx* c );

/* ---- a8: pass 1 for one read = new_kmer_filter + sg_align
 *      (mia_main.c:781-796, mia.c:1500-1665).
 *      out[18] laid out exactly like oracle/ref_harness.c:refh_sess_pass1.
 *      The read is NOT modified; strings are the front/back PWAlnFrag strings
 *      after revcom_PWAF and split_pwaln. */
int orc_pass1( const orc_ctx* c, const char* read, int read_len, int* out,
               char* f_ref, char* f_frag, char* b_ref, char* b_frag,
               unsigned char* mask_f, unsigned char* mask_r );

/* ---- a9: one read of reiterate_assembly (mia_main.c:178-257).
 *      read is in stored orientation (already revcomped if rc).
 *      out[8] = score, as, ae, abr, abc, aer, aec, ref_start (as/ae absolute) */
int orc_realign( const orc_ctx* c, const char* read, int read_len, int rc,
                 int as, int ae, int* out, char* ref_gapped, char* read_gapped );

/* ---- a10-a13: assembly = MapAlignment restated.  Slots persist across rounds
 *      like maln->AlnSeqArray[k]; `dropped` is sticky inside the slot (H10); the
 *      caller owns the per-read front/back slot ids (FragSeq.front_asp/back_asp),
 *      including stale back ids that reiterate_assembly never clears. */
typedef struct orc_asm orc_asm;
orc_asm* orc_asm_new( void );
void     orc_asm_free( orc_asm* a );
void     orc_asm_begin_round( orc_asm* a, int seq_len, int wrap_len );  /* mia_main.c:43-106 */
/* mia_main.c:250-276 / mia.c:1606-1643: end-adjust, split_pwaln,
 * merge_pwaln_into_maln.  start/end are pwaln start/end BEFORE the
 * "end > seq_len" adjustment.  *back_slot written only when split.
 * Returns number of AlnSeq slots used (1 or 2). */
int  orc_asm_add( orc_asm* a, const char* ref_gapped, const char* read_gapped,
                  int start, int end, int revcom, int score,
                  int* front_slot, int* back_slot );
void orc_asm_pop_smp( orc_asm* a, long long n_reads, const int* front,
                      const int* back );                    /* fsdb.c:542-619 */
/* a12: fsdb.c:269-383 */
void orc_score_cut( long long n, const int* seq_len, const int* score,
                    const unsigned char* unique_best, double* slope,
                    double* intercept );
/* a12: mia.c:418-506 for one round; builds the culled entry list
 * (front, then back if any, per read in order) and recomputes gaps[]. */
void orc_asm_cull( orc_asm* a, long long n_reads, const int* front,
                   const int* back, const int* seq_len, const int* score,
                   int hard_cut, int score_cut_set, double slope,
                   double intercept );
/* the same with FragSeq.unique_best (nullable = all 1; -u / -U): a read that is not
 * unique_best is left out of the culled list and of the regression */
void orc_asm_cull_u( orc_asm* a, long long n_reads, const int* front,
                     const int* back, const int* seq_len, const int* score,
                     const unsigned char* unique_best, int hard_cut,
                     int score_cut_set, double slope, double intercept );
/* -D (maln->distant_ref): find_alignable_len (mia.c:69-91) and the cull whose per-read threshold takes it (mia.c:460-463) */
int  orc_alignable_len( const char* ref_wrapped, int wrap_len, int seq_len, int as, int ae );
void orc_asm_cull_d( orc_asm* a, long long n_reads, const int* front,
                     const int* back, const int* seq_len, const int* score,
                     const unsigned char* unique_best, const int* alignable_len,
                     int hard_cut, int score_cut_set, double slope, double intercept );
/* a13: mia.c:515-603 over the culled entry list.  cons must hold
 * seq_len + sum(gaps) + 1 chars.  counts (nullable): 10 ints per base column:
 * As,Cs,Gs,Ts,gaps,cov,sA,sC,sG,sT */
int  orc_asm_consensus( const orc_asm* a, const int* sm_fwd, const int* sm_rc,
                        int cons_code, char* cons, int* counts );
int  orc_asm_num_slots( const orc_asm* a );
int  orc_asm_num_entries( const orc_asm* a );
int  orc_asm_entry( const orc_asm* a, int i );              /* slot id of entry i */
void orc_asm_gaps( const orc_asm* a, int* out );           /* wrap_len+1 ints */
/* out7: start,end,score,revcom,dropped,segment,n_ins ; same text encoding of
 * inserts as refh_sess_aln */
void orc_asm_slot( const orc_asm* a, int i, int* out7, char* seq, char* smp,
                   char* ins );
int  orc_find_consensus( const int* in10, int cons_code ); /* map_align.c:294-391 */
/* f4: trim_frag (mia.c:1318-1368; set-up mia_main.c:692-713): dyn_prog of the adapter
 * (rows) against the read (columns) with the flat matrix and sg5 = 1, the first
 * maximum of the LAST COLUMN in row order, find_align_begin from there.
 * out6: trimmed, trim_point (0 if not trimmed), that maximum, abr, abc, aer. */
int orc_trim( const char* read, int read_len, const char* adapter, int adapter_len, int* out6 );
/* f1: sort_fsdb / sort_fsdb_qscore (fsdb.c:13-88, 90-180, 240-252; key4 = score or
 * qual_sum) as a STABLE sort -- what glibc's qsort is while its merge buffer fits --
 * then set_uniq_in_fsdb (fsdb.c:440-508).  order[k] = input index at sorted
 * position k, unique[i] = unique_best by input index. */
void orc_repeat_filter( long long n, const unsigned char* rc, const int* as,
                        const int* ae, const int* key4,
                        const unsigned char* trimmed, int just_outer_coords,
                        int tolerance, long long* order, unsigned char* unique );

#ifdef __cplusplus
}
#endif
#endif
