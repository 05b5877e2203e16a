// hostio.hpp -- host-side data formats either side of the hot path. Plain C++17, no CUDA.
//
//   f2  miagpu_fastx_*   : the reference's FASTA / FASTQ record readers (io.c:11-25 find_input_type,
//                          io.c:46-167 read_fastq, io.c:190-281 read_fasta) restated as a cursor over a memory-mapped
//                          file that fills batch arrays (bases + offsets + ids + descs + qual_sum) ready for
//                          miagpu_upload_reads -- quirks included, because they decide which reads exist:
//                          ids cut at 100 chars with the cut character going to the description, FASTA
//                          descriptions that repeat their first character (ungetc at io.c:224), reads cut at
//                          256 bases, a record that does not start with '@' / '>' ends the input.
//   f3  miagpu_write_maln: write_ma (map_alignment.c:283-382) fed from what the device returns -- per read
//                          score / as / ae / abr and the packed run lists -- materialising each AlnSeq's
//                          seq / ins (merge_pwaln_into_maln map_align.c:866-954, split_pwaln mia.c:1376-1438),
//                          smp (pop_smp_from_FSDB fsdb.c:542-619), the list order of cull_maln_from_fsdb
//                          (mia.c:463-476) and sort_aln_frags (map_alignment.c:630, glibc's stable merge sort).
//                          Output is byte-identical to the reference's file after line 1 (a time stamp).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <string>
#include <vector>

#include "../../include/miagpu.h"

namespace miagpu { void set_error(const char* fmt, ...); }
using miagpu::set_error;

namespace hostio {

constexpr int kMaxId = 100;     // MAX_ID_LEN   params.h:17
constexpr int kMaxDesc = 128;   // MAX_DESC_LEN params.h:18
constexpr int kMaxRead = 256;   // INIT_ALN_SEQ_LEN params.h:68
constexpr int kDepth = 15;      // PSSM_DEPTH

inline bool c_space(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline int c_upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// ------------------------------------------------------------------------------------------------ f2
struct Fastx {
  const unsigned char* buf = nullptr;
  size_t len = 0, pos = 0;
  int fd = -1;
  bool mapped = false;
  int format = 0;            // 0 fasta, 1 fastq (find_input_type io.c:11-25)
  bool done = false;
  int last_qual_sum = 0;     // FragSeq is reused by the caller's loop: a record without '+' keeps the previous sum
  bool qs_set = false;       // (parallel parse) a quality line was read since this cursor started ...
  bool used_stale = false;   // ... and whether a record took last_qual_sum from before that
  // current batch
  std::vector<uint8_t> bases;
  std::vector<int64_t> off;
  std::vector<char> ids, descs;
  std::vector<int64_t> id_off, desc_off;
  std::vector<int32_t> qual_sum;
  int64_t records_total = 0;

  // the reference keeps what fgetc returned in a `char`: end of input and byte 0xFF are both -1
  int get() { return pos < len ? (int)(signed char)buf[pos++] : (pos++, -1); }
  void unget() { if (pos > 0) pos--; }

  void clear_batch() {
    bases.clear(); off.assign(1, 0); ids.clear(); descs.clear(); id_off.assign(1, 0); desc_off.assign(1, 0); qual_sum.clear();
  }

  void push(const char* id, int idn, const char* desc, int dn, const unsigned char* seq, int sn, int qs) {
    ids.insert(ids.end(), id, id + idn); ids.push_back('\0'); id_off.push_back((int64_t)ids.size());
    descs.insert(descs.end(), desc, desc + dn); descs.push_back('\0'); desc_off.push_back((int64_t)descs.size());
    bases.insert(bases.end(), seq, seq + sn); off.push_back((int64_t)bases.size());
    qual_sum.push_back(qs);
    records_total++;
  }

  // header line shared by both formats (io.c:58-91 / 203-232). Returns false at end of input inside the id.
  bool header(char* id, int& idn, char* desc, int& dn, bool fasta) {
    int c;
    idn = 0;
    for (;;) {
      c = get();
      if (c_space(c) || idn >= kMaxId) break;          // the character that ends an over-long id is consumed here
      if (c == -1) return false;
      id[idn++] = (char)c;
    }
    dn = 0;
    if (c == '\n') return true;
    while (c != '\n' && c_space(c)) c = get();
    if (fasta) unget();                                // io.c:224: the pushed-back character is stored, then read again
    while (c != '\n' && dn < kMaxDesc) { desc[dn++] = (char)c; c = get(); }
    return true;                                       // a description cut at 128 leaves the rest of its line unread
  }

  // one FASTQ record; 0 = stop, 1 = record stored
  int next_fastq() {
    int c = get();
    if (c == -1 || c != '@') return 0;
    char id[kMaxId + 1], desc[kMaxDesc + 1];
    unsigned char seq[kMaxRead + 1];
    int idn, dn, sn = 0;
    if (!header(id, idn, desc, dn, false)) return 0;
    c = get();
    while (c != '\n' && c != -1 && sn < kMaxRead) { if (!c_space(c)) seq[sn++] = (unsigned char)c_upper(c); c = get(); }
    if (sn == kMaxRead) while (c != '\n' && c != -1) c = get();
    c = get();
    if (c != '+') {                                                                // io.c:121-124: accepted as it is
      if (!qs_set) used_stale = true;
      push(id, idn, desc, dn, seq, sn, last_qual_sum);
      return 1;
    }
    do c = get(); while (c != '\n' && c != -1);
    c = get();
    int qn = 0, qs = 0;
    while (c != '\n' && c != -1 && qn < kMaxRead) { if (!c_space(c)) { qs += c - 33; qn++; } c = get(); }
    last_qual_sum = qs;
    qs_set = true;
    if (qn == kMaxRead) while (c != '\n' && c != -1) c = get();
    if (qn != sn) return 0;                                                        // io.c:162-166
    push(id, idn, desc, dn, seq, sn, qs);
    return 1;
  }

  int next_fasta() {
    int c = get();
    if (c == -1 || c != '>') return 0;
    char id[kMaxId + 1], desc[kMaxDesc + 1];
    unsigned char seq[kMaxRead + 1];
    int idn, dn, sn = 0;
    if (!header(id, idn, desc, dn, true)) return 0;
    c = get();
    while (c != '>' && c != -1 && sn < kMaxRead) { if (!c_space(c)) seq[sn++] = (unsigned char)c_upper(c); c = get(); }
    if (c != '>' && sn == kMaxRead) while (c != '>' && c != -1) c = get();
    if (c == '>') unget();
    push(id, idn, desc, dn, seq, sn, 0);
    return 1;
  }
};

// ------------------------------------------------------------------------------------------------ f3
struct Out {
  FILE* f;
  std::vector<char> b;
  explicit Out(FILE* f_) : f(f_) { if (f) b.reserve(1 << 22); }
  void flush() { if (f && !b.empty()) { fwrite(b.data(), 1, b.size(), f); b.clear(); } }   // f == nullptr: a memory buffer
  void room() { if (f && b.size() > (1u << 22) - 4096) flush(); }
  void s(const char* p) { size_t n = strlen(p); b.insert(b.end(), p, p + n); room(); }
  void s(const char* p, size_t n) { b.insert(b.end(), p, p + n); room(); }
  void ch(char c) { b.push_back(c); }
  void i(long v) {
    char t[24]; int n = 0; bool neg = v < 0; unsigned long u = neg ? 0ul - (unsigned long)v : (unsigned long)v;
    do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) b.push_back('-');
    while (n) b.push_back(t[--n]);
    room();
  }
  void kv(const char* k, long v) { s(k); i(v); ch('\n'); }
};

inline double wall_ms() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; }

struct Seg {                  // one AlnSeq of the list
  int64_t read;               // the read whose alignment fills the AlnSeq (its id, description, score ...)
  int32_t start, end;
  int32_t col0, ncol;         // slice of the read's alignment columns
  int32_t smp_n;              // characters of AlnSeq.smp: end - start + 1 (= ncol, except for a back segment whose read starts beyond seq_len)
  char seg;
  uint8_t dropped;
  // the pointer state of the one-call rounds (slots.cuh): an AlnSeq whose smp codes were written by another read's visit
  // (pop_smp_from_FSDB follows stale pointers), and AlnSeqs whose content is an older round's (frozen) alignment
  uint8_t ov = 0;
  int32_t fl = 0, total = 0, bias = 0, bf = 0;
  int32_t fz = -1, fz_score = 0, fz_rc = 0, fz_num_inputs = 1;
};
struct FrozenView { const uint8_t* bases; const uint16_t* runs; const int32_t* nruns; int32_t stride, max_runs; };
int maln_emit(const char* path, const miagpu_maln_header* hd, const miagpu_maln_reads* rd, std::vector<Seg>& segs, const FrozenView* fzv,
              int64_t* n_alnseqs_out);
// the AlnSeq(s) read i's own alignment fills this round (mia_main.c:259-276, split_pwaln mia.c:1376-1438): returns 1 when wrap-split
inline int natural_segs(const miagpu_maln_reads* rd, int L, int64_t i, Seg& f, Seg& b) {
  int start = rd->as[i];
  int end = rd->ae[i] > L ? rd->ae[i] - L : rd->ae[i];                             // mia_main.c:259-263
  int ncol = 0;
  for (int64_t r = rd->run_off[i]; r < rd->run_off[i + 1]; r++) {
    unsigned x = rd->packed[r];
    if (MIAGPU_RUN_TYPE(x) != MIAGPU_RUN_I) ncol += (int)MIAGPU_RUN_LEN(x);
  }
  f = Seg{}; b = Seg{};
  f.read = b.read = i;
  if (start > end) {                                                               // split_pwaln at seq_len
    int nf = L - start;
    if (nf < 0) nf = 0;
    if (nf > ncol) nf = ncol;
    // a read that starts beyond seq_len (as > L: the window rule can leave it there) gets a front AlnSeq of negative length and a
    // back AlnSeq whose end - start + 1 exceeds its columns; pop_smp_from_FSDB and asp_len go by end - start + 1 (fsdb.c:518-530)
    const int over = start > L ? start - L : 0;
    f.start = start; f.end = L - 1; f.col0 = 0; f.ncol = nf; f.smp_n = nf; f.seg = 'f';
    b.start = 0; b.end = end; b.col0 = nf; b.ncol = ncol - nf; b.smp_n = ncol - nf + over; b.seg = 'b';
    return 1;
  }
  f.start = start; f.end = end; f.col0 = 0; f.ncol = ncol; f.smp_n = ncol; f.seg = 'a';
  return 0;
}

inline char smp_code(int from_front, int from_back) {
  if (from_front <= kDepth) return (char)('A' + from_front);
  if (from_back < kDepth) return (char)('A' + 2 * kDepth - from_back);
  return (char)('A' + kDepth);
}

}  // namespace hostio

struct miagpu_fastx { hostio::Fastx x; };

extern "C" int miagpu_fastx_open(miagpu_fastx** out, const char* path) {
  if (!out || !path) { set_error("miagpu_fastx_open: NULL argument"); return 0; }
  int fd = open(path, O_RDONLY);
  if (fd < 0) { set_error("miagpu_fastx_open: cannot open %s", path); return 0; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); set_error("miagpu_fastx_open: cannot stat %s", path); return 0; }
  miagpu_fastx* h = new miagpu_fastx();
  h->x.fd = fd;
  h->x.len = (size_t)st.st_size;
  if (h->x.len) {
    void* p = mmap(nullptr, h->x.len, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) { close(fd); delete h; set_error("miagpu_fastx_open: mmap of %s failed", path); return 0; }
    madvise(p, h->x.len, MADV_SEQUENTIAL);
    h->x.buf = (const unsigned char*)p;
    h->x.mapped = true;
  }
  h->x.format = (h->x.len && h->x.buf[0] == '@') ? 1 : 0;
  h->x.clear_batch();
  *out = h;
  return 1;
}

extern "C" int miagpu_fastx_open_memory(miagpu_fastx** out, const void* text, int64_t len) {
  if (!out || (!text && len) || len < 0) { set_error("miagpu_fastx_open_memory: bad argument"); return 0; }
  miagpu_fastx* h = new miagpu_fastx();
  h->x.buf = (const unsigned char*)text;
  h->x.len = (size_t)len;
  h->x.format = (len && h->x.buf[0] == '@') ? 1 : 0;
  h->x.clear_batch();
  *out = h;
  return 1;
}

extern "C" void miagpu_fastx_close(miagpu_fastx* h) {
  if (!h) return;
  if (h->x.mapped) munmap((void*)h->x.buf, h->x.len);
  if (h->x.fd >= 0) close(h->x.fd);
  delete h;
}

extern "C" int miagpu_fastx_format(miagpu_fastx* h) { return h ? h->x.format : -1; }

// Speculative parallel parse of everything that is left (taken when max_reads cannot bind): the rest of the buffer is cut at guessed
// record starts, every piece is parsed by its own cursor until it reaches or passes the next cut, and a piece is KEPT only if the
// cursor before it stopped exactly on its cut without having ended the input and the piece did not need the previous record's
// quality sum (a record without '+', io.c:121-124) -- i.e. only if the reference's sequential loop would have stood at that
// byte in the same state.  The first piece that fails is parsed again, sequentially, from where the kept cursors ended.
static void fastx_parallel(hostio::Fastx& x, size_t T) {
  using hostio::Fastx;
  const size_t begin = x.pos, end = x.len;
  std::vector<size_t> cut(T + 1, end);
  cut[0] = begin;
  const char mark = x.format ? '@' : '>';
  for (size_t t = 1; t < T; t++) {
    size_t g = begin + (end - begin) / T * t;
    if (g <= cut[t - 1]) g = cut[t - 1] + 1;
    // a line that starts with the record mark; for FASTQ the line after next must start with '+' (a quality line may start with '@')
    size_t p = g;
    while (p < end) {
      const unsigned char* nl = (const unsigned char*)memchr(x.buf + p, '\n', end - p);
      if (!nl) { p = end; break; }
      p = (size_t)(nl - x.buf) + 1;
      if (p < end && x.buf[p] == (unsigned char)mark) {
        if (!x.format) break;
        const unsigned char* l2 = (const unsigned char*)memchr(x.buf + p, '\n', end - p);
        const unsigned char* l3 = l2 ? (const unsigned char*)memchr(l2 + 1, '\n', end - (size_t)(l2 + 1 - x.buf)) : nullptr;
        if (l3 && (size_t)(l3 + 1 - x.buf) < end && l3[1] == '+') break;
      }
    }
    cut[t] = p;
  }
  std::vector<Fastx> part(T);
  std::vector<std::thread> th;
  for (size_t t = 0; t < T; t++) {
    Fastx& c = part[t];
    c.buf = x.buf; c.len = x.len; c.pos = cut[t]; c.format = x.format; c.last_qual_sum = t ? 0 : x.last_qual_sum; c.qs_set = t == 0;
    c.clear_batch();
    if (cut[t] >= end && t) continue;
    th.emplace_back([&c, stop = cut[t + 1]]() {
      while (!c.done && c.pos < stop) if (!(c.format ? c.next_fastq() : c.next_fasta())) c.done = true;
    });
  }
  for (auto& t : th) t.join();
  // keep the prefix of pieces the sequential loop would have produced
  size_t kept = 1;
  while (kept < T && !part[kept - 1].done && part[kept - 1].pos == cut[kept] && cut[kept] < end && !part[kept].used_stale) kept++;
  {
    size_t nb = x.bases.size(), ni = x.ids.size(), nd = x.descs.size(), nr = x.qual_sum.size();
    for (size_t t = 0; t < kept; t++) { nb += part[t].bases.size(); ni += part[t].ids.size(); nd += part[t].descs.size(); nr += part[t].qual_sum.size(); }
    x.bases.reserve(nb); x.ids.reserve(ni); x.descs.reserve(nd); x.qual_sum.reserve(nr);
    x.off.reserve(nr + 1); x.id_off.reserve(nr + 1); x.desc_off.reserve(nr + 1);
  }
  for (size_t t = 0; t < kept; t++) {
    Fastx& c = part[t];
    const int64_t b0 = (int64_t)x.bases.size(), i0 = (int64_t)x.ids.size(), d0 = (int64_t)x.descs.size();
    x.bases.insert(x.bases.end(), c.bases.begin(), c.bases.end());
    x.ids.insert(x.ids.end(), c.ids.begin(), c.ids.end());
    x.descs.insert(x.descs.end(), c.descs.begin(), c.descs.end());
    x.qual_sum.insert(x.qual_sum.end(), c.qual_sum.begin(), c.qual_sum.end());
    for (size_t k = 1; k < c.off.size(); k++) { x.off.push_back(c.off[k] + b0); x.id_off.push_back(c.id_off[k] + i0); x.desc_off.push_back(c.desc_off[k] + d0); }
    x.records_total += c.records_total;
  }
  if (getenv("MIAGPU_TRACE")) fprintf(stderr, "[miagpu trace] fastx: %zu pieces parsed in parallel, %zu kept\n", T, kept);
  const Fastx& last = part[kept - 1];
  x.pos = last.pos; x.done = last.done;
  if (last.qs_set) x.last_qual_sum = last.last_qual_sum;
  for (size_t t = kept - 1; t > 0 && !last.qs_set; t--) if (part[t - 1].qs_set) { x.last_qual_sum = part[t - 1].last_qual_sum; break; }
}

extern "C" int miagpu_fastx_next(miagpu_fastx* h, int64_t max_reads, int64_t* n_out) {
  if (!h || !n_out || max_reads < 0) { set_error("miagpu_fastx_next: bad argument"); return 0; }
  hostio::Fastx& x = h->x;
  x.clear_batch();
  int64_t n = 0;
  {
    const size_t left = x.pos < x.len ? x.len - x.pos : 0;
    size_t par_min = (size_t)8 << 20, T = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (const char* e = getenv("MIAGPU_FASTX_PAR_MIN")) par_min = (size_t)atoll(e);        // tests: 1 = always
    if (const char* e = getenv("MIAGPU_FASTX_THREADS")) T = (size_t)std::max(1, atoi(e));
    if (!x.done && T > 1 && left >= par_min && left > 4 * T && (uint64_t)max_reads >= left / 2 + 1) {   // cannot bind: a record takes at least 2 bytes (">\n")
      fastx_parallel(x, T);
      n = (int64_t)x.qual_sum.size();
    }
  }
  while (!x.done && n < max_reads) {
    int r = x.format ? x.next_fastq() : x.next_fasta();
    if (!r) { x.done = true; break; }
    n++;
  }
  *n_out = n;
  return 1;
}

extern "C" int miagpu_fastx_batch(miagpu_fastx* h, const uint8_t** bases, const int64_t** offsets, const char** ids,
                                  const int64_t** id_off, const char** descs, const int64_t** desc_off, const int32_t** qual_sum) {
  if (!h) { set_error("miagpu_fastx_batch: NULL handle"); return 0; }
  hostio::Fastx& x = h->x;
  if (bases) *bases = x.bases.data();
  if (offsets) *offsets = x.off.data();
  if (ids) *ids = x.ids.data();
  if (id_off) *id_off = x.id_off.data();
  if (descs) *descs = x.descs.data();
  if (desc_off) *desc_off = x.desc_off.data();
  if (qual_sum) *qual_sum = x.qual_sum.data();
  return 1;
}

// a1: read_pssm (io.c:408-503).  31 blocks of "# Matrix for position ..." + four rows of four tab-separated integers + one
// more line; the block after the 15th must say MIDDLE.  Column 4 (read base not ACGT) = N_SCORE -100, row 4 (reference base
// not ACGT) = NR_SCORE -10 (params.h:30-31).  The reference exits on a malformed file; here: return 0 with the message.
extern "C" int miagpu_read_pssm(const char* path, int32_t* fwd) {
  using namespace hostio;
  if (!path || !fwd) { set_error("miagpu_read_pssm: NULL argument"); return 0; }
  FILE* f = fopen(path, "r");
  if (!f) { set_error("miagpu_read_pssm: cannot open %s", path); return 0; }
  char line[4096];
  for (int d = 0; d <= 2 * kDepth; d++) {
    const char* want = (d == kDepth) ? "# Matrix for position: MIDDLE" : (d < kDepth ? "# Matrix for position" : "# Matrix for position:");
    if (!fgets(line, sizeof line, f) || !strstr(line, want)) {
      fclose(f);
      set_error("miagpu_read_pssm: %s: block %d does not start with \"%s\"", path, d + 1, want);
      return 0;
    }
    int32_t* m = fwd + d * 25;
    for (int row = 0; row < 4; row++) {
      int v[4] = {0, 0, 0, 0};
      if (!fgets(line, sizeof line, f) || sscanf(line, "%d\t%d\t%d\t%d", &v[0], &v[1], &v[2], &v[3]) != 4) {
        fclose(f);
        set_error("miagpu_read_pssm: %s: block %d row %d is not four integers", path, d + 1, row + 1);
        return 0;
      }
      for (int col = 0; col < 4; col++) m[row * 5 + col] = v[col];
      m[row * 5 + 4] = -100;
    }
    for (int col = 0; col < 5; col++) m[4 * 5 + col] = -10;
    if (!fgets(line, sizeof line, f)) line[0] = 0;      // the line after a block is skipped whatever it holds
  }
  fclose(f);
  return 1;
}

extern "C" int miagpu_maln_ref_size(int ref_len, int circular) {
  // reiterate_assembly sets size = len + 1 (mia_main.c:67), add_ref_wrap doubles it until the wrap fits (mia.c:669-675)
  long size = (long)ref_len + 1;
  if (circular) {
    int wrap = ref_len < hostio::kMaxRead ? ref_len : hostio::kMaxRead;
    while ((long)ref_len + wrap >= size) size *= 2;
  }
  return (int)size;
}

extern "C" int miagpu_write_maln(const char* path, const miagpu_maln_header* hd, const miagpu_maln_reads* rd, int64_t* n_alnseqs_out) {
  using namespace hostio;
  if (!path || !hd || !rd) { set_error("miagpu_write_maln: NULL argument"); return 0; }
  if (!hd->ref_seq || hd->ref_len <= 0 || !hd->gaps || !hd->fpsm || !hd->rpsm || !hd->ref_id || !hd->ref_desc) {
    set_error("miagpu_write_maln: incomplete header"); return 0;
  }
  const int64_t n = rd->n;
  if (n < 0 || (n && (!rd->bases || !rd->offsets || !rd->rc || !rd->score || !rd->as || !rd->ae || !rd->abr || !rd->run_off || !rd->packed ||
                      !rd->ids || !rd->id_off))) {
    set_error("miagpu_write_maln: incomplete read arrays"); return 0;
  }
  const int L = hd->ref_len;
  // ---- the AlnSeq list in FSDB order (cull_maln_from_fsdb mia.c:463-476), then sort_aln_frags
  std::vector<Seg> segs;
  segs.reserve((size_t)n + 16);
  for (int64_t k = 0; k < n; k++) {
    const int64_t i = rd->fsdb_order ? rd->fsdb_order[k] : k;                      // position k of fsdb->fss holds read i
    if (i < 0 || i >= n) { set_error("miagpu_write_maln: fsdb_order[%lld] = %lld is not a read", (long long)k, (long long)i); return 0; }
    if (rd->unique_best && !rd->unique_best[i]) continue;
    uint8_t df = rd->dropped_front ? rd->dropped_front[i] : 0, db = rd->dropped_back ? rd->dropped_back[i] : df;
    Seg f, b;
    if (natural_segs(rd, L, i, f, b)) {
      f.dropped = df != 0; b.dropped = db != 0;
      segs.push_back(f); segs.push_back(b);
    } else {
      f.dropped = df != 0;
      segs.push_back(f);
    }
  }
  return hostio::maln_emit(path, hd, rd, segs, nullptr, n_alnseqs_out);
}

namespace hostio {
int maln_emit(const char* path, const miagpu_maln_header* hd, const miagpu_maln_reads* rd, std::vector<Seg>& segs, const FrozenView* fzv,
              int64_t* n_alnseqs_out) {
  const int L = hd->ref_len;
  const bool trace = getenv("MIAGPU_TRACE") != nullptr;
  double t_a = wall_ms(), t_b, t_fmt = 0, t_wr = 0;
  std::vector<uint32_t> order(segs.size());
  for (size_t k = 0; k < order.size(); k++) order[k] = (uint32_t)k;
  // sort_aln_frags: stable by (start, end).  Keys are reference columns, so two stable counting passes (end, then start) do it in
  // O(n); anything outside [0, L + 2 * 256] (cannot come out of the window rule) takes the comparison sort instead.
  {
    const int32_t KMAX = L + 2 * kMaxRead + 2;
    bool small = true;
    for (const Seg& sg : segs) if (sg.start < 0 || sg.start >= KMAX || sg.end < 0 || sg.end >= KMAX) { small = false; break; }
    if (small && segs.size() > 4096) {
      std::vector<uint32_t> tmp(order.size()), cnt((size_t)KMAX + 1);
      for (int pass = 0; pass < 2; pass++) {
        std::fill(cnt.begin(), cnt.end(), 0u);
        for (uint32_t k : order) cnt[(size_t)(pass ? segs[k].start : segs[k].end) + 1]++;
        for (size_t v = 1; v < cnt.size(); v++) cnt[v] += cnt[v - 1];
        for (uint32_t k : order) tmp[cnt[(size_t)(pass ? segs[k].start : segs[k].end)]++] = k;
        order.swap(tmp);
      }
    } else {
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (segs[a].start != segs[b].start) return segs[a].start < segs[b].start;
        return segs[a].end < segs[b].end;
      });
    }
  }

  t_b = wall_ms();
  FILE* f = fopen(path, "w");
  if (!f) { set_error("miagpu_write_maln: cannot open %s for writing", path); return 0; }
  Out o(f);
  time_t t = time(nullptr);
  char stamp[64];
  struct tm tmv;
  localtime_r(&t, &tmv);
  asctime_r(&tmv, stamp);
  o.s("/* map_alignment [V1.0] */ "); o.s(stamp);
  o.kv("MALN_NAS ", (long)segs.size());
  o.kv("MALN_SIZ ", hd->maln_size);
  o.kv("MALN_COC ", hd->cons_code);
  o.s("__REFERENCE__\n");
  o.s("ID "); o.s(hd->ref_id); o.ch('\n');
  o.s("DESC "); o.s(hd->ref_desc); o.ch('\n');
  o.kv("LEN ", L);
  o.kv("SIZE ", hd->ref_size > 0 ? hd->ref_size : miagpu_maln_ref_size(L, hd->circular));
  o.s("SEQ "); o.s(hd->ref_seq, (size_t)L); o.ch('\n');
  o.s("GAPS");
  for (int i = 0; i < L; i++) { o.ch(' '); o.i(hd->gaps[i]); }
  o.ch('\n');
  o.s("__PSSM__\n");
  o.kv("DEPTH ", kDepth);
  for (int which = 0; which < 2; which++) {
    const int32_t* sm = which ? hd->rpsm : hd->fpsm;
    o.s(which ? "RPSM:\n" : "FPSM:\n");
    for (int d = 0; d <= 2 * kDepth; d++) {
      for (int row = 0; row < 5; row++) {
        for (int col = 0; col < 5; col++) { if (col) o.ch(' '); o.i(sm[(d * 5 + row) * 5 + col]); }
        o.ch('\n');
      }
      o.ch('\n');
    }
  }
  o.s("__ALNSEQS__\n");

  // ---- per AlnSeq: columns, inserts, smp from the run list
  // Formatting is the expensive part (about 1.3 us per AlnSeq on one core): ranges of the sorted list are formatted into memory
  // buffers by worker threads, wave after wave, and written in order.
  std::atomic<long long> bad_read{-1};
  auto format_range = [&](size_t lo, size_t hi, Out& o) {
    std::vector<char> colc, smp, smp_ov;         // per alignment column of the read: AlnSeq.seq character
    std::vector<int32_t> ins_at, ins_len, actv;  // per column: read row and length of the insert in front of it; pop_smp's running position
    std::string idbuf;
    int64_t cached = -1;
    int front_total = 0, back_total = 0, nfront = 0;
    for (size_t kk = lo; kk < hi; kk++) {
      const Seg& sg = segs[order[kk]];
      const int64_t i = sg.read;
      const uint8_t* read = sg.fz >= 0 ? fzv->bases + (int64_t)sg.fz * fzv->stride : rd->bases + rd->offsets[i];
      const int rlen = sg.fz >= 0 ? fzv->stride : (int)(rd->offsets[i + 1] - rd->offsets[i]);
      const int64_t key = sg.fz >= 0 ? -2 - (int64_t)sg.fz : i;
      if (cached != key) {
        cached = key;
        colc.clear(); ins_at.clear(); ins_len.clear();
        int row = sg.fz >= 0 ? 0 : rd->abr[i], pend_at = 0, pend_len = 0;
        const int64_t r_lo = sg.fz >= 0 ? 0 : rd->run_off[i], r_hi = sg.fz >= 0 ? fzv->nruns[sg.fz] : rd->run_off[i + 1];
        const uint16_t* rl = sg.fz >= 0 ? fzv->runs + (int64_t)sg.fz * fzv->max_runs : rd->packed;
        for (int64_t r = r_lo; r < r_hi; r++) {
          unsigned x = rl[r];
          int ty = (int)MIAGPU_RUN_TYPE(x), ln = (int)MIAGPU_RUN_LEN(x);
          if (ty == MIAGPU_RUN_I) {
            if (!pend_len) pend_at = row;
            pend_len += ln; row += ln;
            continue;
          }
          for (int q = 0; q < ln; q++) {
            ins_at.push_back(pend_at); ins_len.push_back(pend_len); pend_len = 0;
            if (ty == MIAGPU_RUN_M) { colc.push_back(row < rlen ? (char)read[row] : '?'); row++; }
            else colc.push_back('-');
          }
        }
        if (row > rlen || (sg.fz < 0 && rd->abr[i] < 0)) { bad_read.store((long long)i); return; }
        // asp_len of the front and back AlnSeq (fsdb.c:518-530): columns + inserted bases (deletions count as sequence)
        int tot = (int)colc.size();
        int s0 = sg.fz >= 0 ? 0 : rd->as[i], e0 = sg.fz >= 0 ? tot : (rd->ae[i] > L ? rd->ae[i] - L : rd->ae[i]);
        nfront = tot;
        int over = 0;                                                   // see the Seg list: as > L
        front_total = tot; back_total = 0;
        if (s0 > e0) {
          nfront = L - s0; if (nfront < 0) nfront = 0; if (nfront > tot) nfront = tot;
          over = s0 > L ? s0 - L : 0;
          front_total = L - s0;                                         // end - start + 1 of the front AlnSeq: negative when as > L
          back_total = tot - nfront + over;
        }
        for (int c = 0; c < tot; c++) (c < nfront ? front_total : back_total) += ins_len[c];
        // smp over front then back with one running position (fsdb.c:556-616); the `over` positions past the back AlnSeq's string
        // are taken as the reference finds them after a fresh merge: no insert, not a '-'
        smp.resize(colc.size() + (size_t)over);
        actv.resize(colc.size() + (size_t)over);
        int act = 0;
        for (int c = 0; c < tot + over; c++) {
          if (c < tot) act += ins_len[c];
          actv[c] = act;
          int from_front = (c < nfront) ? act : front_total + act;      // fsdb.c:596: the back segment adds the front's length again
          int from_back = front_total + back_total - act - 1;
          smp[c] = smp_code(from_front, from_back);
          if (c >= tot || colc[c] != '-') act++;
        }
      }
      const char* smp_of = smp.data() + sg.col0;
      if (sg.ov) {                                   // the codes another read's visit left in this AlnSeq (fsdb.c:563-614 through a stale pointer)
        smp_ov.resize((size_t)sg.smp_n);
        for (int c = 0; c < sg.smp_n; c++) {
          const int a = sg.bias + actv[std::min<size_t>((size_t)sg.col0 + c, actv.size() - 1)];
          smp_ov[c] = smp_code(sg.bf ? sg.fl + a : a, sg.total - a - 1);
        }
        smp_of = smp_ov.data();
      }
      // id, with split_pwaln's suffix rule (mia.c:1389-1398)
      const char* id = rd->ids + rd->id_off[i];
      idbuf.assign(id);
      if (sg.seg != 'a') {
        size_t cut = std::min<size_t>(idbuf.size(), (size_t)kMaxId - 1);
        idbuf.resize(cut);
        idbuf += (sg.seg == 'f') ? "_f" : "_b";
      }
      o.s("ID "); o.s(idbuf.c_str()); o.ch('\n');
      o.s("DESC "); if (rd->descs && rd->desc_off) o.s(rd->descs + rd->desc_off[i]); o.ch('\n');
      o.kv("SCORE ", sg.fz >= 0 ? sg.fz_score : rd->score[i]);
      o.kv("NUM_INPUTS ", sg.fz >= 0 ? sg.fz_num_inputs : rd->num_inputs ? rd->num_inputs[i] : 1);
      o.kv("START ", sg.start);
      o.kv("END ", sg.end);
      o.kv("RC ", (sg.fz >= 0 ? sg.fz_rc : rd->rc[i]) ? 1 : 0);
      o.kv("TR ", (rd->trimmed && rd->trimmed[i]) ? 1 : 0);
      o.kv("DR ", sg.dropped);
      o.s("SEG "); o.ch(sg.seg); o.ch('\n');
      o.s("SEQ "); o.s(colc.data() + sg.col0, (size_t)sg.ncol); o.ch('\n');
      o.s("SMP "); o.s(smp_of, (size_t)sg.smp_n); o.ch('\n');
      o.s("INS_POS");
      for (int c = 0; c < sg.ncol; c++) {
        int g = sg.col0 + c;
        if (ins_len[g]) { o.ch(' '); o.i(c); o.ch(' '); o.s((const char*)read + ins_at[g], (size_t)ins_len[g]); }
      }
      o.ch('\n');
    }
  };
  const size_t total = order.size();
  unsigned hw = std::thread::hardware_concurrency();
  size_t T = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 1), (size_t)32, (total + 1023) / 1024}));
  if (const char* e = getenv("MIAGPU_MALN_THREADS")) T = std::max(1, atoi(e));              // tests: 1 = one worker
  const size_t per = std::max<size_t>(1, std::min<size_t>(16384, (total + T - 1) / T));
  o.flush();
  // two sets of buffers: the workers format wave k + 1 while this thread writes wave k
  std::vector<Out> bufs[2];
  for (int w = 0; w < 2; w++) for (size_t t = 0; t < T; t++) bufs[w].emplace_back((FILE*)nullptr);
  auto launch = [&](size_t base, int w, std::vector<std::thread>& th) {
    for (size_t t = 0; t < T; t++) {
      size_t lo = std::min(total, base + t * per), hi = std::min(total, lo + per);
      bufs[w][t].b.clear();
      if (lo < hi) th.emplace_back(format_range, lo, hi, std::ref(bufs[w][t]));
    }
  };
  std::vector<std::thread> th_cur, th_next;
  int cur = 0;
  if (total) launch(0, 0, th_cur);
  for (size_t base = 0; base < total; base += T * per) {
    const size_t next = base + T * per;
    if (next < total && bad_read.load() < 0) launch(next, cur ^ 1, th_next);
    const double t0 = wall_ms();
    for (auto& x : th_cur) x.join();
    const double t1 = wall_ms();
    for (size_t t = 0; t < T; t++) if (!bufs[cur][t].b.empty()) fwrite(bufs[cur][t].b.data(), 1, bufs[cur][t].b.size(), f);
    t_fmt += t1 - t0; t_wr += wall_ms() - t1;
    th_cur.clear();
    th_cur.swap(th_next);
    cur ^= 1;
    if (th_cur.empty()) break;
  }
  if (trace) fprintf(stderr, "[miagpu trace] write_maln: sort %.1f ms, waiting for the formatters (%zu threads) %.1f ms, fwrite %.1f ms\n", t_b - t_a, T, t_fmt, t_wr);
  if (bad_read.load() >= 0) {
    fclose(f);
    unlink(path);                                    // no partial file
    set_error("miagpu_write_maln: the runs of read %lld overrun its bases", bad_read.load());
    return 0;
  }
  o.flush();
  bool ok = !ferror(f);
  if (fclose(f) != 0) ok = false;
  if (!ok) { unlink(path); set_error("miagpu_write_maln: write to %s failed", path); return 0; }
  if (n_alnseqs_out) *n_alnseqs_out = (int64_t)segs.size();
  return 1;
}
}  // namespace hostio
