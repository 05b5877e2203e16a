"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``  -> oracle/libmia_oracle.so  (our restatement, oracle/mia_oracle.c)
* ``Ref``     -> oracle/_ref/libmia_ref.so (the unmodified reference + oracle/ref_harness.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmia_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmia_ref.so")
REF_MIA = os.path.join(HERE, "_ref", "mia")
REF_MA = os.path.join(HERE, "_ref", "ma")
REFERENCE_ROOT = "/root/reference"

c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)


def build(ref=True):
    """(Re)build the oracle library and, when the reference sources exist, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if ref and os.path.exists(os.path.join(REFERENCE_ROOT, "src", "mia.c")):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        write_r_mt()


R_MT = os.path.join(HERE, "_ref", "r_mt.fa")


def write_r_mt():
    """R-mt of SURVEY.md 8d: the reference's built-in mt311 consensus (src/mt311.c) with the optional (lower-case) bases
    removed and every IUPAC code replaced by its alphabetically first base -> 16,517 bp with a real mitochondrial
    composition and real low-complexity stretches.  Derived here, where /root/reference exists; the file lives in the
    git-ignored oracle/_ref/ and travels to the GPU box with the other built artefacts."""
    import re
    src = open(os.path.join(REFERENCE_ROOT, "src", "mt311.c")).read()
    seq = "".join(re.findall(r'"([A-Za-z]+)"', src[src.index("mt311_sequence"):]))
    first = dict(R="A", Y="C", S="C", W="A", K="G", M="A", B="C", D="A", H="A", V="A", N="A")
    out = "".join(first.get(ch, ch) for ch in seq if not ch.islower())
    assert set(out) <= set("ACGT"), sorted(set(out))
    os.makedirs(os.path.dirname(R_MT), exist_ok=True)
    with open(R_MT, "w") as f:
        f.write(">R-mt mt311 consensus, lower-case removed, IUPAC -> first base\n")
        for i in range(0, len(out), 60):
            f.write(out[i:i + 60] + "\n")
    return len(out)


def have_ref():
    return os.path.exists(REF_SO)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _bp(a):
    return a.ctypes.data_as(c_ubyte_p)


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


class _AlignMixin:
    """Shared shape of the align() result."""

    @staticmethod
    def _res(out5, rg, fg):
        return dict(score=int(out5[0]), abr=int(out5[1]), abc=int(out5[2]), aer=int(out5[3]), aec=int(out5[4]),
                    ref_gapped=rg.value.decode(), read_gapped=fg.value.decode())


class Oracle(_AlignMixin):
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_kmer_build.restype = C.c_void_p
        L.orc_kmer_build.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_int]
        L.orc_kmer_free.argtypes = [C.c_void_p]
        L.orc_kmer_lookup.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        L.orc_kmer_filter.restype = C.c_uint
        L.orc_kmer_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, c_ubyte_p, c_ubyte_p]
        L.orc_ctx_new.restype = C.c_void_p
        L.orc_ctx_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int]
        L.orc_ctx_free.argtypes = [C.c_void_p]
        L.orc_ctx_wrap_len.argtypes = [C.c_void_p]
        L.orc_ctx_seq.restype = C.c_char_p
        L.orc_ctx_seq.argtypes = [C.c_void_p]
        L.orc_ctx_rcseq.restype = C.c_char_p
        L.orc_ctx_rcseq.argtypes = [C.c_void_p]
        L.orc_pass1.argtypes = [C.c_void_p, C.c_char_p, C.c_int, c_int_p] + [C.c_char_p] * 4 + [c_ubyte_p, c_ubyte_p]
        L.orc_realign.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, C.c_char_p, C.c_char_p]
        L.orc_asm_new.restype = C.c_void_p
        L.orc_asm_new.argtypes = []
        L.orc_asm_free.argtypes = [C.c_void_p]
        L.orc_asm_begin_round.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_asm_add.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p]
        L.orc_asm_pop_smp.argtypes = [C.c_void_p, C.c_longlong, c_int_p, c_int_p]
        L.orc_score_cut.argtypes = [C.c_longlong, c_int_p, c_int_p, c_ubyte_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.orc_asm_cull.argtypes = [C.c_void_p, C.c_longlong, c_int_p, c_int_p, c_int_p, c_int_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_asm_consensus.argtypes = [C.c_void_p, c_int_p, c_int_p, C.c_int, C.c_char_p, c_int_p]
        L.orc_asm_num_slots.argtypes = [C.c_void_p]
        L.orc_asm_num_entries.argtypes = [C.c_void_p]
        L.orc_asm_entry.argtypes = [C.c_void_p, C.c_int]
        L.orc_asm_gaps.argtypes = [C.c_void_p, c_int_p]
        L.orc_asm_slot.argtypes = [C.c_void_p, C.c_int, c_int_p, C.c_char_p, C.c_char_p, C.c_char_p]
        L.orc_revcom_char.restype = C.c_char
        L.orc_revcom_char.argtypes = [C.c_char]

    # -- a1
    def flat_pssm(self):
        sm = np.zeros(775, np.int32)
        self.lib.orc_flat_pssm(_ip(sm))
        return sm

    def parse_pssm(self, text):
        sm = np.zeros(775, np.int32)
        if not self.lib.orc_parse_pssm(_b(text), _ip(sm)):
            raise ValueError("matrix text did not parse")
        return sm

    def revcom_pssm(self, sm):
        out = np.zeros(775, np.int32)
        self.lib.orc_revcom_pssm(_ip(np.ascontiguousarray(sm, np.int32)), _ip(out))
        return out

    def sm_depth(self, row, length):
        return self.lib.orc_sm_depth(row, length)

    def revcom(self, s):
        return "".join(self.lib.orc_revcom_char(_b(ch)).decode() for ch in reversed(s))

    # -- a5..a7
    def align(self, seq1, seq2, sm, sg5=1, mask=None, matrices=False, hp=0):
        """hp: mia -h (homopolymer-discounted gap candidates, mia.c:882-905)"""
        seq1, seq2 = _b(seq1), _b(seq2)
        out5 = np.zeros(5, np.int32)
        rg, fg = C.create_string_buffer(520), C.create_string_buffer(520)
        sm = np.ascontiguousarray(sm, np.int32)
        m = None if mask is None else _bp(np.ascontiguousarray(mask, np.uint8))
        S = T = None
        if matrices:
            S = np.zeros((len(seq2), len(seq1)), np.int32)
            T = np.zeros((len(seq2), len(seq1)), np.int32)
        self.lib.orc_align_hp.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, c_ubyte_p, c_int_p, C.c_int, C.c_int, c_int_p,
                                          C.c_char_p, C.c_char_p, c_int_p, c_int_p]
        ok = self.lib.orc_align_hp(seq1, len(seq1), seq2, len(seq2), m, _ip(sm), sg5, int(hp), _ip(out5), rg, fg,
                                   None if S is None else _ip(S), None if T is None else _ip(T))
        res = self._res(out5, rg, fg)
        res["ok"] = ok
        if matrices:
            res["S"], res["T"] = S, T
        return res

    # -- a2, a3
    def kmer_build(self, seq, k, soft_mask=0):
        seq = _b(seq)
        return self.lib.orc_kmer_build(seq, len(seq), k, soft_mask)

    def kmer_free(self, t):
        self.lib.orc_kmer_free(t)

    def kmer_lookup(self, t, inx):
        out = np.zeros(128, np.uint32)
        n = self.lib.orc_kmer_lookup(t, inx, out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def kmer_filter(self, ft, rt, k, read, len1):
        read = _b(read)
        mf, mr = np.ones(len1, np.uint8), np.ones(len1, np.uint8)
        hits = self.lib.orc_kmer_filter(ft, rt, k, read, len(read), len1, _bp(mf), _bp(mr))
        return hits, mf, mr

    # -- context / pass 1 / realign
    def ctx_new(self, seq, circular, sm_fwd, with_rc=1, k=0, soft_mask=0, distant_ref=0, hp=0):
        seq = _b(seq)
        sm = np.ascontiguousarray(sm_fwd, np.int32)
        c = self.lib.orc_ctx_new(seq, len(seq), circular, with_rc, k, soft_mask, _ip(sm), distant_ref)
        if hp:
            self.lib.orc_ctx_set_hp.argtypes = [C.c_void_p, C.c_int]
            self.lib.orc_ctx_set_hp(c, 1)
        return c

    def ctx_free(self, c):
        self.lib.orc_ctx_free(c)

    def ctx_seq(self, c):
        return self.lib.orc_ctx_seq(c).decode()

    def pass1(self, ctx, read, want_masks=False):
        read = _b(read)
        out = np.zeros(18, np.int32)
        bufs = [C.create_string_buffer(520) for _ in range(4)]
        mf = mr = None
        if want_masks:
            n = self.lib.orc_ctx_wrap_len(ctx)
            mf, mr = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.lib.orc_pass1(ctx, read, len(read), _ip(out), *bufs, None if mf is None else _bp(mf), None if mr is None else _bp(mr))
        return _p1_dict(out, bufs, mf, mr)

    def realign(self, ctx, read, rc, as_, ae):
        read = _b(read)
        out = np.zeros(8, np.int32)
        rg, fg = C.create_string_buffer(520), C.create_string_buffer(520)
        ok = self.lib.orc_realign(ctx, read, len(read), rc, as_, ae, _ip(out), rg, fg)
        return dict(ok=ok, score=int(out[0]), as_=int(out[1]), ae=int(out[2]), abr=int(out[3]), abc=int(out[4]),
                    aer=int(out[5]), aec=int(out[6]), ref_start=int(out[7]),
                    ref_gapped=rg.value.decode(), read_gapped=fg.value.decode())

    # -- assembly
    def asm_new(self):
        return self.lib.orc_asm_new()

    def asm_free(self, a):
        self.lib.orc_asm_free(a)

    def asm_begin_round(self, a, seq_len, wrap_len):
        self.lib.orc_asm_begin_round(a, seq_len, wrap_len)

    def asm_add(self, a, ref_gapped, read_gapped, start, end, revcom, score):
        """Returns (front_slot, back_slot or None)."""
        f, b = C.c_int(-1), C.c_int(-1)
        n = self.lib.orc_asm_add(a, _b(ref_gapped), _b(read_gapped), start, end, revcom, score, C.byref(f), C.byref(b))
        return f.value, (b.value if n == 2 else None)

    def asm_pop_smp(self, a, front, back):
        f, b = np.ascontiguousarray(front, np.int32), np.ascontiguousarray(back, np.int32)
        self.lib.orc_asm_pop_smp(a, len(f), _ip(f), _ip(b))

    def score_cut(self, seq_len, score):
        s, n = C.c_double(), C.c_double()
        sl, sc = np.ascontiguousarray(seq_len, np.int32), np.ascontiguousarray(score, np.int32)
        self.lib.orc_score_cut(len(sl), _ip(sl), _ip(sc), None, C.byref(s), C.byref(n))
        return s.value, n.value

    def asm_cull(self, a, front, back, seq_len, score, hard_cut=0, score_cut_set=0, slope=200.0, intercept=0.0, unique_best=None,
                 alignable_len=None):
        """cull_maln_from_fsdb; alignable_len (per read, -D): the length the threshold takes (find_alignable_len)"""
        f, b = np.ascontiguousarray(front, np.int32), np.ascontiguousarray(back, np.int32)
        sl, sc = np.ascontiguousarray(seq_len, np.int32), np.ascontiguousarray(score, np.int32)
        uq = None if unique_best is None else np.ascontiguousarray(unique_best, np.uint8)
        al = None if alignable_len is None else np.ascontiguousarray(alignable_len, np.int32)
        fn = self.lib.orc_asm_cull_d
        fn.argtypes = [C.c_void_p, C.c_longlong] + [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_double, C.c_double]
        fn.restype = None
        fn(a, len(sl), f.ctypes.data, b.ctypes.data, sl.ctypes.data, sc.ctypes.data, None if uq is None else uq.ctypes.data,
           None if al is None else al.ctypes.data, hard_cut, score_cut_set, slope, intercept)

    def alignable_len(self, ref_wrapped, seq_len, as_, ae):
        fn = self.lib.orc_alignable_len
        fn.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
        rw = _b(ref_wrapped)
        return fn(rw, len(rw), seq_len, as_, ae)

    def asm_consensus(self, a, sm_fwd, sm_rc, cons_code, seq_len, max_extra=1 << 20, counts=False):
        buf = C.create_string_buffer(seq_len + max_extra + 1)
        cnt = np.zeros((seq_len, 10), np.int32) if counts else None
        self.lib.orc_asm_consensus(a, _ip(np.ascontiguousarray(sm_fwd, np.int32)), _ip(np.ascontiguousarray(sm_rc, np.int32)),
                                   cons_code, buf, None if cnt is None else _ip(cnt))
        return (buf.value.decode(), cnt) if counts else buf.value.decode()

    def asm_gaps(self, a, wrap_len):
        g = np.zeros(wrap_len + 1, np.int32)
        self.lib.orc_asm_gaps(a, _ip(g))
        return g

    def asm_slot(self, a, i):
        o = np.zeros(7, np.int32)
        seq, smp, ins = C.create_string_buffer(520), C.create_string_buffer(520), C.create_string_buffer(4096)
        self.lib.orc_asm_slot(a, i, _ip(o), seq, smp, ins)
        return dict(start=int(o[0]), end=int(o[1]), score=int(o[2]), revcom=int(o[3]), dropped=int(o[4]),
                    segment=chr(o[5]), seq=seq.value.decode(), smp=smp.value.decode(), ins=ins.value.decode())

    def asm_entries(self, a):
        """The culled list (= culled_maln->AlnSeqArray before sort_aln_frags)."""
        return [self.asm_slot(a, self.lib.orc_asm_entry(a, i)) for i in range(self.lib.orc_asm_num_entries(a))]

    def trim(self, read, adapter):
        """f4: trim_frag -> dict(trimmed, trim_point, score, abr, abc, aer)"""
        out = np.zeros(6, np.int32)
        f = self.lib.orc_trim
        f.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, c_int_p]
        ok = f(_b(read), len(read), _b(adapter), len(adapter), _ip(out))
        assert ok
        return dict(zip(("trimmed", "trim_point", "score", "abr", "abc", "aer"), map(int, out)))

    def repeat_filter(self, rc, as_, ae, key4, trimmed=None, just_outer_coords=1, tolerance=0):
        """f1: (order int64[n], unique uint8[n] by input index)"""
        n = len(rc)
        rc = np.ascontiguousarray(rc, np.uint8); as_ = np.ascontiguousarray(as_, np.int32); ae = np.ascontiguousarray(ae, np.int32)
        key4 = np.ascontiguousarray(key4, np.int32)
        tr = None if trimmed is None else np.ascontiguousarray(trimmed, np.uint8)
        order, uniq = np.zeros(n, np.int64), np.zeros(n, np.uint8)
        f = self.lib.orc_repeat_filter
        f.argtypes = [C.c_longlong] + [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        f.restype = None
        f(n, rc.ctypes.data, as_.ctypes.data, ae.ctypes.data, key4.ctypes.data, None if tr is None else tr.ctypes.data,
          int(just_outer_coords), int(tolerance), order.ctypes.data, uniq.ctypes.data)
        return order, uniq

    def find_consensus(self, counts10, cons_code=1):
        c = np.ascontiguousarray(counts10, np.int32)
        return chr(self.lib.orc_find_consensus(_ip(c), cons_code))


def _p1_dict(out, bufs, mf=None, mr=None):
    keys = ["hits", "added", "score", "rc", "as_", "ae", "strand_known", "fw_score", "rc_score", "start", "end",
            "split", "b_start", "b_end", "abr", "abc", "aer", "aec"]
    d = {k: int(v) for k, v in zip(keys, out)}
    d["f_ref"], d["f_frag"], d["b_ref"], d["b_frag"] = (b.value.decode() for b in bufs)
    if mf is not None:
        d["mask_f"], d["mask_r"] = mf, mr
    return d


class Ref(_AlignMixin):
    """The unmodified reference, driven through oracle/ref_harness.c."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build(ref=True)
        if not os.path.exists(REF_SO):
            raise RuntimeError("oracle/_ref/libmia_ref.so is missing and /root/reference is not here to build it")
        L = self.lib = C.CDLL(REF_SO)
        L.refh_align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, c_ubyte_p, c_int_p, C.c_int, c_int_p,
                                 C.c_char_p, C.c_char_p, c_int_p, c_int_p]
        L.refh_time_realign.restype = C.c_double
        L.refh_time_realign.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_longlong), c_int_p, c_int_p, c_int_p,
                                        c_int_p, c_int_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.refh_kmer_new.restype = C.c_void_p
        L.refh_kmer_new.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_int]
        L.refh_kmer_lookup.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p]
        L.refh_kmer_free.argtypes = [C.c_void_p, C.c_int]
        L.refh_kmer_filter.restype = C.c_uint
        L.refh_kmer_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int, c_ubyte_p, c_ubyte_p]
        L.refh_sess_new.restype = C.c_void_p
        L.refh_sess_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int]
        L.refh_sess_set_cut.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.refh_sess_pass1.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, c_int_p] + [C.c_char_p] * 4
        L.refh_sess_set_next_qual_sum.argtypes = [C.c_void_p, C.c_int]
        L.refh_sess_masks.argtypes = [C.c_void_p, c_ubyte_p, c_ubyte_p]
        L.refh_sess_end_pass1.argtypes = [C.c_void_p]
        L.refh_sess_iterate.restype = C.c_char_p
        L.refh_sess_iterate.argtypes = [C.c_void_p, C.c_int, c_int_p]
        for f in ("refh_sess_iter_num", "refh_sess_ref_len", "refh_sess_wrap_len", "refh_sess_num_aln"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.refh_sess_ref_seq.restype = C.c_char_p
        L.refh_sess_ref_seq.argtypes = [C.c_void_p]
        L.refh_sess_write_ma.argtypes = [C.c_void_p, C.c_char_p]
        L.refh_sess_gaps.argtypes = [C.c_void_p, c_int_p]
        L.refh_sess_num_fs.restype = C.c_longlong
        L.refh_sess_num_fs.argtypes = [C.c_void_p]
        L.refh_sess_fs.argtypes = [C.c_void_p, C.c_longlong, c_int_p, C.c_char_p]
        L.refh_sess_fs_id.argtypes = [C.c_void_p, C.c_longlong, C.c_char_p]
        L.refh_sess_fs_id.restype = None
        L.refh_sess_set_repeat.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.refh_sess_set_repeat.restype = None
        L.refh_sess_aln.argtypes = [C.c_void_p, C.c_int, c_int_p] + [C.c_char_p] * 4
        L.refh_sess_column.argtypes = [C.c_void_p, C.c_int, c_int_p]
        L.refh_find_consensus.argtypes = [c_int_p, C.c_int]

    def read_pssm(self, path):
        sm = np.zeros(775, np.int32)
        self.lib.refh_read_pssm(_b(path), _ip(sm))
        return sm

    def flat_pssm(self):
        sm = np.zeros(775, np.int32)
        self.lib.refh_flat_pssm(_ip(sm))
        return sm

    def revcom_pssm(self, sm):
        out = np.zeros(775, np.int32)
        self.lib.refh_revcom_pssm(_ip(np.ascontiguousarray(sm, np.int32)), _ip(out))
        return out

    def sm_depth(self, row, length):
        return self.lib.refh_find_sm_depth(row, length)

    def set_hp(self, on):
        """mia -h for the alignments / sessions made from now on (refh_set_hp)"""
        self.lib.refh_set_hp(int(bool(on)))

    def align(self, seq1, seq2, sm, sg5=1, mask=None, matrices=False, hp=0):
        self.set_hp(hp)
        seq1, seq2 = _b(seq1), _b(seq2)
        out5 = np.zeros(5, np.int32)
        rg, fg = C.create_string_buffer(520), C.create_string_buffer(520)
        sm = np.ascontiguousarray(sm, np.int32)
        m = None if mask is None else _bp(np.ascontiguousarray(mask, np.uint8))
        S = T = None
        if matrices:
            S = np.zeros((len(seq2), len(seq1)), np.int32)
            T = np.zeros((len(seq2), len(seq1)), np.int32)
        self.lib.refh_align(seq1, len(seq1), seq2, len(seq2), m, _ip(sm), sg5, _ip(out5), rg, fg,
                            None if S is None else _ip(S), None if T is None else _ip(T))
        res = self._res(out5, rg, fg)
        if matrices:
            res["S"], res["T"] = S, T
        return res

    def time_realign(self, ref, reads, off, win_start, win_len, rc, smf, smr):
        """Wall seconds for the reference's per-read realign sequence over a batch (1 thread)."""
        cells, ck = C.c_longlong(), C.c_longlong()
        off = np.ascontiguousarray(off, np.int64)
        ws, wl, rc = (np.ascontiguousarray(a, np.int32) for a in (win_start, win_len, rc))
        t = self.lib.refh_time_realign(_b(ref), len(off) - 1, reads.ctypes.data_as(C.c_char_p),
                                       off.ctypes.data_as(C.POINTER(C.c_longlong)), _ip(ws), _ip(wl), _ip(rc),
                                       _ip(np.ascontiguousarray(smf, np.int32)), _ip(np.ascontiguousarray(smr, np.int32)),
                                       C.byref(cells), C.byref(ck))
        return t, cells.value, ck.value

    def kmer_new(self, seq, k, soft_mask=0):
        seq = _b(seq)
        return self.lib.refh_kmer_new(seq, len(seq), k, soft_mask)

    def kmer_free(self, t, k):
        self.lib.refh_kmer_free(t, k)

    def kmer_lookup(self, t, inx):
        out = np.zeros(128, np.uint32)
        n = self.lib.refh_kmer_lookup(t, inx, out.ctypes.data_as(C.c_void_p))
        return out[:n].copy()

    def kmer_filter(self, ft, rt, k, read, len1):
        read = _b(read)
        mf, mr = np.ones(len1, np.uint8), np.ones(len1, np.uint8)
        hits = self.lib.refh_kmer_filter(ft, rt, k, read, len(read), len1, _bp(mf), _bp(mr))
        return hits, mf, mr

    # session = mia_main.c main() with in-memory reads
    def sess_new(self, ref_fasta_path, circular, sm, k=0, soft_mask=0, distant_ref=0, cons_code=1, hp=0):
        self.set_hp(hp)
        s = self.lib.refh_sess_new(_b(ref_fasta_path), circular, k, soft_mask, _ip(np.ascontiguousarray(sm, np.int32)),
                                   distant_ref, cons_code)
        if not s:
            raise RuntimeError("reference session failed to read " + ref_fasta_path)
        return s

    def sess_set_repeat(self, s, repeat_filt=1, just_outer_coords=1):
        self.lib.refh_sess_set_repeat(s, int(repeat_filt), int(just_outer_coords))

    def sess_pass1(self, s, rid, read, want_masks=False, qual_sum=0):
        self.lib.refh_sess_set_next_qual_sum(s, int(qual_sum))
        out = np.zeros(18, np.int32)
        bufs = [C.create_string_buffer(520) for _ in range(4)]
        self.lib.refh_sess_pass1(s, _b(rid), _b(read), _ip(out), *bufs)
        mf = mr = None
        if want_masks:
            n = self.lib.refh_sess_wrap_len(s)
            mf, mr = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
            self.lib.refh_sess_masks(s, _bp(mf), _bp(mr))
        return _p1_dict(out, bufs, mf, mr)

    def sess_end_pass1(self, s):
        self.lib.refh_sess_end_pass1(s)

    def sess_iterate(self, s, sort=1):
        conv = C.c_int()
        cons = self.lib.refh_sess_iterate(s, sort, C.byref(conv))
        return cons.decode(), bool(conv.value)

    def sess_ref(self, s):
        return self.lib.refh_sess_ref_seq(s).decode()[: self.lib.refh_sess_ref_len(s)]

    def sess_gaps(self, s):
        g = np.zeros(self.lib.refh_sess_wrap_len(s) + 1, np.int32)
        self.lib.refh_sess_gaps(s, _ip(g))
        return g

    def sess_reads(self, s):
        res = []
        for i in range(self.lib.refh_sess_num_fs(s)):
            o = np.zeros(8, np.int32)
            seq = C.create_string_buffer(260)
            self.lib.refh_sess_fs(s, i, _ip(o), seq)
            rid = C.create_string_buffer(128)
            self.lib.refh_sess_fs_id(s, C.c_longlong(i), rid)
            res.append(dict(seq_len=int(o[0]), score=int(o[1]), rc=int(o[2]), as_=int(o[3]), ae=int(o[4]),
                            strand_known=int(o[5]), unique_best=int(o[6]), has_back=int(o[7]), seq=seq.value.decode(),
                            id=rid.value.decode()))
        return res

    def sess_slots(self, s):
        res = []
        for i in range(self.lib.refh_sess_num_aln(s)):
            o = np.zeros(7, np.int32)
            rid, seq, smp, ins = (C.create_string_buffer(520), C.create_string_buffer(520), C.create_string_buffer(520),
                                  C.create_string_buffer(4096))
            self.lib.refh_sess_aln(s, i, _ip(o), rid, seq, smp, ins)
            res.append(dict(start=int(o[0]), end=int(o[1]), score=int(o[2]), revcom=int(o[3]), dropped=int(o[4]),
                            segment=chr(o[5]), id=rid.value.decode(), seq=seq.value.decode(), smp=smp.value.decode(),
                            ins=ins.value.decode()))
        return res

    def sess_column(self, s, pos):
        o = np.zeros(10, np.int32)
        ch = self.lib.refh_sess_column(s, pos, _ip(o))
        return chr(ch), o

    def sess_write_ma(self, s, path):
        return self.lib.refh_sess_write_ma(s, _b(path))

    def trim(self, read, adapter):
        """the reference's trim_frag -> dict(trimmed, trim_point, score, abr, abc, aer)"""
        out = np.zeros(6, np.int32)
        f = self.lib.refh_trim
        f.argtypes = [C.c_char_p, C.c_char_p, c_int_p]
        f.restype = None
        f(_b(read), _b(adapter), _ip(out))
        return dict(zip(("trimmed", "trim_point", "score", "abr", "abc", "aer"), map(int, out)))

    def repeat_filter(self, rc, as_, ae, key4, trimmed=None, just_outer_coords=1, tolerance=0, use_qscore=0):
        """the reference's sort_fsdb[_qscore] + set_uniq_in_fsdb: (order int64[n], unique uint8[n] by input index)"""
        n = len(rc)
        rc = np.ascontiguousarray(rc, np.uint8); as_ = np.ascontiguousarray(as_, np.int32); ae = np.ascontiguousarray(ae, np.int32)
        key4 = np.ascontiguousarray(key4, np.int32)
        tr = None if trimmed is None else np.ascontiguousarray(trimmed, np.uint8)
        order, uniq = np.zeros(n, np.int64), np.zeros(n, np.uint8)
        f = self.lib.refh_repeat_filter
        f.argtypes = [C.c_longlong] + [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        f.restype = None
        f(n, rc.ctypes.data, as_.ctypes.data, ae.ctypes.data, None if use_qscore else key4.ctypes.data,
          key4.ctypes.data if use_qscore else None, None if tr is None else tr.ctypes.data, int(use_qscore), int(just_outer_coords),
          int(tolerance), order.ctypes.data, uniq.ctypes.data)
        return order, uniq

    def find_consensus(self, counts10, cons_code=1):
        c = np.ascontiguousarray(counts10, np.int32)
        return chr(self.lib.refh_find_consensus(_ip(c), cons_code))
