"""Oracle-side mirror of mia_main.c main() (lines 759-976) built from the
restated pieces in oracle/mia_oracle.c.  Test infrastructure."""
import numpy as np


class OracleRun:
    def __init__(self, o, ref_raw, sm, circular=1, k=0, soft_mask=0, cons_code=1, distant_ref=0, repeat_filt=0, just_outer_coords=1, hp=0):
        self.o, self.sm, self.circular, self.cons_code = o, np.ascontiguousarray(sm, np.int32), circular, cons_code
        self.repeat_filt, self.just_outer_coords = repeat_filt, just_outer_coords
        self.smr = o.revcom_pssm(self.sm)
        self.hp = hp                  # mia -h
        self.ctx = o.ctx_new(ref_raw, circular, self.sm, with_rc=1, k=k, soft_mask=soft_mask, distant_ref=distant_ref, hp=hp)
        self.seq_len = len(ref_raw)
        self.wrap_len = o.lib.orc_ctx_wrap_len(self.ctx)
        self.cur_ref = o.ctx_seq(self.ctx)[: self.seq_len]
        self.asm = o.asm_new()
        o.asm_begin_round(self.asm, self.seq_len, self.wrap_len)
        self.fsdb = []
        self.iter = 0
        self.cons = None
        self.distant_ref = distant_ref
        self.submat_rc = 0            # which matrix a->submat was left pointing at (H6: mia_main.c:126-137 never sets it)

    def pass1(self, read, want_masks=False, qual_sum=0):
        p = self.o.pass1(self.ctx, read, want_masks)
        if p["added"]:
            seq = self.o.revcom(read) if (p["rc"] and p["strand_known"]) else read      # fsdb.c:209-227
            end = p["b_end"] if p["split"] else p["end"]
            f, b = self.o.asm_add(self.asm, p["f_ref"] + p["b_ref"], p["f_frag"] + p["b_frag"], p["start"], end, p["rc"], p["score"])
            self.fsdb.append(dict(rid=len(self.fsdb), unique_best=1, qual_sum=int(qual_sum), seq=seq, seq_len=len(read), score=p["score"], rc=p["rc"], as_=p["as_"], ae=p["ae"],
                                  strand_known=p["strand_known"], front=f, back=-1 if b is None else b))   # mia.c:1626-1642
        return p

    def _arrays(self):
        g = lambda k: np.array([f[k] for f in self.fsdb], np.int32)
        return g("front"), g("back"), g("seq_len"), g("score")

    def _repeat_filter(self):
        """-u: sort_fsdb + set_uniq_in_fsdb (mia_main.c:827-834, 883-886, 938-941): the FSDB itself is re-ordered"""
        if not self.repeat_filt or not self.fsdb:
            return None
        g = lambda k: np.array([f[k] for f in self.fsdb], np.int32)
        key4 = g("qual_sum") if self.repeat_filt == 2 else g("score")                    # -U: sort_fsdb_qscore (fsdb.c:90-180, 250-253)
        order, uniq = self.o.repeat_filter(g("rc").astype(np.uint8), g("as_"), g("ae"), key4, None, self.just_outer_coords, 0)
        for f, u in zip(self.fsdb, uniq):
            f["unique_best"] = int(u)
        self.fsdb = [self.fsdb[k] for k in order]
        return np.array([f["unique_best"] for f in self.fsdb], np.uint8)

    def _alignable(self, ref_wrapped):
        """-D: find_alignable_len of every read against the current (wrapped) reference (mia.c:460-463)"""
        if not self.distant_ref:
            return None
        return np.array([self.o.alignable_len(ref_wrapped, f["seq_len"], f["as_"], f["ae"]) for f in self.fsdb], np.int32)

    def end_pass1(self):
        fr, bk, sl, sc = self._arrays()
        self.o.asm_pop_smp(self.asm, fr, bk)
        uq = self._repeat_filter()
        fr, bk, sl, sc = self._arrays()
        self.o.asm_cull(self.asm, fr, bk, sl, sc, unique_best=uq, alignable_len=self._alignable(self.o.ctx_seq(self.ctx)))
        self.fsdb = [f for f in self.fsdb if f["score"] > 0]                             # clean_FSDB mia.c:400-406
        self.iter = 1
        self.last = self.cur_ref

    def iterate(self):
        o = self.o
        if self.cons is not None:
            self.iter += 1
            self.last = self.cons
        ref = self.last
        ctx = o.ctx_new(ref, self.circular, self.sm, with_rc=0, k=0, hp=self.hp)
        self.seq_len = len(ref)
        self.wrap_len = o.lib.orc_ctx_wrap_len(ctx)
        o.asm_begin_round(self.asm, self.seq_len, self.wrap_len)
        ref_w = o.ctx_seq(ctx)
        for f in self.fsdb:
            if self.distant_ref and not f["strand_known"] and self.iter > 1:        # mia_main.c:120-174
                a = o.align(ref_w, f["seq"], self.smr if self.submat_rc else self.sm, 1, hp=self.hp)       # whatever matrix the last read left (H6)
                if a["score"] > 2000:
                    f["strand_known"], f["rc"], f["as_"], f["ae"], f["score"] = 1, 0, a["abc"], a["aec"], a["score"]
                rcs = o.revcom(f["seq"])
                self.submat_rc = 1                                                  # mia_main.c:151
                a = o.align(ref_w, rcs, self.smr, 1, hp=self.hp)
                if a["score"] > 2000 and a["score"] > f["score"]:
                    f["strand_known"], f["rc"], f["as_"], f["ae"], f["score"], f["seq"] = 1, 1, a["abc"], a["aec"], a["score"], rcs
            if not f["strand_known"]:
                continue
            self.submat_rc = 1 if f["rc"] else 0                                    # mia_main.c:179-184
            r = o.realign(ctx, f["seq"], f["rc"], f["as_"], f["ae"])
            f["as_"], f["ae"], f["score"], f["unique_best"] = r["as_"], r["ae"], r["score"], 1          # mia_main.c:254
            fs, bs = o.asm_add(self.asm, r["ref_gapped"], r["read_gapped"], r["as_"], r["ae"], f["rc"], r["score"])
            f["front"] = fs
            if bs is not None:
                f["back"] = bs          # otherwise the old back slot id stays: mia_main.c:273-276
        al = self._alignable(ref_w)
        o.ctx_free(ctx)
        fr, bk, sl, sc = self._arrays()
        o.asm_pop_smp(self.asm, fr, bk)
        uq = self._repeat_filter()
        fr, bk, sl, sc = self._arrays()
        o.asm_cull(self.asm, fr, bk, sl, sc, unique_best=uq, alignable_len=al)
        self.cons = o.asm_consensus(self.asm, self.sm, self.smr, self.cons_code, self.seq_len)
        return self.cons, self.cons == self.last
