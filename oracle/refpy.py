"""ctypes access to the parity checkers under oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

  RefLib     : oracle/_ref/libhibag_ref.so    = the unmodified reference sources + ref_driver.cpp
  OracleLib  : oracle/_ref/libhibag_oracle.so = this repo's C restatement (hibag_oracle.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under hibag_b200/ does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libhibag_ref.so")
ORACLE_SO = os.path.join(HERE, "_ref", "libhibag_oracle.so")

# reference inst/include/LibHLA_ext.h:261-299 (32 bytes) and :311-352 (48 bytes)
HAPLO_DT = np.dtype([("packed", "<u8", (2,)), ("freq", "<f8"), ("freq_f32", "<f4"),
                     ("hla", "<i4")], align=True)
GENO_DT = np.dtype([("s1", "<u8", (2,)), ("s2", "<u8", (2,)), ("boot", "<i4"),
                    ("a1", "<i4"), ("a2", "<i4"), ("tmp", "<i4")], align=True)
assert HAPLO_DT.itemsize == 32 and GENO_DT.itemsize == 48

NA_INTEGER = -2147483648


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_haplo(packed, freq, hla):
    """THaplotype[] with aux.a2 filled as SetHaploAux_GPU does (reference LibHLA.cpp:565)."""
    h = np.zeros(len(freq), dtype=HAPLO_DT)
    h["packed"] = np.asarray(packed, dtype=np.uint64).reshape(-1, 2)
    h["freq"] = freq
    h["freq_f32"] = np.asarray(freq, dtype=np.float32)
    h["hla"] = hla
    return h


def pack_geno(geno_rows, boot=None, a1=None, a2=None):
    """TGenotype[] from int rows [n][n_snp<=128]; 0/1/2, anything else missing
    (reference LibHLA.cpp:609-622 encoding; unused bits = missing, :671-673)."""
    g = np.asarray(geno_rows)
    n, m = g.shape
    assert m <= 128
    out = np.zeros(n, dtype=GENO_DT)
    s1 = np.zeros((n, 2), dtype=np.uint64)
    s2 = np.full((n, 2), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    for j in range(m):
        w, b = j >> 6, np.uint64(j & 63)
        col = g[:, j]
        one = np.uint64(1) << b
        b1 = (col == 1) | (col == 2)
        b2 = ~((col == 0) | (col == 1))          # 2 or missing -> 1
        s1[:, w] |= np.where(b1, one, np.uint64(0))
        s2[:, w] &= ~np.where(b2, np.uint64(0), one)
    out["s1"], out["s2"] = s1, s2
    if boot is not None:
        out["boot"] = boot
    if a1 is not None:
        lo, hi = np.minimum(a1, a2), np.maximum(a1, a2)
        out["a1"], out["a2"] = lo, hi
    return out


class RefLib:
    """The compiled, unmodified reference (see oracle/ref_driver.cpp)."""

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(path)
        L.ref_last_error.restype = C.c_char_p
        L.ref_cpu_info.restype = C.c_char_p
        L.ref_unif_rand.restype = C.c_double
        L.ref_rng_draws.restype = C.c_ulonglong
        L.ref_exp_log_min_rare_freq.restype = C.POINTER(C.c_double)
        L.ref_model_new.restype = C.c_void_p
        L.ref_model_free.argtypes = [C.c_void_p]
        L.ref_set_gpu_procs.argtypes = [C.c_void_p]
        L.ref_set_interrupt.argtypes = [C.c_double, C.c_longlong]
        L.ref_interrupt_checks.restype = C.c_longlong
        L.ref_model_init_training.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                              C.c_void_p, C.c_void_p]
        L.ref_model_init_predict.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_model_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_longlong, C.c_int]
        L.ref_model_num_classifiers.argtypes = [C.c_void_p]
        L.ref_model_clear.argtypes = [C.c_void_p]
        L.ref_classifier_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.ref_classifier_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.ref_model_add_classifier.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_double]
        L.ref_model_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        for f in (L.ref_best_guess, L.ref_post_prob, L.ref_post_prob2, L.ref_acc_oob, L.ref_acc_ib):
            f.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                          C.c_void_p] + ([C.c_void_p] if f in (L.ref_best_guess, L.ref_post_prob2) else [])
        L.ref_int_to_snp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        assert L.ref_sizeof_haplotype() == 32 and L.ref_sizeof_genotype() == 48

    def _chk(self, rc, allow_interrupt=False):
        if rc == 1 and allow_interrupt:
            return 1
        if rc != 0:
            raise RuntimeError("reference: " + self.lib.ref_last_error().decode())
        return 0

    # -- global state ------------------------------------------------------------------------
    def set_seed(self, seed):
        self.lib.ref_set_seed(C.c_uint(seed))

    def set_target(self, name):
        self._chk(self.lib.ref_set_target(name.encode()))
        return self.lib.ref_cpu_info().decode()

    def table(self):
        return np.ctypeslib.as_array(self.lib.ref_exp_log_min_rare_freq(), shape=(257,)).copy()

    def set_gpu_procs(self, ptr):
        self.lib.ref_set_gpu_procs(C.c_void_p(ptr) if ptr else None)

    def set_interrupt(self, seconds=0.0, after_checks=-1):
        self.lib.ref_set_interrupt(seconds, after_checks)

    # -- scoring kernels ------------------------------------------------------------------------
    def best_guess(self, haplo, n_hla, n_snp, geno, target="base"):
        a1 = np.zeros(len(geno), dtype=np.int32); a2 = np.zeros(len(geno), dtype=np.int32)
        self._chk(self.lib.ref_best_guess(target.encode(), _p(haplo), len(haplo), n_hla, n_snp,
                                          _p(geno), len(geno), _p(a1), _p(a2)))
        return a1, a2

    def post_prob(self, haplo, n_hla, n_snp, geno, target="base"):
        out = np.zeros(len(geno))
        self._chk(self.lib.ref_post_prob(target.encode(), _p(haplo), len(haplo), n_hla, n_snp,
                                         _p(geno), len(geno), _p(out)))
        return out

    def post_prob2(self, haplo, n_hla, n_snp, geno, target="base"):
        nc = n_hla * (n_hla + 1) // 2
        prob = np.zeros((len(geno), nc)); s = np.zeros(len(geno))
        self._chk(self.lib.ref_post_prob2(target.encode(), _p(haplo), len(haplo), n_hla, n_snp,
                                          _p(geno), len(geno), _p(prob), _p(s)))
        return prob, s

    def acc_oob(self, haplo, n_hla, n_snp, geno, target="base"):
        v = C.c_int(0)
        self._chk(self.lib.ref_acc_oob(target.encode(), _p(haplo), len(haplo), n_hla, n_snp,
                                       _p(geno), len(geno), C.byref(v)))
        return v.value

    def acc_ib(self, haplo, n_hla, n_snp, geno, target="base"):
        v = C.c_double(0)
        self._chk(self.lib.ref_acc_ib(target.encode(), _p(haplo), len(haplo), n_hla, n_snp,
                                      _p(geno), len(geno), C.byref(v)))
        return v.value

    def int_to_snp(self, geno_row, index):
        out = np.zeros(1, dtype=GENO_DT)
        g = np.ascontiguousarray(geno_row, dtype=np.int32)
        ix = np.ascontiguousarray(index, dtype=np.int32)
        self._chk(self.lib.ref_int_to_snp(_p(out), len(ix), _p(g), _p(ix)))
        return out

    def hamming(self, geno1, h1, h2, n_snp):
        return self.lib.ref_hamming(_p(geno1), _p(h1), _p(h2), n_snp)

    def prep_haplo_match(self, geno1, haplo, st1, m1, st2, m2, n_snp):
        """reference _PrepHaploMatch_def on two allele ranges; returns [(i1, i2), ...]"""
        cap = max(m1 * max(m1, m2), 1)
        a = np.zeros(cap, dtype=np.int32); b = np.zeros(cap, dtype=np.int32)
        self.lib.ref_prep_haplo_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                  C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        n = self.lib.ref_prep_haplo_match(_p(geno1), _p(haplo), st1, m1, st2, m2, n_snp, _p(a), _p(b), cap)
        return list(zip(a[:n].tolist(), b[:n].tolist()))

    def new_model(self):
        return RefModel(self)


class RefModel:
    """CAttrBag_Model of the compiled reference."""

    def __init__(self, ref):
        self.ref, self.L = ref, ref.lib
        self.h = C.c_void_p(self.L.ref_model_new())
        self.n_snp = self.n_samp = self.n_hla = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_model_free(self.h)
            self.h = None

    def init_training(self, geno, h1, h2, n_hla):
        g = np.ascontiguousarray(geno, dtype=np.int32)
        self.n_samp, self.n_snp = g.shape
        self.n_hla = n_hla
        a = np.ascontiguousarray(h1, dtype=np.int32); b = np.ascontiguousarray(h2, dtype=np.int32)
        self.ref._chk(self.L.ref_model_init_training(self.h, self.n_snp, self.n_samp, _p(g), n_hla,
                                                     _p(a), _p(b)))

    def init_predict(self, n_snp, n_samp, n_hla):
        self.n_snp, self.n_samp, self.n_hla = n_snp, n_samp, n_hla
        self.ref._chk(self.L.ref_model_init_predict(self.h, n_snp, n_samp, n_hla))

    def build(self, nclassifier, mtry, prune=True, verbose=0, reseed_base=-1, first_index=0,
              allow_interrupt=False):
        return self.ref._chk(self.L.ref_model_build(self.h, nclassifier, mtry, int(prune), verbose,
                                                    reseed_base, first_index), allow_interrupt)

    def num_classifiers(self):
        return self.L.ref_model_num_classifiers(self.h)

    def clear(self):
        self.L.ref_model_clear(self.h)

    def classifier(self, k):
        ns, nh, acc = C.c_int(), C.c_int(), C.c_double()
        assert self.L.ref_classifier_info(self.h, k, C.byref(ns), C.byref(nh), C.byref(acc)) == 0
        snpidx = np.zeros(ns.value, dtype=np.int32)
        boot = np.zeros(self.n_samp, dtype=np.int32)
        freq = np.zeros(nh.value); hla = np.zeros(nh.value, dtype=np.int32)
        packed = np.zeros((nh.value, 2), dtype=np.uint64)
        self.L.ref_classifier_get(self.h, k, _p(snpidx), _p(boot), _p(freq), _p(hla), _p(packed))
        return dict(snpidx=snpidx, samp_num=boot, freq=freq, hla=hla, packed=packed,
                    oob_acc=acc.value)

    def add_classifier(self, snpidx, freq, hla, packed, samp_num=None, acc=0.0):
        s = np.ascontiguousarray(snpidx, dtype=np.int32)
        f = np.ascontiguousarray(freq, dtype=np.float64)
        h = np.ascontiguousarray(hla, dtype=np.int32)
        p = np.ascontiguousarray(packed, dtype=np.uint64)
        sn = None if samp_num is None else np.ascontiguousarray(samp_num, dtype=np.int32)
        self.ref._chk(self.L.ref_model_add_classifier(self.h, len(s), _p(s), _p(sn), len(f), _p(f),
                                                      _p(h), _p(p), acc))

    def predict(self, geno, vote=1, want_prob=True, want_dosage=True):
        g = np.ascontiguousarray(geno, dtype=np.int32)
        n = g.shape[0]
        assert g.shape[1] == self.n_snp
        nc = self.n_hla * (self.n_hla + 1) // 2
        h1 = np.zeros(n, dtype=np.int32); h2 = np.zeros(n, dtype=np.int32)
        mp = np.zeros(n); mt = np.zeros(n)
        ds = np.zeros((n, self.n_hla)) if want_dosage else None
        pr = np.zeros((n, nc)) if want_prob else None
        self.ref._chk(self.L.ref_model_predict(self.h, _p(g), n, vote, _p(h1), _p(h2), _p(mp),
                                               _p(mt), _p(ds), _p(pr)))
        return dict(h1=h1, h2=h2, prob=mp, matching=mt, dosage=ds, postprob=pr)


class OracleLib:
    """This repo's C restatement of the scoring path (oracle/hibag_oracle.c)."""

    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle oracle`)")
        L = self.lib = C.CDLL(path)
        common = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.hibag_oracle_table.argtypes = [C.c_void_p]
        L.hibag_oracle_hamming.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.hibag_oracle_best_guess.argtypes = common + [C.c_void_p, C.c_void_p]
        L.hibag_oracle_post_prob.argtypes = common + [C.c_void_p]
        L.hibag_oracle_post_prob2.argtypes = common + [C.c_void_p, C.c_void_p]
        L.hibag_oracle_acc_oob.argtypes = common
        L.hibag_oracle_acc_ib.argtypes = common
        L.hibag_oracle_acc_ib.restype = C.c_double
        L.hibag_oracle_predict_avg.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hibag_oracle_best_guess_cells.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hibag_oracle_dosage.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hibag_oracle_int_to_snp.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.hibag_oracle_classifier_weights.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p, C.c_void_p]

    def table(self):
        t = np.zeros(257)
        self.lib.hibag_oracle_table(_p(t))
        return t

    def bed_decode(self, bed_bytes, n_samp, n_snp, snp_flag=None):
        """reference HIBAG_ConvBED restated: int32 [n_samp][n_kept], NA = INT32_MIN"""
        raw = np.ascontiguousarray(np.frombuffer(bed_bytes, dtype=np.uint8) if not isinstance(bed_bytes, np.ndarray)
                                   else bed_bytes, dtype=np.uint8)
        flag = None if snp_flag is None else np.ascontiguousarray(snp_flag, dtype=np.int32)
        n_save = n_snp if flag is None else int((flag != 0).sum())
        out = np.zeros((n_samp, max(n_save, 1)), dtype=np.int32)
        L = self.lib
        L.hibag_oracle_bed_decode.restype = C.c_int
        L.hibag_oracle_bed_decode.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        rc = L.hibag_oracle_bed_decode(_p(raw), raw.size, n_samp, n_snp, _p(flag), n_save, _p(out))
        if rc != 0:
            raise RuntimeError("Invalid prefix in the PLINK BED file." if rc == -1 else "BED file too short")
        return out[:, :n_save]

    def haplomatch_records(self, haplo, len_per_hla, n_snp, geno):
        """records a build_haplomatch plugin returns: uint32 [n][2] = (in-bag index, (i2<<16)|i1)"""
        L = self.lib
        L.hibag_oracle_haplomatch_records.restype = C.c_long
        L.hibag_oracle_haplomatch_records.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                      C.c_int, C.c_void_p, C.c_long]
        lens = np.ascontiguousarray(len_per_hla, dtype=np.int64)
        cap = 1 << 16
        while True:
            out = np.zeros((cap, 2), dtype=np.uint32)
            n = L.hibag_oracle_haplomatch_records(_p(haplo), _p(lens), len(lens), n_snp, _p(geno), len(geno),
                                                  _p(out), cap)
            if n >= 0:
                return out[:n]
            cap *= 4

    def hamming(self, geno1, h1, h2, n_snp):
        return self.lib.hibag_oracle_hamming(_p(geno1), _p(h1), _p(h2), n_snp)

    def best_guess(self, haplo, n_hla, n_snp, geno):
        a1 = np.zeros(len(geno), dtype=np.int32); a2 = np.zeros(len(geno), dtype=np.int32)
        self.lib.hibag_oracle_best_guess(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno),
                                         _p(a1), _p(a2))
        return a1, a2

    def post_prob(self, haplo, n_hla, n_snp, geno):
        out = np.zeros(len(geno))
        self.lib.hibag_oracle_post_prob(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno), _p(out))
        return out

    def post_prob2(self, haplo, n_hla, n_snp, geno):
        nc = n_hla * (n_hla + 1) // 2
        prob = np.zeros((len(geno), nc)); s = np.zeros(len(geno))
        self.lib.hibag_oracle_post_prob2(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno),
                                         _p(prob), _p(s))
        return prob, s

    def acc_oob(self, haplo, n_hla, n_snp, geno):
        return self.lib.hibag_oracle_acc_oob(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno))

    def acc_ib(self, haplo, n_hla, n_snp, geno):
        return self.lib.hibag_oracle_acc_ib(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno))

    def int_to_snp(self, geno_row, index):
        out = np.zeros(1, dtype=GENO_DT)
        g = np.ascontiguousarray(geno_row, dtype=np.int32)
        ix = np.ascontiguousarray(index, dtype=np.int32)
        self.lib.hibag_oracle_int_to_snp(_p(out), len(ix), _p(g), _p(ix))
        return out

    def best_guess_cells(self, prob, n_hla):
        a1, a2 = C.c_int32(), C.c_int32()
        p = np.ascontiguousarray(prob, dtype=np.float64)
        self.lib.hibag_oracle_best_guess_cells(_p(p), n_hla, C.byref(a1), C.byref(a2))
        return a1.value, a2.value

    def dosage(self, prob, n_hla):
        p = np.ascontiguousarray(prob, dtype=np.float64)
        d = np.zeros(n_hla)
        self.lib.hibag_oracle_dosage(_p(p), n_hla, _p(d))
        return d

    def predict(self, classifiers, n_hla, n_total_snp, geno_rows):
        """Ensemble prediction of int genotype rows [n][n_total_snp] with a list of classifier
        dicts (snpidx, freq, hla, packed): the CPU branch of _PredictHLA + PredictHLA's outputs."""
        nc = n_hla * (n_hla + 1) // 2
        n_cls = len(classifiers)
        haplos = [make_haplo(c["packed"], c["freq"], c["hla"]) for c in classifiers]
        hp = (C.c_void_p * n_cls)(*[h.ctypes.data for h in haplos])
        nh = np.array([len(h) for h in haplos], dtype=np.int32)
        ns = np.array([len(c["snpidx"]) for c in classifiers], dtype=np.int32)
        idx = [np.ascontiguousarray(c["snpidx"], dtype=np.int32) for c in classifiers]
        ip = (C.c_void_p * n_cls)(*[x.ctypes.data for x in idx])
        g = np.ascontiguousarray(geno_rows, dtype=np.int32)
        n = g.shape[0]
        out = dict(h1=np.zeros(n, dtype=np.int32), h2=np.zeros(n, dtype=np.int32), prob=np.zeros(n),
                   matching=np.zeros(n), dosage=np.zeros((n, n_hla)), postprob=np.zeros((n, nc)))
        w = np.zeros(n_cls)
        packed = np.zeros(n_cls, dtype=GENO_DT)
        for s in range(n):
            row = g[s]
            self.lib.hibag_oracle_classifier_weights(n_cls, _p(ns), ip, n_total_snp, _p(row), _p(w))
            for c in range(n_cls):
                self.lib.hibag_oracle_int_to_snp(C.c_void_p(packed.ctypes.data + 48 * c), int(ns[c]),
                                                 _p(row), _p(idx[c]))
            m = C.c_double()
            pr = out["postprob"][s]
            self.lib.hibag_oracle_predict_avg(n_hla, n_cls, hp, _p(nh), _p(ns), _p(packed), _p(w),
                                              _p(pr), C.byref(m))
            out["matching"][s] = m.value
            a1, a2 = self.best_guess_cells(pr, n_hla)
            out["h1"][s], out["h2"][s] = a1, a2
            if a1 != NA_INTEGER:
                lo, hi = min(a1, a2), max(a1, a2)
                out["prob"][s] = pr[hi + lo * (2 * n_hla - lo - 1) // 2]
            out["dosage"][s] = self.dosage(pr, n_hla)
        return out
