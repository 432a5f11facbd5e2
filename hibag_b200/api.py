"""Python mirror of the reference's user-facing interface for the scoring path, over the C ABI
of libhibag_b200.so (include/hibag_b200.h).

Names and argument meaning follow the reference's R front end where one exists:
  hlaAttrBagging(hla, snp, nclassifier, mtry, prune, ...)   reference R/HIBAG.R:48-275
  hlaPredict(model, snp, type=...)                            reference R/HIBAG.R:481-818
  hlaModelToObj / hlaModelFromObj                             reference R/HIBAG.R:1041-1178
  hlaSetKernelTarget is not mirrored: there is one target here, the B200.

The product path needs the CUDA extension: importing works anywhere (so the CPU test-suite can
check the ABI), but every compute call raises if the library or a CUDA device is missing --
there is no CPU fallback.
"""
import ctypes as C
import math
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhibag_b200.so")
NA_INTEGER = -2147483648

# binary layouts of include/hibag_b200.h (== reference LibHLA_ext.h:261-352)
HAPLO_DT = np.dtype([("packed", "<u8", (2,)), ("freq", "<f8"), ("freq_f32", "<f4"),
                     ("hla", "<i4")], align=True)
GENO_DT = np.dtype([("s1", "<u8", (2,)), ("s2", "<u8", (2,)), ("boot", "<i4"),
                    ("a1", "<i4"), ("a2", "<i4"), ("tmp", "<i4")], align=True)


class TrainOpts(C.Structure):
    _fields_ = [("nclassifier", C.c_int), ("mtry", C.c_int), ("prune", C.c_int),
                ("n_threads", C.c_int), ("seed", C.c_int64), ("per_classifier_seed", C.c_int),
                ("first_index", C.c_int), ("index_stride", C.c_int),
                ("use_legacy_hooks", C.c_int), ("verbose", C.c_int), ("n_concurrent", C.c_int),
                ("em_on_device", C.c_int), ("no_screening", C.c_int)]


class TrainStats(C.Structure):
    _fields_ = [("seconds_total", C.c_double), ("seconds_em", C.c_double),
                ("seconds_gpu_wait", C.c_double), ("gpu_kernel_ms", C.c_double),
                ("pair_evals", C.c_uint64), ("popc32_issued", C.c_uint64),
                ("n_oob_evals", C.c_uint64), ("n_ib_evals", C.c_uint64), ("n_em", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("cell_kernel_ms", C.c_double),
                ("cell_kernel_launches", C.c_uint64), ("seconds_prepare", C.c_double),
                ("seconds_phase_oob", C.c_double), ("seconds_phase_ib", C.c_double),
                ("em_kernel_ms", C.c_double), ("n_em_host_fallback", C.c_uint64),
                ("pair_evals_nominal", C.c_uint64), ("n_screen_fallback", C.c_uint64),
                ("gather_kernel_ms", C.c_double), ("gather_kernel_launches", C.c_uint64),
                ("gather_ib_kernel_ms", C.c_double), ("gather_ib_launches", C.c_uint64),
                ("gather_ib_popc32", C.c_uint64), ("em_iterations", C.c_uint64),
                ("em_chain_adds", C.c_uint64), ("em_pair_updates", C.c_uint64)]


class PredictOut(C.Structure):
    _fields_ = [("h1", C.c_void_p), ("h2", C.c_void_p), ("max_prob", C.c_void_p),
                ("matching", C.c_void_p), ("dosage", C.c_void_p), ("post_prob", C.c_void_p)]


class PredictStats(C.Structure):
    _fields_ = [("gpu_kernel_ms", C.c_double), ("cell_kernel_ms", C.c_double),
                ("pair_evals", C.c_uint64), ("popc32_issued", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("cell_kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("pair_evals_nominal", C.c_uint64), ("positions_scored", C.c_uint64),
                ("positions_total", C.c_uint64)]


# every symbol include/hibag_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "hibag_b200_version", "hibag_b200_last_error", "hibag_b200_device_count",
    "hibag_b200_set_device", "hibag_b200_device_info", "hibag_b200_get_procs",
    "hibag_b200_best_guess", "hibag_b200_post_prob", "hibag_b200_post_prob2",
    "hibag_b200_model_new", "hibag_b200_model_free", "hibag_b200_model_set_training",
    "hibag_b200_model_train", "hibag_b200_model_train_stats", "hibag_b200_model_train_trace",
    "hibag_b200_model_num_classifiers", "hibag_b200_model_clear",
    "hibag_b200_model_classifier_info", "hibag_b200_model_classifier_get",
    "hibag_b200_model_classifier_samp_num_len", "hibag_b200_trim_cache", "hibag_b200_sm_time",
    "hibag_b200_model_add_classifier", "hibag_b200_model_predict",
    "hibag_b200_model_predict_device", "hibag_b200_model_predict_stats",
    "hibag_b200_model_predict_partial_device", "hibag_b200_predict_finalize_device",
    "hibag_b200_model_snp_weights", "hibag_b200_pipe_peak",
    "hibag_b200_host_unif_rand", "hibag_b200_host_build_tasks", "hibag_b200_host_screen_constants",
    "hibag_b200_bed_decode", "hibag_b200_bed_decode_device",
    "hibag_b200_get_procs_ex", "hibag_b200_haplomatch", "hibag_b200_free",
    "hibag_b200_host_alloc", "hibag_b200_host_free",
]

_lib = None


def lib():
    """Load libhibag_b200.so (built in-tree by `make -C hibag_b200/csrc` / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    # Training keeps many classifiers in flight, each with its own streams; with the default of 8
    # hardware connections unrelated streams share a queue and wait for each other's kernels
    # (read by the driver when the CUDA context is created; harmless if that has happened)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("hibag_b200: %s is missing -- build it with __graft_entry__.build(); "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.hibag_b200_version.restype = C.c_char_p
    L.hibag_b200_last_error.restype = C.c_char_p
    L.hibag_b200_get_procs.restype = C.c_void_p
    L.hibag_b200_get_procs_ex.restype = C.c_void_p
    L.hibag_b200_get_procs_ex.argtypes = [C.c_int]
    L.hibag_b200_haplomatch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.hibag_b200_free.argtypes = [C.c_void_p]
    L.hibag_b200_device_info.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    for f in (L.hibag_b200_best_guess,):
        f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.hibag_b200_post_prob.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    L.hibag_b200_post_prob2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_void_p]
    L.hibag_b200_model_new.restype = C.c_void_p
    L.hibag_b200_model_new.argtypes = [C.c_int, C.c_int]
    L.hibag_b200_model_free.argtypes = [C.c_void_p]
    L.hibag_b200_model_set_training.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hibag_b200_model_train.argtypes = [C.c_void_p, C.POINTER(TrainOpts)]
    L.hibag_b200_model_train_stats.argtypes = [C.c_void_p, C.POINTER(TrainStats)]
    L.hibag_b200_model_train_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.hibag_b200_model_num_classifiers.argtypes = [C.c_void_p]
    L.hibag_b200_model_clear.argtypes = [C.c_void_p]
    L.hibag_b200_model_classifier_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                                   C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.hibag_b200_model_classifier_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    L.hibag_b200_model_add_classifier.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.hibag_b200_model_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(PredictOut)]
    L.hibag_b200_model_predict_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(PredictOut),
                                                  C.c_void_p, C.c_int]
    L.hibag_b200_model_predict_stats.argtypes = [C.c_void_p, C.POINTER(PredictStats)]
    L.hibag_b200_model_predict_partial_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                          C.c_void_p, C.c_void_p, C.c_int]
    L.hibag_b200_predict_finalize_device.argtypes = [C.c_int, C.c_int, C.c_void_p, C.POINTER(PredictOut),
                                                     C.c_void_p, C.c_int]
    L.hibag_b200_model_snp_weights.argtypes = [C.c_void_p, C.c_void_p]
    L.hibag_b200_host_unif_rand.argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    L.hibag_b200_host_build_tasks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    L.hibag_b200_bed_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int), C.POINTER(C.c_double)]
    L.hibag_b200_bed_decode_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_void_p]
    L.hibag_b200_pipe_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hibag_b200_host_alloc.restype = C.c_void_p
    L.hibag_b200_host_alloc.argtypes = [C.c_size_t]
    L.hibag_b200_host_free.argtypes = [C.c_void_p]
    _lib = L
    return L


def _chk(rc):
    if rc != 0:
        raise RuntimeError("hibag_b200: " + lib().hibag_b200_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count():
    return lib().hibag_b200_device_count()


def set_device(i):
    _chk(lib().hibag_b200_set_device(i))


SM_TIME_CLASSES = ["gather_oob", "gather_ib", "em", "screen_bound", "screen_need", "screen_tasks", "reduce_oob",
                   "reduce_ib", "cell_pass", "em_prep", "em_cta_cycles", "screen_dedup"]


def sm_time(reset=False):
    """held SM-time per kernel class in SM-cycles (see hibag_b200_sm_time); synchronises the device"""
    out = np.zeros(16, dtype=np.uint64)
    _chk(lib().hibag_b200_sm_time(_p(out), int(reset)))
    return {k: float(out[i]) / 1024.0 for i, k in enumerate(SM_TIME_CLASSES)}


def trim_cache():
    """release the library's cached device / pinned blocks; returns the bytes given back"""
    lib().hibag_b200_trim_cache.restype = C.c_size_t
    return int(lib().hibag_b200_trim_cache())


def pinned_empty(shape, dtype=np.float64):
    """numpy array over a page-locked block of the library's cache (hibag_b200_host_alloc); the
    block goes back to the cache when the array (and every view of it) is garbage-collected"""
    import weakref
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = lib().hibag_b200_host_alloc(max(n, 1))
    if not p:
        raise RuntimeError(lib().hibag_b200_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    weakref.finalize(buf, lib().hibag_b200_host_free, C.c_void_p(p))
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)


def device_info():
    name = C.create_string_buffer(128)
    sm, khz = C.c_int(), C.c_int()
    _chk(lib().hibag_b200_device_info(name, 128, C.byref(sm), C.byref(khz)))
    return dict(name=name.value.decode(), sm_count=sm.value, clock_khz=khz.value)


def get_procs(with_haplomatch=False):
    """Address of the TypeGPUExtProc-compatible hook struct (reference LibHLA_ext.h:358-388);
    with_haplomatch also fills the optional build_haplomatch hook."""
    if with_haplomatch:
        return lib().hibag_b200_get_procs_ex(1)
    return lib().hibag_b200_get_procs()


def haplomatch(haplo, len_per_hla, n_snp, geno):
    """GPU body of the build_haplomatch hook on host arrays: uint32 [n][2] records
    (in-bag index, (i2 << 16) | i1) for every genotype with boot > 0."""
    lens = np.ascontiguousarray(len_per_hla, dtype=np.uint64)
    buf, n = C.c_void_p(), C.c_size_t()
    _chk(lib().hibag_b200_haplomatch(_p(haplo), _p(lens), len(lens), n_snp, _p(geno), len(geno),
                                     C.byref(buf), C.byref(n)))
    try:
        arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint32)), shape=(n.value,)).copy()
    finally:
        lib().hibag_b200_free(buf)
    assert arr[0] == n.value - 1
    return arr[1:].reshape(-1, 2)


def pipe_peak(which):
    ops, ms = C.c_double(), C.c_double()
    _chk(lib().hibag_b200_pipe_peak(which, C.byref(ops), C.byref(ms)))
    return ops.value, ms.value


def host_unif_rand(seed, n):
    out = np.zeros(n)
    _chk(lib().hibag_b200_host_unif_rand(seed, n, _p(out)))
    return out


def host_build_tasks(haplo, n_hla, n_snp, target_chunks=512):
    nc = n_hla * (n_hla + 1) // 2
    cells = np.zeros((nc, 8), dtype=np.int32); chunks = np.zeros((nc, 2), dtype=np.int32)
    n, pairs = C.c_int(), C.c_uint64()
    _chk(lib().hibag_b200_host_build_tasks(_p(haplo), len(haplo), n_hla, n_snp, target_chunks, _p(cells),
                                           _p(chunks), C.byref(n), C.byref(pairs)))
    return cells, chunks[:n.value], pairs.value


def host_screen_constants():
    """(T[257], T'[257], K) of the exact screening"""
    t = np.zeros(257); f = np.zeros(257); k = C.c_double()
    _chk(lib().hibag_b200_host_screen_constants(_p(t), _p(f), C.byref(k)))
    return t, f, k.value


# ---- kernel-level batched scoring -----------------------------------------------------------

def best_guess(haplo, n_hla, n_snp, geno):
    a1 = np.zeros(len(geno), dtype=np.int32); a2 = np.zeros(len(geno), dtype=np.int32)
    _chk(lib().hibag_b200_best_guess(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno), _p(a1), _p(a2)))
    return a1, a2


def post_prob(haplo, n_hla, n_snp, geno):
    out = np.zeros(len(geno))
    _chk(lib().hibag_b200_post_prob(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno), _p(out)))
    return out


def post_prob2(haplo, n_hla, n_snp, geno):
    nc = n_hla * (n_hla + 1) // 2
    prob = np.zeros((len(geno), nc)); s = np.zeros(len(geno))
    _chk(lib().hibag_b200_post_prob2(_p(haplo), len(haplo), n_hla, n_snp, _p(geno), len(geno), _p(prob), _p(s)))
    return prob, s


# ---- model ----------------------------------------------------------------------------------

class HLAModel:
    """An attribute-bagging model (reference class "hlaAttrBagClass", R/HIBAG.R:236-247)."""

    def __init__(self, n_snp, n_hla, hla_allele=None, snp_id=None):
        self.n_snp, self.n_hla = int(n_snp), int(n_hla)
        self.hla_allele = list(hla_allele) if hla_allele is not None else [str(i) for i in range(n_hla)]
        self.snp_id = list(snp_id) if snp_id is not None else None
        self.n_samp = 0
        self._h = C.c_void_p(lib().hibag_b200_model_new(self.n_snp, self.n_hla))
        if not self._h:
            raise RuntimeError("hibag_b200: " + lib().hibag_b200_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.hibag_b200_model_free(self._h)
            self._h = None

    @property
    def n_cells(self):
        return self.n_hla * (self.n_hla + 1) // 2

    def set_training(self, geno, h1, h2):
        g = np.ascontiguousarray(geno, dtype=np.int8)
        assert g.ndim == 2 and g.shape[1] == self.n_snp
        self.n_samp = g.shape[0]
        a = np.ascontiguousarray(h1, dtype=np.int32); b = np.ascontiguousarray(h2, dtype=np.int32)
        _chk(lib().hibag_b200_model_set_training(self._h, self.n_samp, _p(g), _p(a), _p(b)))

    def train(self, nclassifier, mtry, prune=True, seed=100, n_threads=0, per_classifier_seed=False,
              first_index=0, index_stride=1, use_legacy_hooks=False, verbose=0, n_concurrent=1,
              em_on_device=True, screening=True):
        o = TrainOpts(nclassifier, mtry, int(prune), n_threads, seed, int(per_classifier_seed),
                      first_index, index_stride, int(use_legacy_hooks), verbose, int(n_concurrent),
                      int(em_on_device), int(not screening))
        _chk(lib().hibag_b200_model_train(self._h, C.byref(o)))

    def train_stats(self):
        s = TrainStats()
        _chk(lib().hibag_b200_model_train_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in TrainStats._fields_}

    def train_trace(self):
        """rows {classifier, accepted SNPs (-1 = classifier finished), cumulative pair evaluations,
        cumulative candidate EM runs}"""
        n = lib().hibag_b200_model_train_trace(self._h, None, 0)
        out = np.zeros((max(n, 1), 4), dtype=np.int64)
        lib().hibag_b200_model_train_trace(self._h, _p(out), n)
        return out[:n]

    def predict_stats(self):
        s = PredictStats()
        _chk(lib().hibag_b200_model_predict_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in PredictStats._fields_}

    def num_classifiers(self):
        return lib().hibag_b200_model_num_classifiers(self._h)

    def clear(self):
        _chk(lib().hibag_b200_model_clear(self._h))

    def classifier(self, k):
        ns, nh, acc = C.c_int(), C.c_int(), C.c_double()
        _chk(lib().hibag_b200_model_classifier_info(self._h, k, C.byref(ns), C.byref(nh), C.byref(acc)))
        snpidx = np.zeros(ns.value, dtype=np.int32)
        n_sn = lib().hibag_b200_model_classifier_samp_num_len(self._h, k)
        samp = np.zeros(max(n_sn, 1), dtype=np.int32)
        freq = np.zeros(nh.value); hla = np.zeros(nh.value, dtype=np.int32)
        packed = np.zeros((nh.value, 2), dtype=np.uint64)
        _chk(lib().hibag_b200_model_classifier_get(self._h, k, _p(snpidx), _p(samp), _p(freq), _p(hla), _p(packed)))
        return dict(snpidx=snpidx, samp_num=samp[:max(n_sn, 0)], freq=freq, hla=hla, packed=packed,
                    oob_acc=acc.value)

    def add_classifier(self, snpidx, freq, hla, packed, samp_num=None, oob_acc=0.0):
        s = np.ascontiguousarray(snpidx, dtype=np.int32)
        f = np.ascontiguousarray(freq, dtype=np.float64)
        h = np.ascontiguousarray(hla, dtype=np.int32)
        p = np.ascontiguousarray(packed, dtype=np.uint64)
        sn = None if samp_num is None else np.ascontiguousarray(samp_num, dtype=np.int32)
        _chk(lib().hibag_b200_model_add_classifier(self._h, len(s), _p(s), _p(sn),
                                                  0 if sn is None else len(sn), len(f), _p(f), _p(h), _p(p),
                                                  oob_acc))

    def snp_weights(self):
        w = np.zeros(self.n_snp, dtype=np.int32)
        _chk(lib().hibag_b200_model_snp_weights(self._h, _p(w)))
        return w

    def predict(self, geno, want_prob=True, want_dosage=True):
        """Host-buffer prediction: H2D of the raw genotypes and D2H of the results inside."""
        t0 = time.time()
        g = np.ascontiguousarray(geno, dtype=np.int8)
        assert g.ndim == 2 and g.shape[1] == self.n_snp
        n = g.shape[0]
        # large results land in page-locked memory (D2H at PCIe rate behind the next tile's scoring)
        def new(shape, dtype):
            if n * self.n_cells * 8 >= (8 << 20):
                try:
                    return pinned_empty(shape, dtype)
                except RuntimeError:              # host cannot page-lock that much: pageable (slower copy back)
                    pass
            return np.empty(shape, dtype)
        h1 = np.zeros(n, dtype=np.int32); h2 = np.zeros(n, dtype=np.int32)
        mp = np.zeros(n); mt = np.zeros(n)
        ds = new((n, self.n_hla), np.float64) if want_dosage else None
        pr = new((n, self.n_cells), np.float64) if want_prob else None
        out = PredictOut(_p(h1).value, _p(h2).value, _p(mp).value, _p(mt).value,
                         None if ds is None else _p(ds).value, None if pr is None else _p(pr).value)
        t1 = time.time()
        _chk(lib().hibag_b200_model_predict(self._h, _p(g), n, C.byref(out)))
        if os.environ.get("HIBAG_B200_PREDICT_DEBUG"):
            print("HLAModel.predict: result buffers %.1f ms, library call %.1f ms" % (
                1e3 * (t1 - t0), 1e3 * (time.time() - t1)), file=sys.stderr)
        return dict(h1=h1, h2=h2, prob=mp, matching=mt, dosage=ds, postprob=pr)

    def predict_device(self, geno_ptr, n_samp, h1=0, h2=0, max_prob=0, matching=0, dosage=0,
                       post_prob=0, stream=0, sync=True):
        """Device-resident prediction; all arguments are raw device pointers (ints)."""
        out = PredictOut(h1 or None, h2 or None, max_prob or None, matching or None,
                         dosage or None, post_prob or None)
        _chk(lib().hibag_b200_model_predict_device(self._h, C.c_void_p(geno_ptr), n_samp, C.byref(out),
                                                  C.c_void_p(stream) if stream else None, int(sync)))

    def predict_partial_device(self, geno_ptr, n_samp, snp_weight_ptr, acc_ptr, stream=0, sync=True):
        _chk(lib().hibag_b200_model_predict_partial_device(
            self._h, C.c_void_p(geno_ptr), n_samp, C.c_void_p(snp_weight_ptr) if snp_weight_ptr else None,
            C.c_void_p(acc_ptr), C.c_void_p(stream) if stream else None, int(sync)))

    # reference hlaModelToObj / hlaModelFromObj (R/HIBAG.R:1041-1178): the interchange format of
    # trained models (C side: HIBAG_GetClassifierList / HIBAG_NewClassifierHaplo, src/HIBAG.cpp:817-958)
    def to_obj(self):
        """The "hlaAttrBagObj" list as a dict: n_samp, n_snp, hla_allele, snp_id, classifiers =
        [{samp_num, snpidx (1-based), haplos {freq, hla (labels), haplo ("0101.." strings)},
        outofbag_acc}]."""
        cls = []
        for k in range(self.num_classifiers()):
            c = self.classifier(k)
            cls.append(dict(samp_num=c["samp_num"], snpidx=c["snpidx"] + 1,
                            haplos=dict(freq=c["freq"], hla=[self.hla_allele[i] for i in c["hla"]],
                                        haplo=[_bits(p, len(c["snpidx"])) for p in c["packed"]]),
                            outofbag_acc=c["oob_acc"]))
        return dict(n_samp=self.n_samp, n_snp=self.n_snp, hla_allele=list(self.hla_allele),
                    snp_id=self.snp_id, classifiers=cls)

    @staticmethod
    def from_obj(obj):
        m = HLAModel(obj["n_snp"], len(obj["hla_allele"]), obj["hla_allele"], obj.get("snp_id"))
        m.n_samp = int(obj.get("n_samp") or 0)
        alleles = [str(a) for a in obj["hla_allele"]]
        for c in obj["classifiers"]:
            hla = [alleles.index(str(x)) for x in c["haplos"]["hla"]]
            strs = list(c["haplos"]["haplo"])
            n_snp_c = len(np.atleast_1d(c["snpidx"]))
            if any(len(st) != n_snp_c for st in strs):
                raise ValueError("from_obj: haplotype string length differs from the number of SNPs")
            packed = np.array([_unbits(st) for st in strs], dtype=np.uint64).reshape(-1, 2)
            sn = c.get("samp_num")
            if sn is not None and len(sn) == 0:
                sn = None
            if sn is not None and m.n_samp and len(sn) != m.n_samp:
                raise ValueError("from_obj: samp_num has %d entries, the model %d samples" % (len(sn), m.n_samp))
            m.add_classifier(np.atleast_1d(np.asarray(c["snpidx"])) - 1, c["haplos"]["freq"], hla, packed,
                             samp_num=sn, oob_acc=float(c.get("outofbag_acc", 0)))
        return m

    # ---- on-disk forms ------------------------------------------------------------------------
    def save(self, path):
        """Write the model: *.json = the hlaAttrBagObj dict as JSON (frequencies as C99 hex floats, so
        the round trip is bit-exact); anything else = NumPy .npz (flat arrays with offsets)."""
        obj = self.to_obj()
        if str(path).endswith(".json"):
            import json
            js = dict(format="hibag_b200.hlaAttrBagObj/1", n_samp=int(obj["n_samp"]), n_snp=int(obj["n_snp"]),
                      hla_allele=[str(a) for a in obj["hla_allele"]],
                      snp_id=None if obj["snp_id"] is None else [str(x) for x in obj["snp_id"]],
                      classifiers=[dict(samp_num=[int(x) for x in c["samp_num"]],
                                        snpidx=[int(x) for x in c["snpidx"]],
                                        haplos=dict(freq=[float(x).hex() for x in c["haplos"]["freq"]],
                                                    hla=[str(x) for x in c["haplos"]["hla"]],
                                                    haplo=list(c["haplos"]["haplo"])),
                                        outofbag_acc=float(c["outofbag_acc"]).hex()) for c in obj["classifiers"]])
            with open(path, "w") as f:
                json.dump(js, f)
            return
        cls = [self.classifier(k) for k in range(self.num_classifiers())]
        cat = lambda key, dt: np.concatenate([np.asarray(c[key], dtype=dt).reshape(-1) for c in cls]) \
            if cls else np.zeros(0, dtype=dt)
        np.savez_compressed(
            path, format=np.array("hibag_b200.model/1"), n_snp=np.int64(self.n_snp), n_samp=np.int64(self.n_samp),
            hla_allele=np.array([str(a) for a in self.hla_allele]),
            snp_id=np.array([] if self.snp_id is None else [str(x) for x in self.snp_id]),
            snp_off=np.cumsum([0] + [len(c["snpidx"]) for c in cls]), snpidx=cat("snpidx", np.int32),
            hap_off=np.cumsum([0] + [len(c["freq"]) for c in cls]), freq=cat("freq", np.float64),
            hla=cat("hla", np.int32), packed=cat("packed", np.uint64).reshape(-1, 2),
            samp_off=np.cumsum([0] + [len(c["samp_num"]) for c in cls]), samp_num=cat("samp_num", np.int32),
            oob_acc=np.array([c["oob_acc"] for c in cls], dtype=np.float64))

    @staticmethod
    def load(path):
        """Read a model written by save() (.json / .npz) or an R workspace (.RData/.rdata/.rda/.rds-less
        RDX2 stream) holding one hlaAttrBagObj -- see hlaModelFromRData for workspaces with several."""
        p = str(path)
        if p.endswith(".json"):
            import json
            js = json.load(open(p))
            for c in js["classifiers"]:
                c["haplos"]["freq"] = [float.fromhex(x) for x in c["haplos"]["freq"]]
                c["outofbag_acc"] = float.fromhex(c["outofbag_acc"])
            return HLAModel.from_obj(js)
        if p.lower().endswith((".rdata", ".rda")):
            return hlaModelFromRData(p)
        z = np.load(p, allow_pickle=False)
        alleles = [str(a) for a in z["hla_allele"]]
        m = HLAModel(int(z["n_snp"]), len(alleles), alleles, [str(x) for x in z["snp_id"]] or None)
        m.n_samp = int(z["n_samp"])
        for k in range(len(z["oob_acc"])):
            a, b = z["snp_off"][k:k + 2]; q, r = z["hap_off"][k:k + 2]; u, v = z["samp_off"][k:k + 2]
            m.add_classifier(z["snpidx"][a:b], z["freq"][q:r], z["hla"][q:r], z["packed"][q:r],
                             samp_num=z["samp_num"][u:v] if v > u else None, oob_acc=float(z["oob_acc"][k]))
        return m


def _find_model_objs(obj, path=()):
    """(path, RObj) of every hlaAttrBagObj-shaped list inside a loaded R workspace"""
    from . import rdx2
    out = []
    if isinstance(obj, rdx2.RObj) and isinstance(obj.value, list) and obj.names():
        names = obj.names()
        if "classifiers" in names and "hla.allele" in names and "n.snp" in names:
            return [(path, obj)]
        for nm, v in zip(names, obj.value):
            out += _find_model_objs(v, path + (nm,))
    return out


def hlaModelFromRData(path, name=None):
    """Load a pre-fit model from an R workspace without R (reference: load() + hlaModelFromObj,
    R/HIBAG.R:1135-1178; e.g. inst/extdata/ModelList.RData holds `modellist$A`). `name` selects
    the object ("modellist/A" or just "A") when the file holds several."""
    from . import rdx2
    ws = rdx2.load(path)
    found = []
    for top, val in ws.items():
        found += _find_model_objs(val, (top,))
    if name is not None:
        found = [f for f in found if "/".join(f[0]) == name or f[0][-1] == name]
    if len(found) != 1:
        raise ValueError("hlaModelFromRData: %d model objects match in %s (%s)" % (
            len(found), path, ", ".join("/".join(f[0]) for f in found)))
    A = found[0][1]
    vec = lambda o: list(o.value) if isinstance(o.value, list) else o.value
    cls = []
    for c in A["classifiers"].value:
        h = c["haplos"]
        cls.append(dict(samp_num=np.asarray(c["samp.num"].value, dtype=np.int32),
                        snpidx=np.asarray(c["snpidx"].value, dtype=np.int32),
                        haplos=dict(freq=np.asarray(h["freq"].value, dtype=np.float64), hla=vec(h["hla"]),
                                    haplo=vec(h["haplo"])),
                        outofbag_acc=float(c["outofbag.acc"].value[0])))
    obj = dict(n_samp=int(A["n.samp"].value[0]), n_snp=int(A["n.snp"].value[0]),
               hla_allele=[str(a) for a in A["hla.allele"].value], snp_id=[str(x) for x in vec(A["snp.id"])],
               classifiers=cls)
    m = HLAModel.from_obj(obj)
    names = A.names()
    m.sample_id = vec(A["sample.id"]) if "sample.id" in names else None
    m.snp_position = np.asarray(A["snp.position"].value) if "snp.position" in names else None
    m.hla_locus = A["hla.locus"].value[0] if "hla.locus" in names else None
    return m


hlaModelToObj = HLAModel.to_obj
hlaModelFromObj = HLAModel.from_obj


def _bits(p, n):
    return "".join("1" if (int(p[j >> 6]) >> (j & 63)) & 1 else "0" for j in range(n))


def _unbits(s):
    w = [0, 0]
    for j, ch in enumerate(s):
        if ch == "1":
            w[j >> 6] |= 1 << (j & 63)
    return w


def default_mtry(n_snp, mtry="sqrt"):
    """reference R/HIBAG.R:180-208"""
    if mtry == "sqrt":
        v = math.ceil(math.sqrt(n_snp))
    elif mtry == "all":
        v = n_snp
    elif mtry == "one":
        v = 1
    else:
        v = float(mtry)
        if 0 < v < 1:
            v = n_snp * v
        v = min(math.ceil(v), n_snp)
    return max(int(v), 1)


def hlaAttrBagging(hla, snp, nclassifier=100, mtry="sqrt", prune=True, mono_rm=True, seed=100,
                   nthread=0, per_classifier_seed=False, use_legacy_hooks=False, verbose=False,
                   hla_allele=None, first_index=0, index_stride=1, n_concurrent=1, em_on_device=True,
                   screening=True):
    """Train a model. hla = (h1, h2) integer allele indices (or labels with hla_allele given),
    snp = int matrix [n_samp, n_snp] with 0/1/2 and anything else missing.
    Mirrors reference hlaAttrBagging (R/HIBAG.R:48-275): monomorphic SNPs are removed when mono_rm,
    mtry rule, then HIBAG_Training + HIBAG_NewClassifiers."""
    h1, h2 = (np.asarray(x) for x in hla)
    g = np.asarray(snp)
    assert g.ndim == 2 and len(h1) == g.shape[0] == len(h2)
    if hla_allele is None:
        n_hla = int(max(h1.max(), h2.max())) + 1
    else:
        n_hla = len(hla_allele)
    keep = np.arange(g.shape[1])
    if mono_rm:                                   # R/HIBAG.R:117-155
        valid = (g >= 0) & (g <= 2)
        cnt = valid.sum(axis=0)
        mf = np.where(cnt > 0, (g * valid).sum(axis=0) / np.maximum(cnt, 1) * 0.5, 0.0)
        mf = np.minimum(mf, 1 - mf)
        keep = np.nonzero(mf > 0)[0]
        g = g[:, keep]
    model = HLAModel(g.shape[1], n_hla, hla_allele)
    model.snp_sel = keep
    model.set_training(g, h1, h2)
    model.train(nclassifier, default_mtry(g.shape[1], mtry), prune=prune, seed=seed, n_threads=nthread,
                per_classifier_seed=per_classifier_seed, first_index=first_index,
                index_stride=index_stride, use_legacy_hooks=use_legacy_hooks, verbose=int(verbose),
                n_concurrent=n_concurrent, em_on_device=em_on_device, screening=screening)
    return model


def bed_decode(bed_bytes, n_samp, n_snp, snp_flag=None, return_ms=False):
    """PLINK .bed bytes (whole file, prefix included) -> int8 [n_samp][n_kept] on the GPU
    (reference HIBAG_ConvBED, src/HIBAG.cpp:1094-1191; missing = -1)"""
    raw = np.frombuffer(bed_bytes, dtype=np.uint8) if not isinstance(bed_bytes, np.ndarray) else \
        np.ascontiguousarray(bed_bytes, dtype=np.uint8)
    flag = None if snp_flag is None else np.ascontiguousarray(snp_flag, dtype=np.int32)
    n_save, ms = C.c_int(), C.c_double()
    _chk(lib().hibag_b200_bed_decode(_p(raw), raw.size, n_samp, n_snp, _p(flag), None, C.byref(n_save), None))
    out = np.zeros((n_samp, n_save.value), dtype=np.int8)
    if n_save.value > 0:
        _chk(lib().hibag_b200_bed_decode(_p(raw), raw.size, n_samp, n_snp, _p(flag), _p(out), C.byref(n_save),
                                         C.byref(ms)))
    return (out, ms.value) if return_ms else out


def bed_decode_device(payload_ptr, mode, n_samp, n_snp, sel_ptr, n_save, out_ptr, stream=0):
    """device-resident BED decoding: payload / selection / output are device pointers"""
    _chk(lib().hibag_b200_bed_decode_device(C.c_void_p(payload_ptr), mode, n_samp, n_snp,
                                            C.c_void_p(sel_ptr) if sel_ptr else None, n_save,
                                            C.c_void_p(out_ptr), C.c_void_p(stream)))


def hlaBED2Geno(bed_fn, fam_fn, bim_fn, import_chr="", region=None, snp_flag=None):
    """Import a PLINK binary file set (reference hlaBED2Geno, R/DataUtilities.R:703-780): sample ids
    from the .fam (InvID, or FamilyID-InvID when not unique), SNP ids / positions / alleles from the
    .bim, genotypes decoded on the GPU. SNP selection: import_chr = "" keeps every SNP, a
    chromosome name or list keeps those with a positive position (R/DataUtilities.R:688-697);
    region = (chr, start, end) keeps a position window (the reference's "xMHC" preset needs its
    per-assembly gene table, which is R-side data: pass the window instead); snp_flag overrides.
    Returns a dict shaped like hlaSNPGenoClass with genotype as int8 [n_samp][n_snp] (sample-major,
    what hlaAttrBagging / hlaPredict take here)."""
    fam = [ln.split() for ln in open(fam_fn) if ln.strip()]
    bim = [ln.split() for ln in open(bim_fn) if ln.strip()]
    inv = [f[1] for f in fam]
    if len(set(inv)) == len(inv):
        sample_id = inv
    else:
        sample_id = ["%s-%s" % (f[0], f[1]) for f in fam]
        if len(set(sample_id)) != len(sample_id):
            raise ValueError("IDs in PLINK bed are not unique!")
    snp_id = [b[1] for b in bim]
    if len(set(snp_id)) != len(snp_id):
        raise ValueError("The SNP IDs in the PLINK binary file should be unique!")
    chrom = np.array([b[0] for b in bim])
    pos = np.array([int(b[3]) if b[3].lstrip("-").isdigit() else 0 for b in bim], dtype=np.int64)
    if snp_flag is not None:
        flag = np.asarray(snp_flag, dtype=bool)
    elif region is not None:
        flag = (chrom == str(region[0])) & (pos >= region[1]) & (pos <= region[2])
    elif import_chr == "" or import_chr is None:
        flag = np.ones(len(bim), dtype=bool)
    else:
        want = [str(c) for c in (import_chr if isinstance(import_chr, (list, tuple)) else [import_chr])]
        flag = np.isin(chrom, want) & (pos > 0)
    if not flag.any():
        raise ValueError("There is no SNP imported.")
    g = bed_decode(np.fromfile(bed_fn, dtype=np.uint8), len(fam), len(bim), flag.astype(np.int32))
    keep = np.nonzero(flag)[0]
    return dict(genotype=g, sample_id=sample_id, snp_id=[snp_id[i] for i in keep], snp_position=pos[keep],
                snp_allele=["%s/%s" % (bim[i][4], bim[i][5]) for i in keep])


def hlaPredict(model, snp, type="response+prob"):
    """Predict HLA types (reference hlaPredict, R/HIBAG.R:481-818, vote = "prob").
    type: "response" (best guess, prob, matching), "dosage", "prob", "response+dosage",
    "response+prob"."""
    g = np.asarray(snp)
    if getattr(model, "snp_sel", None) is not None and g.shape[1] != model.n_snp:
        g = g[:, model.snp_sel]
    want_prob = type in ("prob", "response+prob")
    want_dosage = type in ("dosage", "response+dosage", "response+prob")
    r = model.predict(g, want_prob=want_prob, want_dosage=want_dosage)
    r["allele1"] = [model.hla_allele[i] if i != NA_INTEGER else None for i in r["h1"]]
    r["allele2"] = [model.hla_allele[i] if i != NA_INTEGER else None for i in r["h2"]]
    return r
