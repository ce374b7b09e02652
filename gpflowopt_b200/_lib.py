"""ctypes binding of libgpfo.so (include/gpfo.h).  Fails loudly when the CUDA library is missing: there is
no CPU fallback for the acquisition path."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpfo.so")

OK, ERR_NOT_SPD, ERR_BAD_SHAPE, ERR_CUDA, ERR_UNSUPPORTED, ERR_BAD_PROGRAM, ERR_STATE = range(7)
KERN_RBF, KERN_MATERN52, KERN_MATERN32 = 0, 1, 2
OP_EI, OP_POI, OP_POF, OP_LCB, OP_HVPOI, OP_SUM, OP_PROD, OP_SCALE, OP_MEAN, OP_VAR, OP_MES = range(1, 12)
MAX_GPS, MAX_SLOTS, MAX_OPS, MAX_PARAMS, MAX_DIM = 16, 16, 64, 64, 64

# every symbol include/gpfo.h declares
EXPORTS = ["gpfo_version", "gpfo_ctx_create", "gpfo_ctx_destroy", "gpfo_last_error", "gpfo_status_string",
           "gpfo_gp_set", "gpfo_gp_count_set", "gpfo_gp_get_factor", "gpfo_program_set", "gpfo_cells_set",
           "gpfo_eval", "gpfo_predict", "gpfo_timings_get", "gpfo_set_variant", "gpfo_measure_dmma_peak",
           "gpfo_measure_fp64_pipes", "gpfo_pareto_cells", "gpfo_eval_random", "gpfo_random_rows", "gpfo_ndtr_fast_host",
           "gpfo_non_dominated_sort", "gpfo_select_variant", "gpfo_mt19937_design_rows"]


class Timings(ctypes.Structure):
    _fields_ = [("factor_ms", ctypes.c_double), ("h2d_ms", ctypes.c_double), ("eval_ms", ctypes.c_double),
                ("d2h_ms", ctypes.c_double), ("hot_kernel_ms", ctypes.c_double), ("launches", ctypes.c_int64),
                ("variant", ctypes.c_int64)]


class GpfoError(RuntimeError):
    def __init__(self, status, message):
        super(GpfoError, self).__init__("libgpfo status %d: %s" % (status, message))
        self.status = status


class NotPositiveDefiniteError(GpfoError):
    """Cholesky failure: the role tf.errors.InvalidArgumentError plays in the reference
    (gpflowopt/acquisition/acquisition.py:117, gpflowopt/bo.py:51)."""


_lib = None


def load():
    """Loads libgpfo.so (building is the job of gpflowopt_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgpfo.so is missing (%s). Build it with `python -m gpflowopt_b200.build`; "
                          "there is no CPU fallback for the acquisition path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_dp = ctypes.c_void_p
    vp = ctypes.c_void_p
    lib.gpfo_version.restype = ctypes.c_int
    lib.gpfo_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.gpfo_ctx_destroy.argtypes = [vp]
    lib.gpfo_last_error.argtypes = [vp]
    lib.gpfo_last_error.restype = ctypes.c_char_p
    lib.gpfo_status_string.argtypes = [ctypes.c_int]
    lib.gpfo_status_string.restype = ctypes.c_char_p
    lib.gpfo_gp_set.argtypes = [vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, c_dp, c_dp, ctypes.c_int,
                                ctypes.c_int, c_dp, ctypes.c_double, ctypes.c_double, c_dp, c_dp, c_dp, c_dp,
                                ctypes.c_int]
    lib.gpfo_gp_count_set.argtypes = [vp, ctypes.c_int]
    lib.gpfo_gp_get_factor.argtypes = [vp, ctypes.c_int, c_dp, c_dp, c_dp]
    lib.gpfo_program_set.argtypes = [vp, c_dp, ctypes.c_int, c_dp, ctypes.c_int]
    lib.gpfo_cells_set.argtypes = [vp, c_dp, ctypes.c_int, ctypes.c_int, c_dp, c_dp, ctypes.c_int64]
    lib.gpfo_eval.argtypes = [vp, c_dp, ctypes.c_int64, ctypes.c_int, c_dp, c_dp, ctypes.POINTER(ctypes.c_double),
                              ctypes.POINTER(ctypes.c_int64)]
    lib.gpfo_predict.argtypes = [vp, ctypes.c_int, c_dp, ctypes.c_int64, ctypes.c_int, c_dp, c_dp]
    lib.gpfo_timings_get.argtypes = [vp, ctypes.POINTER(Timings)]
    lib.gpfo_set_variant.argtypes = [vp, ctypes.c_int]
    lib.gpfo_measure_dmma_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.gpfo_measure_fp64_pipes.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                            ctypes.POINTER(ctypes.c_double)]
    lib.gpfo_eval_random.argtypes = [vp, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_dp, c_dp, c_dp,
                                     ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    lib.gpfo_random_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_dp, c_dp, c_dp]
    lib.gpfo_ndtr_fast_host.argtypes = [c_dp, ctypes.c_int64, c_dp]
    lib.gpfo_non_dominated_sort.argtypes = [c_dp, ctypes.c_int64, ctypes.c_int, c_dp]
    lib.gpfo_select_variant.argtypes = [vp, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
    lib.gpfo_mt19937_design_rows.argtypes = [c_dp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int64, ctypes.c_int, c_dp, c_dp, c_dp]
    lib.gpfo_pareto_cells.argtypes = [c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_double, c_dp, c_dp, ctypes.c_int64,
                                      ctypes.POINTER(ctypes.c_int64)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is ctypes.c_int and name not in ("gpfo_version",):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def numpy_stream_design_rows(n, scale, shift, out=None):
    """``np.random.rand(n, D) * scale + shift`` drawn natively from -- and advancing -- NumPy's global legacy RandomState
    (csrc/pareto_host.cu::gpfo_mt19937_design_rows): the same bytes, about 4 x faster.  Returns None when the global generator
    is not the MT19937 RandomState (then the caller falls back to NumPy)."""
    state = np.random.get_state()
    if state[0] != 'MT19937':
        return None
    lib = load()
    scale, shift = _f64(scale).ravel(), _f64(shift).ravel()
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    pos = ctypes.c_int32(int(state[2]))
    if out is None:
        out = np.empty((int(n), scale.size))
    assert out.shape == (int(n), scale.size) and out.dtype == np.float64 and out.flags.c_contiguous
    st = lib.gpfo_mt19937_design_rows(_ptr(key), ctypes.byref(pos), int(n), scale.size, _ptr(scale), _ptr(shift), _ptr(out) if n else None)
    if st != OK:
        raise GpfoError(st, "gpfo_mt19937_design_rows failed")
    np.random.set_state((state[0], key, pos.value) + tuple(state[3:]))
    return out


def non_dominated_counts(Y):
    """Dominance counts of gpflowopt/pareto.py:78-90 (native host code, csrc/pareto_host.cu); no GPU needed."""
    lib = load()
    Y = _f64(Y)
    if Y.ndim != 2:
        raise GpfoError(ERR_BAD_SHAPE, "objectives must be n x R")
    out = np.zeros(Y.shape[0], dtype=np.int64)
    st = lib.gpfo_non_dominated_sort(_ptr(Y) if Y.size else None, Y.shape[0], Y.shape[1], _ptr(out) if Y.size else None) \
        if Y.shape[0] else OK
    if st != OK:
        raise GpfoError(st, "gpfo_non_dominated_sort failed")
    return out


def ndtr_fast_host(z):
    """The HVPoI kernel's branch-free Normal CDF (csrc/ndtr_fast.cuh), evaluated on the host (tests; no GPU needed)."""
    lib = load()
    z = _f64(z).ravel()
    out = np.empty_like(z)
    st = lib.gpfo_ndtr_fast_host(_ptr(z), z.size, _ptr(out))
    if st != OK:
        raise GpfoError(st, "gpfo_ndtr_fast_host failed")
    return out


def random_rows(seed, first_row, n, lower, upper):
    """Rows [first_row, first_row + n) of the device candidate stream, regenerated on the host (no GPU needed)."""
    lib = load()
    lower, upper = _f64(lower).ravel(), _f64(upper).ravel()
    out = np.empty((int(n), lower.size))
    st = lib.gpfo_random_rows(int(seed), int(first_row), int(n), lower.size, _ptr(lower), _ptr(upper), _ptr(out) if n else None)
    if st != OK and n:
        raise GpfoError(st, "gpfo_random_rows failed")
    return out


def pareto_cells(front, order, threshold=0.0):
    """Cell decomposition of the non-dominated region (native host code, no GPU needed).  Returns (lb, ub) int32 C x R."""
    lib = load()
    front = _f64(front)
    P, R = front.shape
    order = np.ascontiguousarray(order, dtype=np.int32)
    cap = max(64, 4 * P * P)
    while True:
        lb = np.empty((cap, R), dtype=np.int32)
        ub = np.empty((cap, R), dtype=np.int32)
        n = ctypes.c_int64(0)
        st = lib.gpfo_pareto_cells(_ptr(front), _ptr(order), P, R, float(threshold), _ptr(lb), _ptr(ub), cap, ctypes.byref(n))
        if st == OK:
            return lb[:n.value].copy(), ub[:n.value].copy()
        if n.value > cap:
            cap = int(n.value)
            continue
        raise GpfoError(st, "gpfo_pareto_cells failed")


class Context(object):
    """Owns one gpfo_ctx (one CUDA device, one stream).  Not thread-safe."""

    def __init__(self, device=0):
        self._lib = load()
        h = ctypes.c_void_p()
        st = self._lib.gpfo_ctx_create(int(device), ctypes.byref(h))
        if st != OK:
            raise GpfoError(st, "gpfo_ctx_create(device=%d) failed: %s (a CUDA sm_100a device is required; there is "
                                "no CPU fallback)" % (device, self._lib.gpfo_status_string(st).decode()))
        self._h = h
        self.device = int(device)
        self.variant = 0
        self.cum = dict(eval_ms=0.0, hot_kernel_ms=0.0, h2d_ms=0.0, launches=0, calls=0)   # running totals over gpfo_eval calls

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gpfo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st == OK:
            return
        msg = self._lib.gpfo_last_error(self._h).decode()
        if st == ERR_NOT_SPD:
            raise NotPositiveDefiniteError(st, msg)
        raise GpfoError(st, msg)

    # ---- model state -------------------------------------------------------------------------------
    def gp_set(self, gp_id, X, Y, kind, lengthscales, variance, noise_variance, in_scale=None, in_shift=None,
               out_scale=None, out_shift=None, slot0=0):
        X = _f64(X)
        Y = _f64(Y)
        if X.ndim != 2 or Y.ndim != 2 or X.shape[0] != Y.shape[0]:
            raise GpfoError(ERR_BAD_SHAPE, "X must be M x D and Y M x R")
        M, D = X.shape
        R = Y.shape[1]
        ell = np.atleast_1d(_f64(lengthscales)).ravel()
        ard = 1 if ell.size > 1 or D == 1 else 0
        if ell.size not in (1, D):
            raise GpfoError(ERR_BAD_SHAPE, "lengthscales must have 1 or D entries")
        vecs = []
        for v, n, default in ((in_scale, D, 1.0), (in_shift, D, 0.0), (out_scale, R, 1.0), (out_shift, R, 0.0)):
            vecs.append(np.full(n, default) if v is None else _f64(v, (n,)))
        self._check(self._lib.gpfo_gp_set(self._h, gp_id, M, D, R, _ptr(X), _ptr(Y), int(kind), ard, _ptr(ell),
                                          float(variance), float(noise_variance), _ptr(vecs[0]), _ptr(vecs[1]),
                                          _ptr(vecs[2]), _ptr(vecs[3]), int(slot0)))

    def gp_count_set(self, n):
        self._check(self._lib.gpfo_gp_count_set(self._h, int(n)))

    def gp_get_factor(self, gp_id, M, R):
        L, Linv, V = np.empty((M, M)), np.empty((M, M)), np.empty((M, R))
        self._check(self._lib.gpfo_gp_get_factor(self._h, gp_id, _ptr(L), _ptr(Linv), _ptr(V)))
        return L, Linv, V

    def program_set(self, ops, params):
        ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
        params = _f64(params).ravel()
        self._check(self._lib.gpfo_program_set(self._h, _ptr(ops), ops.shape[0], _ptr(params) if params.size else None,
                                               params.size))

    def cells_set(self, pf_ext, lb, ub):
        pf_ext = _f64(pf_ext)
        lb = np.ascontiguousarray(lb, dtype=np.int32).reshape(-1, pf_ext.shape[1])
        ub = np.ascontiguousarray(ub, dtype=np.int32).reshape(-1, pf_ext.shape[1])
        self._check(self._lib.gpfo_cells_set(self._h, _ptr(pf_ext), pf_ext.shape[0], pf_ext.shape[1],
                                             _ptr(lb) if lb.size else None, _ptr(ub) if ub.size else None, lb.shape[0]))

    # ---- hot path ----------------------------------------------------------------------------------
    def eval(self, X, want_values=True, want_grads=False):
        """X: N x D float64 ndarray (host).  Returns (values N x 1 or None, grads N x D or None, best_value, best_index)."""
        X = _f64(X)
        if X.ndim != 2:
            raise GpfoError(ERR_BAD_SHAPE, "candidates must be a 2-D float64 array")
        N, D = X.shape
        vals = np.empty((N, 1)) if want_values else None
        grads = np.empty((N, D)) if want_grads else None
        bv, bi = ctypes.c_double(0.0), ctypes.c_int64(-1)
        self._check(self._lib.gpfo_eval(self._h, _ptr(X) if N else None, N, D, _ptr(vals) if N else None,
                                        _ptr(grads) if (want_grads and N) else None, ctypes.byref(bv), ctypes.byref(bi)))
        self._accumulate()
        return vals, grads, bv.value, bi.value

    def eval_device(self, x_ptr, N, D, values_ptr=None, grads_ptr=None):
        """Device-resident candidates / outputs given as raw device pointers (e.g. torch.Tensor.data_ptr())."""
        bv, bi = ctypes.c_double(0.0), ctypes.c_int64(-1)
        self._check(self._lib.gpfo_eval(self._h, ctypes.c_void_p(x_ptr), int(N), int(D),
                                        ctypes.c_void_p(values_ptr) if values_ptr else None,
                                        ctypes.c_void_p(grads_ptr) if grads_ptr else None,
                                        ctypes.byref(bv), ctypes.byref(bi)))
        self._accumulate()
        return bv.value, bi.value

    def eval_random(self, seed, first_row, N, lower, upper, want_values=False):
        """Scores N device-generated candidates (rows first_row .. first_row+N-1 of the stream keyed by seed)."""
        lower, upper = _f64(lower).ravel(), _f64(upper).ravel()
        vals = np.empty((N, 1)) if want_values else None
        bv, bi = ctypes.c_double(0.0), ctypes.c_int64(-1)
        self._check(self._lib.gpfo_eval_random(self._h, int(seed), int(first_row), int(N), lower.size, _ptr(lower), _ptr(upper),
                                               _ptr(vals) if (want_values and N) else None, ctypes.byref(bv), ctypes.byref(bi)))
        self._accumulate()
        return vals, bv.value, bi.value

    def predict(self, gp_id, X, R):
        X = _f64(X)
        N, D = X.shape
        mean, var = np.empty((N, R)), np.empty((N, R))
        self._check(self._lib.gpfo_predict(self._h, gp_id, _ptr(X) if N else None, N, D, _ptr(mean) if N else None,
                                           _ptr(var) if N else None))
        return mean, var

    def _accumulate(self):
        t = Timings()
        if self._lib.gpfo_timings_get(self._h, ctypes.byref(t)) == OK:
            c = self.cum
            c["eval_ms"] += t.eval_ms; c["hot_kernel_ms"] += t.hot_kernel_ms; c["h2d_ms"] += t.h2d_ms
            c["launches"] += t.launches; c["calls"] += 1

    def reset_cum(self):
        for k in self.cum:
            self.cum[k] = 0 if k in ("launches", "calls") else 0.0

    def timings(self):
        t = Timings()
        self._check(self._lib.gpfo_timings_get(self._h, ctypes.byref(t)))
        return {k: getattr(t, k) for k, _ in Timings._fields_}

    def set_variant(self, v):
        self._check(self._lib.gpfo_set_variant(self._h, int(v)))
        self.variant = int(v)

    def select_variant(self, n):
        """Forward kernel the library would pick for a batch of n candidates (see gpfo_select_variant)."""
        v = ctypes.c_int(0)
        self._check(self._lib.gpfo_select_variant(self._h, int(n), ctypes.byref(v)))
        return v.value

    def measure_fp64_pipes(self, mode):
        a, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
        self._check(self._lib.gpfo_measure_fp64_pipes(self._h, int(mode), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def measure_dmma_peak(self):
        v = ctypes.c_double(0.0)
        self._check(self._lib.gpfo_measure_dmma_peak(self._h, ctypes.byref(v)))
        return v.value
