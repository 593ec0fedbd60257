"""ctypes binding of include/ntn_b200.h (the C-ABI shared library ``libntn_b200.so``).

No fallback of any kind: if the library has not been built (``python -c 'import
__graft_entry__ as g; g.build()'`` or ``make -C .../csrc``) loading raises, and every
compute call raises ``NTNError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libntn_b200.so")

NTN_OK, NTN_EINVAL, NTN_ECUDA, NTN_ESTATE, NTN_EUNSUP = 0, -1, -2, -3, -4
ACT_TANH, ACT_SIGMOID = 0, 1
WV_FROZEN, WV_THETA = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
SCORE_REFERENCE_EVAL, SCORE_TRAIN = 0, 1
NUM_PHASES = 9


class NTNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ntn_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("d", C.c_int32), ("k", C.c_int32), ("R", C.c_int32), ("Ne", C.c_int32), ("Nw", C.c_int32),
                ("batch_divisor", C.c_int32), ("lambda_", C.c_float), ("activation", C.c_int32),
                ("wv_source", C.c_int32), ("gemm_path", C.c_int32), ("device", C.c_int32),
                ("reg_scale", C.c_float)]


REDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64)


class LbfgsOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("history", C.c_int32), ("tol", C.c_double),
                ("reduce_fn", REDUCE_FN), ("reduce_ctx", C.c_void_p)]


class LbfgsResult(C.Structure):
    _fields_ = [("min_obj_val", C.c_double), ("first_cost", C.c_double), ("iters", C.c_int32), ("evals", C.c_int32),
                ("did_converge", C.c_int32), ("n_active_last", C.c_int64)]


SIDES_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_uint8))

# name -> (restype, argtypes): every symbol include/ntn_b200.h declares
_P = C.c_void_p
_I32P = C.POINTER(C.c_int32)
SYMBOLS = {
    "ntn_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "ntn_destroy": (None, [_P]),
    "ntn_last_error": (C.c_char_p, [_P]),
    "ntn_dimension": (C.c_int64, [_P]),
    "ntn_set_entity_words": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ntn_set_frozen_wv": (C.c_int, [_P, _P]),
    "ntn_set_batch": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "ntn_set_batch_dev": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "ntn_set_training_triples": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "ntn_generate_batch_dev": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32]),
    "ntn_get_batch_columns": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P]),
    "ntn_eval": (C.c_int, [_P, _P, _P, C.POINTER(C.c_double), _P, C.POINTER(C.c_int64)]),
    "ntn_eval_f32": (C.c_int, [_P, _P, _P, C.POINTER(C.c_double), _P, C.POINTER(C.c_int64)]),
    "ntn_eval_dev": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ntn_last_cost": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ntn_result_dev": (_P, [_P]),
    "ntn_set_grad_tail": (C.c_int, [_P, C.c_int]),
    "ntn_exchange_bind": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P, C.c_uint64, C.c_int64, _P, C.c_int64]),
    "ntn_exchange_scratch_numel": (C.c_int64, [_P]),
    "ntn_exchange_bind_ge": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_int64, _P, C.c_int64]),
    "ntn_exchange_ge_scratch_numel": (C.c_int64, [_P]),
    "ntn_exchange_unbind": (C.c_int, [_P]),
    "ntn_exchange_check": (C.c_int, [_P]),
    "ntn_exchange_probe": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "ntn_exchange_mode": (C.c_char_p, [_P]),
    "ntn_internal_dimension": (C.c_int64, [_P]),
    "ntn_to_internal_dev": (C.c_int, [_P, _P, _P]),
    "ntn_from_internal_dev": (C.c_int, [_P, _P, _P]),
    "ntn_upload_theta": (C.c_int, [_P, _P, _P]),
    "ntn_download_theta": (C.c_int, [_P, _P, _P]),
    "ntn_set_stream": (C.c_int, [_P, _P]),
    "ntn_reset_stream": (C.c_int, [_P]),
    "ntn_get_stream": (_P, [_P]),
    "ntn_lbfgs_minimize": (C.c_int, [_P, _P, C.POINTER(LbfgsOpts), SIDES_FN, _P, C.POINTER(LbfgsResult)]),
    "ntn_score": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int32, _P]),
    "ntn_last_launch_count": (C.c_int, [_P]),
    "ntn_active_gemm_path": (C.c_int, [_P]),
    "ntn_set_profiling": (C.c_int, [_P, C.c_int]),
    "ntn_last_phase_ms": (C.c_int, [_P, _P]),
    "ntn_phase_name": (C.c_char_p, [C.c_int]),
    "ntn_debug_fetch": (C.c_int64, [_P, C.c_char_p, _P, C.c_int64]),
    "ntn_host_java_random": (C.c_int, [C.c_int64, C.c_int32, C.c_int64, _P, C.c_int64, _P]),
    "ntn_host_random_ints": (C.c_int, [C.POINTER(C.c_uint64), C.c_int32, C.c_int64, _P]),
    "ntn_host_sides_draw": (None, [_P, _P]),
    "ntn_host_entity_word_maps": (C.c_int, [C.c_char_p, C.c_char_p, _P, _P, _P, _P, _P, C.c_int64,
                                            C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ntn_host_collections_shuffle": (C.c_int, [C.c_int64, C.c_int32, _P]),
    "ntn_host_fastdiv": (C.c_int, [C.c_uint32, C.c_int64, _P, _P]),
    "ntn_host_word_chunks": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P]),
    "ntn_host_convert_f64_f32": (C.c_int, [_P, _P, C.c_int64]),
    "ntn_host_convert_f32_f64": (C.c_int, [_P, _P, C.c_int64]),
    "ntn_host_generate_batch": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "ntn_host_theta_checkpoint_roundtrip": (C.c_int, [C.c_char_p, C.c_int32, _P, C.c_int64, _P]),
    "ntn_version": (C.c_char_p, []),
}

_lib = None


def load():
    """dlopen the library and type every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)        # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Handle:
    """Thin RAII wrapper over ntn_handle*.  All arrays are NumPy; device pointers are ints."""

    def __init__(self, d, k, R, Ne, Nw, batch_divisor, lam, activation="tanh", wv_source="frozen",
                 gemm_path="auto", device=0, reg_scale=1.0):
        self.lib = load()
        cfg = Config(d, k, R, Ne, Nw, batch_divisor, lam,
                     {"tanh": ACT_TANH, "sigmoid": ACT_SIGMOID}[activation],
                     {"frozen": WV_FROZEN, "theta": WV_THETA}[wv_source],
                     {"auto": GEMM_AUTO, "simt": GEMM_SIMT, "tcgen05": GEMM_TCGEN05}[gemm_path],
                     device, reg_scale)
        self.h = C.c_void_p()
        rc = self.lib.ntn_create(C.byref(cfg), C.byref(self.h))
        if rc != NTN_OK:
            self.h = None
            raise NTNError(rc, self.lib.ntn_last_error(None).decode())
        self.P = self.lib.ntn_dimension(self.h)
        self.Pint = self.lib.ntn_internal_dimension(self.h)
        self.R = R

    def _ck(self, rc):
        if rc != NTN_OK:
            raise NTNError(rc, self.lib.ntn_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.ntn_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- state ----
    def set_entity_words(self, fwd_off, fwd_idx, bwd_off=None, bwd_idx=None, len_bwd=None):
        fo, fi = _i32(fwd_off), _i32(fwd_idx)
        bo = None if bwd_off is None else _i32(bwd_off)
        bi = None if bwd_idx is None else _i32(bwd_idx)
        lb = _i32(np.diff(fo) if len_bwd is None else len_bwd)
        self._ck(self.lib.ntn_set_entity_words(self.h, _ptr(fo), _ptr(fi), _ptr(bo), _ptr(bi), _ptr(lb)))

    def set_frozen_wv(self, wv):
        wv = np.ascontiguousarray(wv, dtype=np.float32)
        self._ck(self.lib.ntn_set_frozen_wv(self.h, _ptr(wv)))

    def set_batch(self, e1, rel, e2, e3):
        e1, rel, e2, e3 = _i32(e1), _i32(rel), _i32(e2), _i32(e3)
        assert len(e1) == len(rel) == len(e2) == len(e3)
        self._ck(self.lib.ntn_set_batch(self.h, len(e1), _ptr(e1), _ptr(rel), _ptr(e2), _ptr(e3)))

    def set_batch_dev(self, n: int, d_e1: int, d_rel: int, d_e2: int, d_e3: int):
        """The batch columns as int32 device pointers (e.g. torch tensors' data_ptr())."""
        self._ck(self.lib.ntn_set_batch_dev(self.h, int(n), d_e1, d_rel, d_e2, d_e3))

    def set_training_triples(self, train):
        """DataFactory.trainingTripples (T x 3: e1, rel, e2) into HBM, for generate_batch_dev."""
        t = np.asarray(train)
        e1, rel, e2 = _i32(t[:, 0]), _i32(t[:, 1]), _i32(t[:, 2])
        self._ck(self.lib.ntn_set_training_triples(self.h, len(e1), _ptr(e1), _ptr(rel), _ptr(e2)))

    def generate_batch_dev(self, first, batch, corrupt, seed, job, filter_true=False):
        """Batch job from training triples [first, first+batch) with corrupt entities sampled on the device."""
        self._ck(self.lib.ntn_generate_batch_dev(self.h, int(first), int(batch), int(corrupt), int(seed) & (2**64 - 1),
                                                 int(job), 1 if filter_true else 0))

    def get_batch_columns(self, n):
        cols = [np.empty(n, np.int32) for _ in range(4)]
        self._ck(self.lib.ntn_get_batch_columns(self.h, int(n), *(_ptr(c) for c in cols)))
        return tuple(cols)

    # ---- hot path ----
    def _side(self, corrupt_side):
        s = np.ascontiguousarray(corrupt_side, dtype=np.uint8)
        assert s.size == self.R
        return s

    def eval(self, theta, corrupt_side, grad_out=None):
        """computeAt with host double arrays in the reference's order -> (cost, grad, n_active)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        assert theta.size == self.P
        grad = np.empty(self.P, np.float64) if grad_out is None else grad_out
        s = self._side(corrupt_side)
        cost, nact = C.c_double(), C.c_int64()
        self._ck(self.lib.ntn_eval(self.h, _ptr(theta), _ptr(s), C.byref(cost), _ptr(grad), C.byref(nact)))
        return cost.value, grad, nact.value

    def eval_f32(self, theta, corrupt_side, grad_out=None):
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        assert theta.size == self.P
        grad = np.empty(self.P, np.float32) if grad_out is None else grad_out
        s = self._side(corrupt_side)
        cost, nact = C.c_double(), C.c_int64()
        self._ck(self.lib.ntn_eval_f32(self.h, _ptr(theta), _ptr(s), C.byref(cost), _ptr(grad), C.byref(nact)))
        return cost.value, grad, nact.value

    def eval_dev(self, d_theta: int, corrupt_side, d_grad: int, sync=True):
        s = self._side(corrupt_side)
        if sync:
            cost, nact = C.c_double(), C.c_int64()
            self._ck(self.lib.ntn_eval_dev(self.h, d_theta, _ptr(s), d_grad, C.byref(cost), C.byref(nact)))
            return cost.value, nact.value
        self._ck(self.lib.ntn_eval_dev(self.h, d_theta, _ptr(s), d_grad, None, None))
        return None

    def last_cost(self):
        cost, nact = C.c_double(), C.c_int64()
        self._ck(self.lib.ntn_last_cost(self.h, C.byref(cost), C.byref(nact)))
        return cost.value, nact.value

    def set_grad_tail(self, on=True):
        self._ck(self.lib.ntn_set_grad_tail(self.h, int(on)))

    def exchange_bind(self, rank, world, buffer_ptrs, signal_pad_ptrs, multicast_ptr, numel, scratch_ptrs=None, scratch_numel=0):
        """Bind one symmetric gradient buffer for the in-evaluation exchange (include/ntn_b200.h: ntn_exchange_bind)."""
        bp = np.asarray(buffer_ptrs, np.uint64)
        sp = np.asarray(signal_pad_ptrs, np.uint64)
        cp = None if scratch_ptrs is None else np.asarray(scratch_ptrs, np.uint64)
        self._ck(self.lib.ntn_exchange_bind(self.h, int(rank), int(world), _ptr(bp), _ptr(sp), int(multicast_ptr), int(numel),
                                            _ptr(cp), int(scratch_numel)))
        return True

    def exchange_bind_ge(self, buffer_ptrs, signal_pad_ptrs, multicast_ptr, numel, scratch_ptrs=None, scratch_numel=0):
        """Bind a symmetric entity-gradient buffer (include/ntn_b200.h: ntn_exchange_bind_ge)."""
        bp, sp = np.asarray(buffer_ptrs, np.uint64), np.asarray(signal_pad_ptrs, np.uint64)
        cp = None if scratch_ptrs is None else np.asarray(scratch_ptrs, np.uint64)
        self._ck(self.lib.ntn_exchange_bind_ge(self.h, _ptr(bp), _ptr(sp), int(multicast_ptr), int(numel), _ptr(cp), int(scratch_numel)))
        return True

    @property
    def exchange_ge_scratch_numel(self):
        return int(self.lib.ntn_exchange_ge_scratch_numel(self.h))

    @property
    def exchange_scratch_numel(self):
        return int(self.lib.ntn_exchange_scratch_numel(self.h))

    def exchange_unbind(self):
        self._ck(self.lib.ntn_exchange_unbind(self.h))

    def exchange_probe(self, d_grad, off, n, last, iters):
        self._ck(self.lib.ntn_exchange_probe(self.h, d_grad, int(off), int(n), int(last), int(iters)))

    def exchange_check(self):
        return bool(self.lib.ntn_exchange_check(self.h))

    @property
    def exchange_mode(self):
        return self.lib.ntn_exchange_mode(self.h).decode()

    @staticmethod
    def unpack_tail(tail):
        """(cost, n_active) from the 4 summed tail floats."""
        t = [float(x) for x in tail]
        return t[0] + t[1], int(round(t[2])) * 4096 + int(round(t[3]))

    def result_tensor(self):
        """torch view (4 float64, on the device) of {cost, n_active, data term, sum theta^2}."""
        import torch

        class _Buf:
            pass
        b = _Buf()
        b.__cuda_array_interface__ = {"shape": (4,), "typestr": "<f8", "data": (int(self.lib.ntn_result_dev(self.h)), False),
                                      "version": 2, "strides": None}
        return torch.as_tensor(b, device="cuda")

    def to_internal_dev(self, d_ref: int, d_int: int):
        self._ck(self.lib.ntn_to_internal_dev(self.h, d_ref, d_int))

    def from_internal_dev(self, d_int: int, d_ref: int):
        self._ck(self.lib.ntn_from_internal_dev(self.h, d_int, d_ref))

    def upload_theta(self, theta, d_int: int):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        assert theta.size == self.P
        self._ck(self.lib.ntn_upload_theta(self.h, _ptr(theta), d_int))

    def download_theta(self, d_int: int, out=None):
        out = np.empty(self.P, np.float64) if out is None else out
        self._ck(self.lib.ntn_download_theta(self.h, d_int, _ptr(out)))
        return out

    def set_stream(self, cuda_stream: int):
        """Run on this cudaStream_t (0 = CUDA's legacy default stream, e.g. torch's default)."""
        self._ck(self.lib.ntn_set_stream(self.h, cuda_stream or None))

    def reset_stream(self):
        self._ck(self.lib.ntn_reset_stream(self.h))

    def score(self, theta, e1, rel, e2, mode="reference_eval"):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        e1, rel, e2 = _i32(e1), _i32(rel), _i32(e2)
        out = np.empty(len(e1), np.float64)
        m = {"reference_eval": SCORE_REFERENCE_EVAL, "train": SCORE_TRAIN}[mode]
        self._ck(self.lib.ntn_score(self.h, _ptr(theta), len(e1), _ptr(e1), _ptr(rel), _ptr(e2), m, _ptr(out)))
        return out

    def lbfgs_minimize(self, d_theta: int, draw_sides, max_iters=5, history=0, tol=0.0, reduce=None):
        """Device-resident L-BFGS on the current batch job; ``draw_sides()`` returns this evaluation's
        corrupt sides (uint8[R]), or ``draw_sides`` is the uint64 state array of the native drawer (see below).  Returns the LbfgsResult; theta is updated in place on the device.
        Data parallel: ``reduce(ptr, n)`` must SUM all-reduce the n device floats at ``ptr`` in place."""
        R = self.R
        ctx = None
        if isinstance(draw_sides, np.ndarray):
            # native drawer (ntn_host_sides_draw): uint64[2 + R] = {java.util.Random state, R, presence flags}; the
            # state is advanced in place -- no interpreter call per evaluation
            assert draw_sides.dtype == np.uint64 and draw_sides.size == 2 + R and int(draw_sides[1]) == R
            cb = C.cast(self.lib.ntn_host_sides_draw, SIDES_FN)
            ctx = _ptr(draw_sides)
        else:
            def _cb(_ctx, out):
                s = np.ascontiguousarray(draw_sides(), np.uint8)
                for r in range(R):
                    out[r] = int(s[r])
            cb = SIDES_FN(_cb)
        rcb = REDUCE_FN(lambda _ctx, ptr, n: reduce(int(ptr), int(n))) if reduce is not None else REDUCE_FN()
        opts, res = LbfgsOpts(max_iters, history, tol, rcb, None), LbfgsResult()
        self._ck(self.lib.ntn_lbfgs_minimize(self.h, d_theta, C.byref(opts), cb, ctx, C.byref(res)))
        return res

    def debug_fetch(self, name):
        n = self.lib.ntn_debug_fetch(self.h, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, np.float32)
        self.lib.ntn_debug_fetch(self.h, name.encode(), _ptr(out), n)
        return out.view(np.int32) if name.startswith("i:") else out

    # ---- introspection ----
    @property
    def launches(self):
        return self.lib.ntn_last_launch_count(self.h)

    @property
    def gemm_path(self):
        return {GEMM_SIMT: "simt", GEMM_TCGEN05: "tcgen05"}.get(self.lib.ntn_active_gemm_path(self.h), "?")

    def set_profiling(self, on=True):
        self._ck(self.lib.ntn_set_profiling(self.h, int(on)))

    def phase_ms(self):
        ms = np.zeros(NUM_PHASES, np.float32)
        self._ck(self.lib.ntn_last_phase_ms(self.h, _ptr(ms)))
        return {self.lib.ntn_phase_name(i).decode(): float(ms[i]) for i in range(NUM_PHASES)}
