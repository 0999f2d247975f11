"""ctypes binding of ``libgumbi_b200.so`` (the C ABI declared in ``include/gumbi_b200.h``).

This is the whole reference-side binding: a maintainer of Gumbi would vendor this file next to
``gumbi/regression/b200/GP.py``.  There is deliberately no fallback: if the shared
library is missing or no B200 is visible, importing the symbols works but every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

MAX_TERMS, MAX_D, MAX_LIN, MAX_COREG, MAX_P = 4, 16, 8, 3, 16
N_TIMINGS = 8
# flat gradient layout of gb2_mll_grad (include/gumbi_b200.h)
GRAD_LS, GRAD_ETA, GRAD_C, GRAD_TAU, GRAD_B = 0, MAX_D, MAX_D + 1, MAX_D + 1 + MAX_LIN, MAX_D + 2 + MAX_LIN
GRAD_TERM = GRAD_B + MAX_COREG * MAX_P * MAX_P
GRAD_SIGMA = MAX_TERMS * GRAD_TERM
GRAD_NOISE_B = GRAD_SIGMA + 1
GRAD_LEN = GRAD_NOISE_B + MAX_P * MAX_P
KIND_IDS = {"ExpQuad": 0, "Matern52": 1, "Matern32": 2, "Matern12": 3, "Exponential": 4}
FP64, TF32 = 0, 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgumbi_b200.so")


class Term(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("d", C.c_int32),
        ("cont_idx", C.c_int32 * MAX_D),
        ("ls", C.c_double * MAX_D),
        ("eta", C.c_double),
        ("n_lin", C.c_int32),
        ("lin_idx", C.c_int32 * MAX_LIN),
        ("c", C.c_double * MAX_LIN),
        ("tau", C.c_double),
        ("n_coreg", C.c_int32),
        ("coreg_col", C.c_int32 * MAX_COREG),
        ("coreg_P", C.c_int32 * MAX_COREG),
        ("coreg_B", C.POINTER(C.c_double) * MAX_COREG),
    ]


class Kernel(C.Structure):
    _fields_ = [
        ("n_terms", C.c_int32),
        ("terms", Term * MAX_TERMS),
        ("sigma", C.c_double),
        ("noise_col", C.c_int32),
        ("noise_P", C.c_int32),
        ("noise_B", C.POINTER(C.c_double)),
        ("jitter", C.c_double),
    ]


EXPORTS = [
    "gb2_abi_version", "gb2_create", "gb2_destroy", "gb2_last_error", "gb2_set_train", "gb2_set_train_dev",
    "gb2_set_kernel", "gb2_factorize", "gb2_mll", "gb2_mll_grad", "gb2_get_alpha", "gb2_predict", "gb2_predict_dev", "gb2_predict_full", "gb2_fitc_factorize", "gb2_fitc_mll", "gb2_fitc_predict", "gb2_factorize_predict", "gb2_factorize_predict_dev", "gb2_get_K", "gb2_get_L",
    "gb2_get_v", "gb2_get_trace", "gb2_get_timings", "gb2_set_option", "gb2_mark", "gb2_elapsed_ms",
    "gb2_nccl_unique_id", "gb2_dist_init", "gb2_dist_finalize", "gb2_dist_allgather_dev",
]

_lib = None


class BackendUnavailable(RuntimeError):
    """The CUDA core cannot be used (library not built, or no B200 visible).  There is no CPU fallback."""


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library and declare the prototypes.  Raises BackendUnavailable if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise BackendUnavailable(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  gumbi_b200 has no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    H = C.c_void_p
    dp = C.POINTER(C.c_double)
    lib.gb2_abi_version.restype = C.c_int
    lib.gb2_create.argtypes = [C.POINTER(H), C.c_int, C.c_int]
    lib.gb2_destroy.argtypes = [H]
    lib.gb2_last_error.argtypes = [H]
    lib.gb2_last_error.restype = C.c_char_p
    lib.gb2_set_train.argtypes = [H, dp, C.c_int64, C.c_int32, dp]
    lib.gb2_set_train_dev.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.gb2_set_kernel.argtypes = [H, C.POINTER(Kernel)]
    lib.gb2_factorize.argtypes = [H]
    lib.gb2_mll.argtypes = [H, dp]
    lib.gb2_mll_grad.argtypes = [H, dp, dp]
    lib.gb2_predict.argtypes = [H, dp, C.c_int64, C.c_int32, dp, dp]
    lib.gb2_predict_full.argtypes = [H, dp, C.c_int64, C.c_int32, dp, dp]
    lib.gb2_predict_dev.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gb2_factorize_predict_dev.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    lib.gb2_fitc_factorize.argtypes = [H, dp, C.c_int64]
    lib.gb2_fitc_mll.argtypes = [H, dp]
    lib.gb2_fitc_predict.argtypes = [H, dp, C.c_int64, C.c_int32, dp, dp]
    lib.gb2_get_K.argtypes = [H, dp]
    lib.gb2_get_L.argtypes = [H, dp]
    lib.gb2_get_v.argtypes = [H, dp]
    lib.gb2_get_alpha.argtypes = [H, dp]
    lib.gb2_get_trace.argtypes = [H, C.POINTER(C.c_uint64), C.c_int64]
    lib.gb2_factorize_predict.argtypes = [H, dp, C.c_int64, C.c_int32, dp, dp]
    lib.gb2_get_timings.argtypes = [H, dp]
    lib.gb2_set_option.argtypes = [H, C.c_char_p, C.c_int]
    lib.gb2_nccl_unique_id.argtypes = [C.c_char_p]
    lib.gb2_dist_init.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    lib.gb2_dist_finalize.argtypes = [H]
    lib.gb2_dist_allgather_dev.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64]
    lib.gb2_mark.argtypes = [H, C.c_int]
    lib.gb2_elapsed_ms.argtypes = [H, C.c_int, C.c_int, dp]
    for name in EXPORTS:
        if name not in ("gb2_last_error",):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def as_dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_kernel_struct(spec: dict):
    """Plain-dict model description (see gumbi_b200.spec / oracle docstring) -> (Kernel struct, keep-alive list)."""
    keep = []
    k = Kernel()
    terms = spec["terms"]
    if not 1 <= len(terms) <= MAX_TERMS:
        raise ValueError(f"between 1 and {MAX_TERMS} additive terms are supported, got {len(terms)}")
    k.n_terms = len(terms)
    for t, term in enumerate(terms):
        T = k.terms[t]
        if term["kind"] not in KIND_IDS:
            raise ValueError(f"Continuous kernel must be one of {sorted(KIND_IDS)}, got {term['kind']!r}")
        T.kind = KIND_IDS[term["kind"]]
        ci = list(term["cont_idx"])
        ls = np.atleast_1d(np.asarray(term["ls"], dtype=np.float64))
        if ls.size == 1 and len(ci) > 1:  # ARD=False: one shared lengthscale (GP.py:400)
            ls = np.repeat(ls, len(ci))
        if len(ci) > MAX_D:
            raise ValueError(f"at most {MAX_D} continuous dimensions are supported")
        if ls.size != len(ci):
            raise ValueError("ls must have one entry per continuous dimension (or a single shared value)")
        T.d = len(ci)
        for i, (c_, l_) in enumerate(zip(ci, ls)):
            T.cont_idx[i] = int(c_)
            T.ls[i] = float(l_)
        T.eta = float(term["eta"])
        li = list(term.get("lin_idx") or [])
        if len(li) > MAX_LIN:
            raise ValueError(f"at most {MAX_LIN} linear dimensions are supported")
        T.n_lin = len(li)
        cc = np.atleast_1d(np.asarray(term.get("c") if len(li) else [], dtype=np.float64))
        if len(li) and cc.size != len(li):
            raise ValueError("c must have one entry per linear dimension")
        for i, c_ in enumerate(li):
            T.lin_idx[i] = int(c_)
            T.c[i] = float(cc[i])
        T.tau = float(term.get("tau", 0.0) or 0.0)
        cgs = list(term.get("coreg") or [])
        if len(cgs) > MAX_COREG:
            raise ValueError(f"at most {MAX_COREG} Coregion factors per term are supported")
        T.n_coreg = len(cgs)
        for f, cg in enumerate(cgs):
            B = coregion_B(cg["W"], cg["kappa"])
            if B.shape[0] > MAX_P:
                raise ValueError(f"at most {MAX_P} levels per Coregion factor are supported")
            keep.append(B)
            T.coreg_col[f] = int(cg["col"])
            T.coreg_P[f] = B.shape[0]
            T.coreg_B[f] = as_dp(B)
    k.sigma = float(spec["sigma"])
    k.jitter = float(spec.get("jitter", 1e-6))
    ncg = spec.get("noise_coreg")
    if ncg:
        B = coregion_B(ncg["W"], ncg["kappa"])
        keep.append(B)
        k.noise_col = int(ncg["col"])
        k.noise_P = B.shape[0]
        k.noise_B = as_dp(B)
    else:
        k.noise_col = -1
        k.noise_P = 0
    return k, keep


def coregion_B(W, kappa) -> np.ndarray:
    """B = W W^T + diag(kappa)  (pm.gp.cov.Coregion, gumbi/regression/pymc/GP.py:457-464); tiny, host side."""
    W = np.atleast_2d(np.asarray(W, dtype=np.float64))
    kappa = np.atleast_1d(np.asarray(kappa, dtype=np.float64))
    if W.shape[0] != kappa.shape[0]:
        raise ValueError("W and kappa disagree on the number of levels")
    return np.ascontiguousarray(W @ W.T + np.diag(kappa))
