"""GPEngine: one fitted-model handle on one B200, over the C ABI.

Mirrors what ``pm.gp.Marginal`` does for ``PymcGP`` (gumbi/regression/pymc/GP.py:580, :845-847) but keeps the
factor ``L`` and ``v = L^-1 y`` resident in HBM between calls (the reference rebuilds and re-factorises on
every ``predict``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

TIMING_KEYS = ("prep_ms", "kbuild_ms", "cholesky_ms", "kstar_ms", "solve_ms", "reduce_ms", "launches_factorize", "launches_predict")


def _c_f64(a, ndim=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if ndim is not None and a.ndim != ndim:
        raise ValueError(f"expected a {ndim}-D array, got shape {a.shape}")
    return a


class GPEngine:
    """Opaque handle owner.  ``precision`` is "fp64" (default) or "tf32"."""

    def __init__(self, device: int = 0, precision: str = "fp64"):
        self._lib = _lib.load()
        if precision not in ("fp64", "tf32"):
            raise ValueError('precision must be "fp64" or "tf32"')
        self.precision = precision
        self.device = int(device)
        h = C.c_void_p()
        rc = self._lib.gb2_create(C.byref(h), self.device, _lib.FP64 if precision == "fp64" else _lib.TF32)
        if rc != 0:
            msg = self._lib.gb2_last_error(None).decode()
            raise _lib.BackendUnavailable(f"gb2_create failed ({rc}): {msg}")
        self._h = h
        self._keep = None
        self.N = 0
        self.D_in = 0
        self.rank, self.world = 0, 1
        self.options = {}

    # -- lifetime ------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gb2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc == 0:
            return
        msg = self._lib.gb2_last_error(self._h).decode()
        if rc > 0:
            # same exception type the reference surfaces from scipy/pytensor's Cholesky
            raise np.linalg.LinAlgError(f"{what}: {msg}")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")

    # -- data / hyper-parameters ----------------------------------------------------------------------
    def set_train(self, X, y):
        X = _c_f64(np.atleast_2d(X), 2)
        y = _c_f64(y).reshape(-1)
        if X.shape[0] != y.shape[0]:
            raise ValueError(f"X has {X.shape[0]} rows but y has {y.shape[0]} entries")
        if not (np.all(np.isfinite(X)) and np.all(np.isfinite(y))):
            raise ValueError("X and y must be finite (get_shaped_data drops NaN rows, base.py:469-471)")
        self._check(self._lib.gb2_set_train(self._h, _lib.as_dp(X), X.shape[0], X.shape[1], _lib.as_dp(y)), "set_train")
        self.N, self.D_in = X.shape

    def set_train_device(self, dX_ptr: int, N: int, D_in: int, dy_ptr: int):
        """Device-resident inputs (e.g. ``torch.Tensor.data_ptr()`` of float64 C-contiguous CUDA tensors)."""
        self._check(self._lib.gb2_set_train_dev(self._h, C.c_void_p(dX_ptr), N, D_in, C.c_void_p(dy_ptr)), "set_train_dev")
        self.N, self.D_in = int(N), int(D_in)

    def set_kernel(self, spec: dict):
        k, keep = _lib.make_kernel_struct(spec)
        self._check(self._lib.gb2_set_kernel(self._h, C.byref(k)), "set_kernel")
        self._keep = (k, keep)

    def set_option(self, name: str, value: int):
        self._check(self._lib.gb2_set_option(self._h, name.encode(), int(value)), "set_option")
        self.options[name] = int(value)

    @property
    def shard_storage(self) -> bool:
        """True: the factor is stored row-block-sharded across the ranks and predict() is a collective over ALL points."""
        return self.world > 1 and bool(self.options.get("shard_storage"))

    # -- multi-GPU ------------------------------------------------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        rc = self._lib.gb2_nccl_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"gb2_nccl_unique_id failed ({rc}): {self._lib.gb2_last_error(None).decode()}")
        return buf.raw

    def dist_init(self, rank: int, world: int, unique_id: bytes):
        """Collective over all ranks: afterwards ``factorize`` shards K by 128-row blocks across the ranks' GPUs."""
        if world > 1 and len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of an ncclUniqueId")
        self._check(self._lib.gb2_dist_init(self._h, int(rank), int(world), unique_id), "dist_init")
        self.rank, self.world = int(rank), int(world)

    def allgather_device(self, dsend_ptr: int, drecv_ptr: int, count: int):
        """Rank-major gather of ``count`` float64 per rank on the handle's stream (device pointers)."""
        self._check(self._lib.gb2_dist_allgather_dev(self._h, C.c_void_p(dsend_ptr), C.c_void_p(drecv_ptr), int(count)), "allgather")

    def dist_finalize(self):
        self._check(self._lib.gb2_dist_finalize(self._h), "dist_finalize")
        self.rank, self.world = 0, 1

    # -- compute --------------------------------------------------------------------------------------
    def factorize(self):
        self._check(self._lib.gb2_factorize(self._h), "factorize")

    def mll(self) -> float:
        out = C.c_double()
        self._check(self._lib.gb2_mll(self._h, C.byref(out)), "mll")
        return out.value

    def mll_grad(self, spec: dict):
        """(log p(y|X,theta), gradient) with the gradient shaped like ``spec``: per term ``ls``, ``eta``, ``c``, ``tau`` and
        per Coregion factor ``W``/``kappa`` (chain rule through B = W W^T + diag(kappa) applied here), plus ``sigma`` and
        the noise Coregion.  ``spec`` must be the one last passed to ``set_kernel``."""
        L = _lib
        val = C.c_double()
        g = np.zeros(L.GRAD_LEN, dtype=np.float64)
        self._check(self._lib.gb2_mll_grad(self._h, C.byref(val), L.as_dp(g)), "mll_grad")

        def coreg_grad(cg, GB):
            W = np.atleast_2d(np.asarray(cg["W"], dtype=np.float64))
            return {"W": (GB + GB.T) @ W, "kappa": np.diag(GB).copy()}

        out = {"terms": [], "sigma": float(g[L.GRAD_SIGMA]), "noise_coreg": None}
        for t, term in enumerate(spec["terms"]):
            gt = g[t * L.GRAD_TERM:(t + 1) * L.GRAD_TERM]
            d = len(term["cont_idx"])
            ls = np.atleast_1d(np.asarray(term["ls"], dtype=np.float64))
            gls = gt[L.GRAD_LS:L.GRAD_LS + d].copy()
            if ls.size == 1 and d > 1:  # shared lengthscale: sum over dimensions
                gls = np.array([gls.sum()])
            n_lin = len(term.get("lin_idx") or [])
            tg = {"ls": gls, "eta": float(gt[L.GRAD_ETA]), "c": gt[L.GRAD_C:L.GRAD_C + n_lin].copy(),
                  "tau": float(gt[L.GRAD_TAU]), "coreg": []}
            for f, cg in enumerate(term.get("coreg") or []):
                P = len(cg["kappa"])
                GB = gt[L.GRAD_B + f * L.MAX_P ** 2:L.GRAD_B + f * L.MAX_P ** 2 + P * P].reshape(P, P)
                tg["coreg"].append(coreg_grad(cg, GB))
            out["terms"].append(tg)
        ncg = spec.get("noise_coreg")
        if ncg:
            P = len(ncg["kappa"])
            GB = g[L.GRAD_NOISE_B:L.GRAD_NOISE_B + P * P].reshape(P, P)
            out["noise_coreg"] = coreg_grad(ncg, GB)
        return val.value, out

    def predict(self, Xs, pred_noise: bool = True):
        Xs = _c_f64(np.atleast_2d(Xs), 2)
        if Xs.shape[1] != self.D_in:
            raise ValueError(f"points_array has {Xs.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xs)):
            raise ValueError("points_array must be finite")
        M = Xs.shape[0]
        mean = np.empty(M, dtype=np.float64)
        var = np.empty(M, dtype=np.float64)
        self._check(self._lib.gb2_predict(self._h, _lib.as_dp(Xs), M, int(bool(pred_noise)), _lib.as_dp(mean), _lib.as_dp(var)), "predict")
        return mean, var

    def factorize_predict(self, Xs, pred_noise: bool = True):
        """``factorize()`` + ``predict()`` in one pass (one cold reference ``predict`` call): the prediction points ride through the
        factorisation as extra rows of the factor.  fp64, replicated storage; leaves the handle factorised."""
        Xs = _c_f64(np.atleast_2d(Xs), 2)
        if Xs.shape[1] != self.D_in:
            raise ValueError(f"points_array has {Xs.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xs)):
            raise ValueError("points_array must be finite")
        M = Xs.shape[0]
        mean = np.empty(M, dtype=np.float64)
        var = np.empty(M, dtype=np.float64)
        self._check(self._lib.gb2_factorize_predict(self._h, _lib.as_dp(Xs), M, int(bool(pred_noise)), _lib.as_dp(mean), _lib.as_dp(var)),
                    "factorize_predict")
        return mean, var

    def predict_full(self, Xs, pred_noise: bool = False):
        """Posterior mean (M,) and full covariance (M, M) -- the parameters of ``gp.conditional(name, Xnew)`` (GP.py:913)."""
        Xs = _c_f64(np.atleast_2d(Xs), 2)
        if Xs.shape[1] != self.D_in:
            raise ValueError(f"points_array has {Xs.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xs)):
            raise ValueError("points_array must be finite")
        M = Xs.shape[0]
        mean = np.empty(M, dtype=np.float64)
        cov = np.empty((M, M), dtype=np.float64)
        self._check(self._lib.gb2_predict_full(self._h, _lib.as_dp(Xs), M, int(bool(pred_noise)), _lib.as_dp(mean), _lib.as_dp(cov)), "predict_full")
        return mean, cov

    # -- sparse FITC approximation (gb2_fitc_*; pm.gp.MarginalSparse(approx="FITC"), GP.py:571-578) ------------------------
    def fitc_factorize(self, Xu):
        """Factorise the FITC approximation with inducing points ``Xu`` (m, D_in) for the current training set and kernel."""
        Xu = _c_f64(np.atleast_2d(Xu), 2)
        if Xu.shape[1] != self.D_in:
            raise ValueError(f"Xu has {Xu.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xu)):
            raise ValueError("Xu must be finite")
        self._check(self._lib.gb2_fitc_factorize(self._h, _lib.as_dp(Xu), Xu.shape[0]), "fitc_factorize")

    def fitc_mll(self) -> float:
        out = C.c_double(0.0)
        self._check(self._lib.gb2_fitc_mll(self._h, C.byref(out)), "fitc_mll")
        return float(out.value)

    def fitc_predict(self, Xs, pred_noise: bool = True):
        Xs = _c_f64(np.atleast_2d(Xs), 2)
        if Xs.shape[1] != self.D_in:
            raise ValueError(f"points_array has {Xs.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xs)):
            raise ValueError("points_array must be finite")
        M = Xs.shape[0]
        mean = np.empty(M, dtype=np.float64)
        var = np.empty(M, dtype=np.float64)
        self._check(self._lib.gb2_fitc_predict(self._h, _lib.as_dp(Xs), M, int(bool(pred_noise)), _lib.as_dp(mean), _lib.as_dp(var)), "fitc_predict")
        return mean, var

    def predict_device(self, dXs_ptr: int, M: int, pred_noise: bool, dmean_ptr: int, dvar_ptr: int):
        self._check(
            self._lib.gb2_predict_dev(self._h, C.c_void_p(dXs_ptr), int(M), int(bool(pred_noise)), C.c_void_p(dmean_ptr), C.c_void_p(dvar_ptr)),
            "predict_dev",
        )

    def factorize_predict_device(self, dXs_ptr: int, M: int, pred_noise: bool, dmean_ptr: int, dvar_ptr: int):
        """``factorize_predict`` on device pointers (no host copies)."""
        self._check(
            self._lib.gb2_factorize_predict_dev(self._h, C.c_void_p(dXs_ptr), int(M), int(bool(pred_noise)), C.c_void_p(dmean_ptr),
                                                C.c_void_p(dvar_ptr)),
            "factorize_predict_dev",
        )

    def mark(self, slot: int):
        """Record a timing mark on the handle's stream (device-side timing for benchmarks)."""
        self._check(self._lib.gb2_mark(self._h, int(slot)), "mark")

    def elapsed_ms(self, a: int, b: int) -> float:
        out = C.c_double()
        self._check(self._lib.gb2_elapsed_ms(self._h, int(a), int(b), C.byref(out)), "elapsed_ms")
        return out.value

    # -- test hooks -----------------------------------------------------------------------------------
    def get_K(self):
        K = np.empty((self.N, self.N), dtype=np.float64)
        self._check(self._lib.gb2_get_K(self._h, _lib.as_dp(K)), "get_K")
        return K

    def get_L(self):
        L = np.empty((self.N, self.N), dtype=np.float64)
        self._check(self._lib.gb2_get_L(self._h, _lib.as_dp(L)), "get_L")
        return L

    def get_v(self):
        v = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.gb2_get_v(self._h, _lib.as_dp(v)), "get_v")
        return v

    def get_alpha(self):
        """alpha = K^-1 y of the last ``mll_grad`` call (valid until the next factorisation)."""
        a = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.gb2_get_alpha(self._h, _lib.as_dp(a)), "get_alpha")
        return a

    def get_trace(self):
        """(steps, 6) uint64 %globaltimer stamps (ns) of the last factorisation, after ``set_option("trace", 1)``; columns: diagonal
        kernel eligible / done, panel solve done, next-column update done (panel stream), bulk update eligible / done (main stream)."""
        steps = (self.N + 1 + 127) // 128
        out = np.zeros(steps * 6, dtype=np.uint64)
        self._check(self._lib.gb2_get_trace(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size), "get_trace")
        return out.reshape(steps, 6)

    def timings(self) -> dict:
        out = np.zeros(_lib.N_TIMINGS, dtype=np.float64)
        self._check(self._lib.gb2_get_timings(self._h, _lib.as_dp(out)), "get_timings")
        return dict(zip(TIMING_KEYS, out.tolist()))
