"""Kronecker-aware multi-output solve: P independent n x n problems instead of one (nP) x (nP) problem.

When every output is observed at the same n input rows (what ``Regressor.get_shaped_data`` stacks when no observation is
missing, gumbi/regression/base.py:459-464) and the model is not additive, the covariance the reference hands to
``pm.gp.Marginal`` (gumbi/regression/pymc/GP.py:724-727, :560-569, :580) is

    K  =  B (x) Kx  +  D (x) I_n ,      B = W W^T + diag(kappa)                    (output Coregion, P x P)
                                        D = diag(sigma^2 * Bn[p,p] + jitter)       (output noise [* Coregion "Output_noise"] + 1e-6)
                                        Kx = (eta^2 k(ls) + tau Linear(c)) o prod(other Coregion factors)   over the n rows

With the P x P eigen-decomposition  D^-1/2 B D^-1/2 = U diag(lam) U^T  and  T = D^-1/2 U  (T^T B T = diag(lam), T^T D T = I):

    K = (T^-T (x) I) blockdiag_q(lam_q Kx + I_n) (T^-1 (x) I)

so the Cholesky factorisation, the solves and the marginal likelihood split into P single-output problems
``K_q = lam_q Kx + I`` with rotated observations ``y~_q = sum_p T[p,q] y_p`` -- each one is exactly what the CUDA core already
does for a single-output model with ``eta_q = eta sqrt(lam_q)``, ``tau_q = tau lam_q``, ``sigma = 1``, ``jitter = 0``.  Cost
P n^3/3 + P n^2 M instead of (nP)^3/3 + (nP)^2 M P_pred: P^2 fewer flops in the factorisation (BASELINE config 5, P = 4: 16x),
the blocks fit one GPU each and need NO exchange when they are spread over GPUs (one process per GPU, blocks dealt round-robin).
The reference's BoTorch backend makes the same structural choice (``KroneckerMultiTaskGP`` when all inputs are shared,
gumbi/regression/botorch/GP.py:219-241).

Everything here is O(P^3 + M P^2) host arithmetic around block engines that expose the ``GPEngine`` interface; the numbers are
those of the dense path up to rounding (tests/test_kron.py checks 1e-9 against the dense oracle and the dense CUDA path):

    posterior at (x*, p*):  mean = sum_q a_q mu_q(x*),  var = sum_q a_q^2 var_q(x*)  (+ sigma^2 Bn[p*,p*]),   a = T^-1[:, p*]
    log p(y)             :  sum_q mll_q  -  n/2 sum_p log D_p
    gradient             :  ls, c, other Coregion factors: sum_q of the block gradients; eta: sum_q sqrt(lam_q) d/d eta_q;
                            tau: sum_q lam_q d/d tau_q;  B, D: from alpha~_q = K_q^-1 y~_q and tr K_q^-1 (see ``mll_grad``).
"""
from __future__ import annotations

import numpy as np


class NotAligned(ValueError):
    """The stacked training rows are not P aligned copies of the same inputs (or the model has no Kronecker structure)."""


def aligned_blocks(X, pcol: int, P: int):
    """(Xb, rows): ``Xb`` (n, D_in-1) the shared inputs without the output column, ``rows`` (P, n) indices of each output's rows
    in the stacked arrays -- or raises ``NotAligned``.  Rows must appear in the same order for every output, which is what
    base.py:459-464 produces (one copy of X per output, in coordinate order)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    N = X.shape[0]
    p = X[:, pcol]
    pi = p.astype(np.int64)
    if N == 0 or not np.array_equal(pi, p) or pi.min() < 0 or pi.max() >= P:
        raise NotAligned("output column must hold integer coordinates in [0, P)")
    if N % P:
        raise NotAligned(f"{N} stacked rows are not a multiple of {P} outputs (missing observations?)")
    n = N // P
    rows = [np.flatnonzero(pi == q) for q in range(P)]
    if any(len(r) != n for r in rows):
        raise NotAligned("outputs have different numbers of observations")
    rest = np.delete(X, pcol, axis=1)
    Xb = np.ascontiguousarray(rest[rows[0]])
    for r in rows[1:]:
        if not np.array_equal(rest[r], Xb):
            raise NotAligned("outputs are observed at different inputs")
    return Xb, np.stack(rows)


def split_spec(spec: dict, pcol: int):
    """Dense multi-output ``spec`` -> (block term without the output Coregion, index of that factor, B, D).  Raises
    ``NotAligned`` for structures without the Kronecker form (additive models: one output Coregion per term with different Kx)."""
    if len(spec["terms"]) != 1:
        raise NotAligned("additive models (several terms) have no single B (x) Kx form")
    term = spec["terms"][0]
    out_f = [f for f, cg in enumerate(term.get("coreg") or []) if cg["col"] == pcol]
    if len(out_f) != 1:
        raise NotAligned("the model has no Coregion factor over the output column")
    f_out = out_f[0]
    cg = term["coreg"][f_out]
    W = np.atleast_2d(np.asarray(cg["W"], dtype=np.float64))
    kappa = np.asarray(cg["kappa"], dtype=np.float64).reshape(-1)
    B = W @ W.T + np.diag(kappa)
    P = len(kappa)
    sigma, jitter = float(spec["sigma"]), float(spec.get("jitter", 1e-6))
    ncg = spec.get("noise_coreg")
    if ncg:
        if ncg["col"] != pcol:
            raise NotAligned("noise Coregion over a column other than the output column")
        Wn = np.atleast_2d(np.asarray(ncg["W"], dtype=np.float64))
        bn = (Wn ** 2).sum(1) + np.asarray(ncg["kappa"], dtype=np.float64).reshape(-1)
    else:
        bn = np.ones(P)
    D = sigma ** 2 * bn + jitter

    def shift(col):  # column index once the output column is removed
        return col - 1 if col > pcol else col

    block = {"kind": term["kind"], "cont_idx": [shift(c) for c in term["cont_idx"]], "ls": list(term["ls"]),
             "eta": float(term["eta"]), "lin_idx": [shift(c) for c in (term.get("lin_idx") or [])],
             "c": list(term.get("c") or []), "tau": float(term.get("tau") or 0.0),
             "coreg": [{"col": shift(c["col"]), "W": c["W"], "kappa": c["kappa"]}
                       for f, c in enumerate(term.get("coreg") or []) if f != f_out]}
    return block, f_out, B, D, bn


def rotation(B, D):
    """lam (P,), T (P, P), Tinv (P, P) with T^T B T = diag(lam), T^T diag(D) T = I."""
    s = 1.0 / np.sqrt(D)
    lam, U = np.linalg.eigh(B * s[:, None] * s[None, :])
    if lam.min() < -1e-12 * max(lam.max(), 1.0):
        raise np.linalg.LinAlgError("output Coregion matrix is not positive semi-definite")
    lam = np.maximum(lam, 1e-300)   # kappa > 0 keeps B positive definite; the floor only guards the divisions in mll_grad
    return lam, U * s[:, None], U.T / s[None, :]


class KronEngine:
    """Stands where ``GPEngine`` stands in ``B200Backend`` for an aligned multi-output model.

    ``make_engine()`` returns a fresh block engine (a ``GPEngine`` on this process's GPU).  ``rank``/``world``: blocks q with
    ``q % world == rank`` are factorised here; per-block results (O(n) numbers) are exchanged with ``gather`` -- a callable
    ``gather(obj) -> [obj_rank0, ...]`` (``torch.distributed.all_gather_object``), the only communication of this path.
    """

    def __init__(self, make_engine, pcol: int, P: int, rank: int = 0, world: int = 1, gather=None, threads: int = 1):
        if world > 1 and gather is None:
            raise ValueError("world > 1 needs a gather callable")
        # threads > 1: this rank's blocks are driven from that many host threads at once (every block engine owns its handle and
        # streams, the C ABI keeps no shared mutable state and ctypes drops the GIL), so that on one GPU the factorisation of one
        # block fills the SM slots another block's critical chain leaves idle.  Same arithmetic per block: identical results.
        self.threads = int(threads)
        self.make_engine = make_engine
        self.pcol, self.P = int(pcol), int(P)
        self.kron_rank, self.kron_world, self.gather = int(rank), int(world), gather
        self.mine = [q for q in range(self.P) if q % self.kron_world == self.kron_rank]
        self.blocks = {}
        self.spec = None
        self.factorized = False
        self.options = {}

    # -- GPEngine interface -------------------------------------------------------------------------------------------
    def set_train(self, X, y):
        self.Xb, self.rows = aligned_blocks(X, self.pcol, self.P)
        self.n = self.Xb.shape[0]
        self.Y = np.asarray(y, dtype=np.float64).reshape(-1)[self.rows]          # (P, n)
        self.N, self.D_in = self.n * self.P, self.Xb.shape[1] + 1
        self.factorized = False

    def set_kernel(self, spec):
        self.block_term, self.f_out, self.B, self.D, self.bn = split_spec(spec, self.pcol)
        self.spec = spec
        self.factorized = False

    def set_option(self, name, value):
        self.options[name] = int(value)
        for e in self.blocks.values():
            e.set_option(name, value)

    def block_spec(self, q):
        t = dict(self.block_term)
        t["eta"] = self.block_term["eta"] * np.sqrt(self.lam[q])
        t["tau"] = self.block_term["tau"] * self.lam[q]
        return {"terms": [t], "sigma": 1.0, "noise_coreg": None, "jitter": 0.0}

    def factorize(self, _predict_points=None):
        if self.spec is None:
            raise RuntimeError("factorize called before set_kernel")
        self.lam, self.T, self.Tinv = rotation(self.B, self.D)
        self.Yt = self.T.T @ self.Y                                               # (P, n): rotated observations
        for q in self.mine:
            if q not in self.blocks:
                self.blocks[q] = self.make_engine()
                for k, v in self.options.items():
                    self.blocks[q].set_option(k, v)

        def one(q):
            e = self.blocks[q]
            e.set_train(self.Xb, self.Yt[q])
            e.set_kernel(self.block_spec(q))
            if _predict_points is None:
                e.factorize()
                return None
            return e.factorize_predict(_predict_points, pred_noise=False)

        # A non-PD block on one rank must fail the call on EVERY rank: the ranks own different blocks, and a rank that raised
        # alone would skip the next gather while the others wait in it (and later pair with a stale one).
        def guarded(q):
            try:
                return ("ok", one(q))
            except np.linalg.LinAlgError as ex:
                return ("not_pd", str(ex))

        self.factorized = False
        local = self._each(guarded)
        verdicts = {q: (v[0], v[1] if v[0] != "ok" else None) for q, v in local.items()}
        parts = [verdicts] if self.kron_world == 1 else self.gather(verdicts)
        bad = sorted((q, v[1]) for part in parts for q, v in part.items() if v[0] != "ok")
        if bad:
            raise np.linalg.LinAlgError(f"Kronecker block {bad[0][0]}: {bad[0][1]}")
        self._block_posteriors = {q: v[1] for q, v in local.items()}
        self.factorized = True

    def factorize_predict(self, Xs, pred_noise: bool = True):
        """``factorize()`` + ``predict()`` with every block going through its engine's one-pass entry point (gb2_factorize_predict)."""
        if self.spec is None:
            raise RuntimeError("factorize_predict called before set_kernel")
        rest, pstar, uniq, inv = self._split_points(Xs)
        self.factorize(_predict_points=uniq)
        return self._combine(self._collect(self._block_posteriors), pstar, inv, pred_noise)

    def _each(self, fn):
        """{q: fn(q)} over this rank's blocks, on ``self.threads`` host threads when asked to."""
        if self.threads > 1 and len(self.mine) > 1:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=min(self.threads, len(self.mine))) as pool:
                return dict(zip(self.mine, pool.map(fn, self.mine)))
        return {q: fn(q) for q in self.mine}

    def _collect(self, local: dict):
        """{q: value} from every rank -> list indexed by q."""
        parts = [local] if self.kron_world == 1 else self.gather(local)
        merged = {}
        for part in parts:
            merged.update(part)
        return [merged[q] for q in range(self.P)]

    def mll(self):
        self._need_factor()
        vals = self._collect(self._each(lambda q: float(self.blocks[q].mll())))
        return float(np.sum(vals) - 0.5 * self.n * np.log(self.D).sum())

    def predict(self, Xs, pred_noise: bool = True):
        self._need_factor()
        rest, pstar, uniq, inv = self._split_points(Xs)
        res = self._collect(self._each(lambda q: self.blocks[q].predict(uniq, pred_noise=False)))
        return self._combine(res, pstar, inv, pred_noise)

    def _combine(self, res, pstar, inv, pred_noise):
        """Block posteriors (mu_q, var_q at the distinct inputs) -> posterior at the requested (input, output) pairs."""
        A = self.Tinv[:, pstar]                                                   # (P, M')
        mean = np.zeros(len(pstar))
        var = np.zeros(len(pstar))
        for q, (mu_q, var_q) in enumerate(res):
            mean += A[q] * np.asarray(mu_q)[inv]
            var += A[q] ** 2 * np.asarray(var_q)[inv]
        if pred_noise:
            var += float(self.spec["sigma"]) ** 2 * self.bn[pstar]
        return mean, var

    def predict_full(self, Xs, pred_noise: bool = False):
        self._need_factor()
        rest, pstar, uniq, inv = self._split_points(Xs)
        res = self._collect(self._each(lambda q: self.blocks[q].predict_full(uniq, pred_noise=False)))
        A = self.Tinv[:, pstar]
        mean = np.zeros(len(pstar))
        cov = np.zeros((len(pstar), len(pstar)))
        for q, (mu_q, cov_q) in enumerate(res):
            mean += A[q] * np.asarray(mu_q)[inv]
            cov += (A[q][:, None] * A[q][None, :]) * np.asarray(cov_q)[np.ix_(inv, inv)]
        if pred_noise:
            cov[np.diag_indices_from(cov)] += float(self.spec["sigma"]) ** 2 * self.bn[pstar]
        return mean, cov

    def mll_grad(self, spec):
        """(log p(y), gradient shaped like the dense ``spec``) -- same contract as ``GPEngine.mll_grad``.

        With G = 1/2 (alpha alpha^T - K^-1) the exact gradient w.r.t. the entries of K, alpha = (T (x) I) alpha~ and
        K^-1 = (T (x) I) blockdiag(K_q^-1) (T^T (x) I):
          dB[p,p'] = sum G o (E_pp' (x) Kx) = 1/2 [T (Ma - diag((n - tr K_q^-1)/lam_q)) T^T]_pp',   Ma[q,q'] = alpha~_q^T Kx alpha~_q'
          dD[p]    = sum G o (E_pp  (x) I ) = 1/2 [T (Aa - diag(tr K_q^-1)) T^T]_pp,                 Aa[q,q'] = alpha~_q^T alpha~_q'
        with Kx alpha~_q = (y~_q - alpha~_q)/lam_q and tr K_q^-1 = alpha~_q^T alpha~_q - d mll_q/d sigma_q (sigma_q = 1)."""
        self._need_factor()
        def one(q):
            val, g = self.blocks[q].mll_grad(self.block_spec(q))
            return float(val), g, np.asarray(self.blocks[q].get_alpha(), dtype=np.float64)

        res = self._collect(self._each(one))
        P, n, lam, T = self.P, self.n, self.lam, self.T
        val = float(sum(r[0] for r in res) - 0.5 * n * np.log(self.D).sum())
        alpha = np.stack([r[2] for r in res])                                     # (P, n) alpha~_q
        Aa = alpha @ alpha.T
        Ma = alpha @ ((self.Yt - alpha) / lam[:, None]).T
        Ma = 0.5 * (Ma + Ma.T)
        tr_inv = np.array([Aa[q, q] - float(res[q][1]["sigma"]) for q in range(P)])
        GB = 0.5 * (T @ (Ma - np.diag((n - tr_inv) / lam)) @ T.T)
        Gd = 0.5 * np.diag(T @ (Aa - np.diag(tr_inv)) @ T.T)

        term = spec["terms"][0]
        bt = [r[1]["terms"][0] for r in res]
        tg = {"ls": sum(np.asarray(b["ls"], dtype=np.float64) for b in bt),
              "eta": float(sum(np.sqrt(lam[q]) * bt[q]["eta"] for q in range(P))),
              "c": sum(np.asarray(b["c"], dtype=np.float64) for b in bt),
              "tau": float(sum(lam[q] * bt[q]["tau"] for q in range(P))), "coreg": []}
        fb = 0
        for f, cg in enumerate(term.get("coreg") or []):
            if f == self.f_out:
                W = np.atleast_2d(np.asarray(cg["W"], dtype=np.float64))
                tg["coreg"].append({"W": (GB + GB.T) @ W, "kappa": np.diag(GB).copy()})
            else:
                tg["coreg"].append({"W": sum(np.asarray(b["coreg"][fb]["W"]) for b in bt),
                                    "kappa": sum(np.asarray(b["coreg"][fb]["kappa"]) for b in bt)})
                fb += 1
        sigma = float(spec["sigma"])
        out = {"terms": [tg], "sigma": float(2.0 * sigma * (Gd * self.bn).sum()), "noise_coreg": None}
        ncg = spec.get("noise_coreg")
        if ncg:
            Wn = np.atleast_2d(np.asarray(ncg["W"], dtype=np.float64))
            gbn = Gd * sigma ** 2
            out["noise_coreg"] = {"W": 2.0 * gbn[:, None] * Wn, "kappa": gbn.copy()}
        return val, out

    def close(self):
        for e in self.blocks.values():
            if hasattr(e, "close"):
                e.close()
        self.blocks = {}
        self.factorized = False

    # -- helpers ------------------------------------------------------------------------------------------------------
    def _need_factor(self):
        if not self.factorized:
            raise RuntimeError("called before a successful factorize")

    def _split_points(self, Xs):
        Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
        if Xs.shape[1] != self.D_in:
            raise ValueError(f"points_array has {Xs.shape[1]} columns, model has {self.D_in} dims")
        if not np.all(np.isfinite(Xs)):
            raise ValueError("points_array must be finite")
        p = Xs[:, self.pcol]
        pstar = p.astype(np.int64)
        if not np.array_equal(pstar, p) or (len(p) and (pstar.min() < 0 or pstar.max() >= self.P)):
            raise ValueError("output coordinates of the prediction points must be integers in [0, P)")
        rest = np.delete(Xs, self.pcol, axis=1)
        # each distinct input row is solved once per block, whatever the number of outputs predicted there (base.py:533-536 tiles
        # the points once per output)
        uniq, inv = np.unique(rest, axis=0, return_inverse=True)
        return rest, pstar, np.ascontiguousarray(uniq), np.asarray(inv).reshape(-1)
