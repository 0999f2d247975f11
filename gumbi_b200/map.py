"""``find_MAP`` for the B200 backend: host-side L-BFGS-B over device-evaluated objective and gradient.

Restates what ``PymcGP.find_MAP`` -> ``pm.find_MAP()`` does (gumbi/regression/pymc/GP.py:799-813): SciPy L-BFGS-B on
``-(log p(y|X,theta) + sum of log-priors)`` in the *transformed* space (positive variables are optimised through their
logs, ``jacobian=False`` so no Jacobian term is added), started from the model's initial point (prior moments;
``W`` from ``default_rng(seed).standard_normal`` as GP.py:459), ``maxeval=5000``.  The returned dict carries both the
constrained values and the ``*_log__`` keys, as PyMC's does.

Priors (cited per line): ls ~ InverseGamma(alpha, beta) from ``get_ls_prior`` (gumbi/utils/gp_utils.py:51-87) ->
``pm.find_constrained_prior`` restated with SciPy below; eta ~ Gamma(2, 1) (GP.py:408); c ~ Normal(0, 10),
tau ~ HalfNormal(10) (GP.py:451-452); W ~ Normal(0, 3), kappa ~ Gamma(1.5, 1) (GP.py:460-461); sigma ~ Exponential(1)
(GP.py:560).

Every O(N^2)-and-up quantity (K, Cholesky, K^-1, the N^2 gradient contraction) is computed on the GPU through
``gb2_factorize`` / ``gb2_mll_grad``; this file only does the O(#parameters) bookkeeping.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy import optimize, stats
from scipy.special import gammaln

POSITIVE = ("ls", "η", "τ", "κ", "σ")


# ----------------------------------------------------------------------------------------------------------------------
# length-scale prior (gp_utils.py:15-87)
# ----------------------------------------------------------------------------------------------------------------------
def _min_max_nonzero_distance(points):
    """min and max non-zero pairwise Euclidean distance.  1-D columns (the ARD case) are done by sorting, which gives
    exactly ``pdist``'s answer without the O(N^2) memory of gp_utils.py:34."""
    points = np.asarray(points, dtype=np.float64)
    if points.ndim == 1 or points.shape[1] == 1:
        u = np.unique(points.reshape(-1))
        if u.size < 2:
            return None, None
        return float(np.diff(u).min()), float(u[-1] - u[0])
    from scipy.spatial.distance import pdist

    n = points.shape[0]
    if n <= 4096:
        d = pdist(points)
        d = d[d != 0]
        return (float(d.min()), float(d.max())) if d.size else (None, None)
    from scipy.spatial.distance import cdist

    lo, hi = np.inf, 0.0
    for s in range(0, n, 1024):  # blocked, O(N * 1024) memory; direct differences (as pdist), so identical points give an exact 0
        d = cdist(points[s:s + 1024], points[s:])   # pairs (i, j >= i): every unordered pair once, plus the exact-zero self pairs
        nz = d[d != 0]
        if nz.size:
            lo, hi = min(lo, float(nz.min())), max(hi, float(nz.max()))
    return (lo, hi) if hi > 0 else (None, None)


def parse_ls_limits(X, ARD=True, lower=None, upper=None):
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    cols = [X[:, [j]] for j in range(X.shape[1])] if ARD else [X]

    def spread(v):
        v = [None] if v is None else list(np.atleast_1d(v))
        if len(v) == 1:
            v = v * len(cols)
        if len(v) != len(cols):
            raise ValueError("Number of bounds must match number of dimensions")
        return v

    lowers, uppers = spread(lower), spread(upper)
    for i, pts in enumerate(cols):
        dmin, dmax = _min_max_nonzero_distance(pts)
        default_lower = dmin if dmin is not None else 0.01
        lo = default_lower if lowers[i] is None else lowers[i]
        lowers[i] = max(lo, default_lower, 0.01)
        if uppers[i] is None:
            uppers[i] = dmax if dmax is not None else 1
    return lowers, uppers


def find_constrained_invgamma(lower, upper, mass=0.98, exact=False):
    """``pm.find_constrained_prior(pm.InverseGamma, lower, upper, init_guess={alpha: lower, beta: upper}, mass)``.

    PyMC poses: minimise (CDF(lower) - (1-mass)/2)^2 subject to CDF(upper) - CDF(lower) = mass, and hands it to
    ``scipy.optimize.minimize`` with a ``NonlinearConstraint`` (-> SLSQP) started at [lower, upper].  What the reference actually
    gets is NOT the exact optimum: the objective is at most 1e-4 in magnitude, so SLSQP meets the mass constraint in a few
    steps and declares convergence with CDF(lower) ~ 0 (e.g. lower=0.01, upper=3.5: alpha=4.07, beta=3.68, whereas the exact
    solution of both conditions is alpha=1.06, beta=0.047 -- a prior with its mode 30x lower).  Parity with the reference means
    reproducing that, so the same SciPy call is made here (``exact=True`` solves both conditions instead: beta is a pure scale,
    which leaves one monotone equation in alpha).  Raises ValueError('Optimization of parameters failed') like PyMC."""
    if not (0 < lower < upper) or not (0 < mass < 1):
        raise ValueError("Optimization of parameters failed.")
    tail = (1.0 - mass) / 2.0
    if not exact:
        def cdf(p, x):
            return stats.invgamma.cdf(x, p[0], scale=p[1]) if p[0] > 0 and p[1] > 0 else np.nan

        cons = optimize.NonlinearConstraint(lambda p: cdf(p, upper) - cdf(p, lower), lb=mass, ub=mass)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            opt = optimize.minimize(lambda p: (cdf(p, lower) - tail) ** 2, x0=[lower, upper], constraints=cons)
        if not opt.success or not np.all(np.isfinite(opt.x)) or np.any(opt.x <= 0):
            raise ValueError("Optimization of parameters failed.")
        return {"alpha": float(opt.x[0]), "beta": float(opt.x[1])}
    target = np.log(upper / lower)

    def ratio(log_a):
        a = np.exp(log_a)
        return np.log(stats.invgamma.ppf(1.0 - tail, a) / stats.invgamma.ppf(tail, a)) - target

    lo, hi = np.log(1e-2), np.log(1e7)
    try:
        f_lo, f_hi = ratio(lo), ratio(hi)
        if not (np.isfinite(f_lo) and np.isfinite(f_hi)) or f_lo * f_hi > 0:
            raise ValueError
        log_a = optimize.brentq(ratio, lo, hi, xtol=1e-12, rtol=1e-12)
    except ValueError:
        raise ValueError("Optimization of parameters failed.") from None
    alpha = float(np.exp(log_a))
    beta = float(lower / stats.invgamma.ppf(tail, alpha))
    return {"alpha": alpha, "beta": beta}


def get_ls_prior(X, ARD=True, lower=None, upper=None, mass=0.98):
    lowers, uppers = parse_ls_limits(X, ARD=ARD, lower=lower, upper=upper)
    alphas, betas = [], []
    for i, (lo, hi) in enumerate(zip(lowers, uppers)):
        mass_ = mass
        while True:
            try:
                p = find_constrained_invgamma(lo, hi, mass_)
                break
            except ValueError:
                mass_ -= 0.01  # gp_utils.py:72-74
                if mass_ <= 0.5:
                    raise
        if mass_ != mass:
            warnings.warn(f"Mass of constrained lengthscale prior was reduced from {mass:.3f} to {mass_:.3f} to enable "
                          f"convergence for dimension {i}.")
        alphas.append(p["alpha"])
        betas.append(p["beta"])
    return {"alpha": np.array(alphas), "beta": np.array(betas)}


# ----------------------------------------------------------------------------------------------------------------------
# priors: (logp, dlogp/dx, initial value)
# ----------------------------------------------------------------------------------------------------------------------
def _invgamma(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    init = np.where(a > 1, b / np.maximum(a - 1, 1e-300), b / (a + 1))
    return (lambda x: np.sum(a * np.log(b) - gammaln(a) - (a + 1) * np.log(x) - b / x),
            lambda x: -(a + 1) / x + b / x ** 2, lambda shape: np.broadcast_to(init, shape).copy())


def _gamma(a, b):
    return (lambda x: np.sum(a * np.log(b) - gammaln(a) + (a - 1) * np.log(x) - b * x),
            lambda x: (a - 1) / x - b, lambda shape: np.full(shape, a / b))


def _normal(mu, sd):
    return (lambda x: np.sum(-0.5 * ((x - mu) / sd) ** 2 - np.log(sd * np.sqrt(2 * np.pi))),
            lambda x: -(x - mu) / sd ** 2, lambda shape: np.full(shape, float(mu)))


def _halfnormal(sd):
    return (lambda x: np.sum(-0.5 * (x / sd) ** 2 + np.log(np.sqrt(2 / np.pi) / sd)),
            lambda x: -x / sd ** 2, lambda shape: np.full(shape, float(sd)))


def _exponential(lam):
    return (lambda x: np.sum(np.log(lam) - lam * x), lambda x: np.full(np.shape(x), -lam), lambda shape: np.full(shape, 1.0 / lam))


def ls_bounds_z(gp):
    """(lower, upper) standardized lengthscale bounds from ``gp.ls_bounds`` -- ``PymcGP._prepare_lengthscales`` (GP.py:630-646).

    ``ls_bounds`` is a ParameterArray with one layer per bounded continuous dimension holding [lower, upper] in natural units
    (NaN = unbounded); on plain arrays (``ArrayGP``) a dict {dim: (lower_z, upper_z)} already in standardized units.  Only the
    dimensions named in ``ls_bounds`` contribute, in ``continuous_dims`` order, and the reference's check
    ``not ARD and len(lower) != 1 or len(upper) != 1`` is kept as written (it rejects more than one bounded dimension)."""
    lb = getattr(gp, "ls_bounds", None)
    if lb is None:
        return None, None
    is_parray = hasattr(lb, "names")
    names = list(lb.names) if is_parray else list(lb)
    zbounds = []
    for dim in gp.continuous_dims:
        if dim in names:
            vals = np.asarray(lb[dim].z.values() if is_parray else lb[dim], dtype=np.float64).squeeze()
            zbounds.append([None if np.isnan(b) else float(b) for b in vals])
    lower, upper = list(zip(*zbounds))
    if not gp.ARD and len(lower) != 1 or len(upper) != 1:
        raise ValueError("Bounds must be specified for only a single dimension if ARD is False")
    return lower, upper


def build_priors(gp):
    """name -> (logp, dlogp, init) for every free hyper-parameter of ``gp`` (a built B200Backend)."""
    lay = gp._layout
    X = gp._X
    Xs = X[:, lay["idx_s"]]
    ls_prior = getattr(gp, "ls_prior", "InverseGamma")
    if ls_prior not in ("InverseGamma", "Gamma(2,1)"):
        raise ValueError(f"ls_prior must be 'InverseGamma' or 'Gamma(2,1)', got {ls_prior!r}")
    if ls_prior == "InverseGamma":     # today's reference: constrained InverseGamma from the data's distance range (GP.py:385,407)
        lower, upper = ls_bounds_z(gp)
        ls_params = get_ls_prior(Xs, ARD=gp.ARD, lower=lower, upper=upper, mass=gp.mass)
    pri = {}
    for name, shape in gp.param_shapes().items():
        kind = name.split("_")[0]
        if kind == "ls":
            # "Gamma(2,1)": the prior of the reference's older code, kept in its source as a comment (GP.py:408) -- the one the
            # executed Multioutput_Regression notebook was run with (tests/test_notebook_parity.py)
            pri[name] = _invgamma(ls_params["alpha"], ls_params["beta"]) if ls_prior == "InverseGamma" else _gamma(2.0, 1.0)
        elif kind == "η":
            pri[name] = _gamma(2.0, 1.0)
        elif kind == "c":
            pri[name] = _normal(0.0, 10.0)
        elif kind == "τ":
            pri[name] = _halfnormal(10.0)
        elif kind == "W":
            lp, dlp, _ = _normal(0.0, 3.0)
            seed = gp.seed
            pri[name] = (lp, dlp, lambda shape, seed=seed: np.random.default_rng(seed).standard_normal(size=shape))
        elif kind == "κ":
            pri[name] = _gamma(1.5, 1.0)
        elif kind == "σ":
            pri[name] = _exponential(1.0)
        else:  # pragma: no cover
            raise KeyError(name)
    return pri


# ----------------------------------------------------------------------------------------------------------------------
# gradient of the spec-shaped device gradient -> named hyper-parameters
# ----------------------------------------------------------------------------------------------------------------------
def named_gradient(gp, gspec):
    lay = gp._layout
    out = {}

    def add(name, val):
        out[name] = out.get(name, 0.0) + np.asarray(val, dtype=np.float64)

    for t, tg in zip(lay["terms"], gspec["terms"]):
        sfx = t["suffix"]
        add(f"ls_{sfx}", gp._fold_ls_gradient(tg["ls"]))
        add(f"η_{sfx}", tg["eta"])
        if lay["n_l"] > 0:
            add(f"c_{sfx}", tg["c"])
            add(f"τ_{sfx}", tg["tau"])
        for (name, _, _), cg in zip(t["coreg"], tg["coreg"]):
            add(f"W_{name}", cg["W"])       # the output Coregion is shared by all terms: contributions add up
            add(f"κ_{name}", cg["kappa"])
    add("σ", gspec["sigma"])
    if lay["noise_coreg"]:
        name = lay["noise_coreg"][0]
        add(f"W_{name}", gspec["noise_coreg"]["W"])
        add(f"κ_{name}", gspec["noise_coreg"]["kappa"])
    return out


def make_objective(gp, start=None):
    """Packed objective of ``pm.find_MAP``: returns (fun, x0, unpack, names, positive) where ``fun(x)`` is
    (-(log-likelihood + log-priors), its gradient) in the transformed space and ``unpack(x)`` the constrained point."""
    shapes = gp.param_shapes()
    pri = build_priors(gp)
    names = list(shapes)
    sizes = [int(np.prod(shapes[n])) if shapes[n] != () else 1 for n in names]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(int)
    positive = [n.split("_")[0] in POSITIVE for n in names]

    x0 = np.zeros(offs[-1])
    for i, n in enumerate(names):
        shape = shapes[n] if shapes[n] != () else (1,)
        val = pri[n][2](shape)
        if start and n in start:
            val = np.broadcast_to(np.asarray(start[n], dtype=np.float64), shape)
        x0[offs[i]:offs[i + 1]] = (np.log(val) if positive[i] else val).reshape(-1)

    def unpack(x):
        point = {}
        for i, n in enumerate(names):
            v = x[offs[i]:offs[i + 1]]
            v = np.exp(v) if positive[i] else v.copy()
            point[n] = v.reshape(shapes[n]) if shapes[n] != () else float(v[0])
        return point

    def lockstep(val, grad):
        """Multi-GPU: every rank runs this optimiser over a collective objective.  The device gradient is summed with fp64 atomics, so
        it is not bitwise identical between ranks; rank 0's (value, gradient) is broadcast so that all ranks take the same steps,
        stop at the same evaluation and never meet in different collectives."""
        if getattr(gp.engine, "world", 1) <= 1 and getattr(gp.engine, "kron_world", 1) <= 1:
            return val, grad
        import torch.distributed as tdist

        if not (tdist.is_available() and tdist.is_initialized()) or tdist.get_world_size() == 1:
            return val, grad
        payload = [(float(val), np.asarray(grad, dtype=np.float64)) if tdist.get_rank() == 0 else None]
        tdist.broadcast_object_list(payload, src=0)
        return payload[0][0], payload[0][1].copy()

    def fun(x):
        return lockstep(*fun_local(x))

    def fitc_value(xv):
        gp.engine.set_kernel(gp.spec_from_point(unpack(xv)))
        gp.engine.fitc_factorize(gp._engine_points(gp._Xu))
        return gp.engine.fitc_mll()

    def fun_sparse(x):
        """sparse=True: the FITC marginal likelihood (gb2_fitc_mll) with a central-difference gradient in the optimiser's own
        (log-transformed) coordinates -- 2 p + 1 device factorisations of the n_u x n_u systems per evaluation, each O(N n_u^2).
        (PyMC differentiates the same logp by reverse mode; the step 1e-5 keeps the gradient error ~1e-8 relative.)"""
        fun.n_eval += 1
        point = unpack(x)
        try:
            val = fitc_value(x)
            gx = np.zeros_like(x)
            h = 1e-5
            for k in range(len(x)):
                xp = x.copy(); xp[k] += h
                xm = x.copy(); xm[k] -= h
                gx[k] = (fitc_value(xp) - fitc_value(xm)) / (2 * h)
        except (np.linalg.LinAlgError, FloatingPointError):
            return 1e100, np.zeros_like(x)
        grad = gx
        for i, n in enumerate(names):
            xv = np.asarray(point[n], dtype=np.float64).reshape(-1)
            val += float(pri[n][0](np.asarray(point[n], dtype=np.float64)))
            gi = np.asarray(pri[n][1](xv), dtype=np.float64).reshape(-1)
            grad[offs[i]:offs[i + 1]] += gi * xv if positive[i] else gi
        if not np.isfinite(val):
            return 1e100, np.zeros_like(x)
        return -val, -grad

    def fun_local(x):
        if getattr(gp, "sparse", False):
            return fun_sparse(x)
        fun.n_eval += 1
        point = unpack(x)
        spec = gp.spec_from_point(point)
        try:
            gp.engine.set_kernel(spec)
            gp.engine.factorize()
            val, gspec = gp.engine.mll_grad(spec)
        except (np.linalg.LinAlgError, FloatingPointError):
            return 1e100, np.zeros_like(x)
        g = named_gradient(gp, gspec)
        grad = np.zeros_like(x)
        for i, n in enumerate(names):
            xv = np.asarray(point[n], dtype=np.float64).reshape(-1)
            val += float(pri[n][0](np.asarray(point[n], dtype=np.float64)))
            gi = np.asarray(g[n], dtype=np.float64).reshape(-1) + np.asarray(pri[n][1](xv), dtype=np.float64).reshape(-1)
            grad[offs[i]:offs[i + 1]] = gi * xv if positive[i] else gi  # d/dlog(x) = x d/dx ; jacobian=False
        if not np.isfinite(val):
            return 1e100, np.zeros_like(x)
        return -val, -grad

    fun.n_eval = 0
    return fun, x0, unpack, names, positive


def find_map(gp, start=None, method="L-BFGS-B", maxeval=5000, return_raw=False, progressbar=False, **kwargs):
    """Restatement of ``pm.find_MAP(start, method="L-BFGS-B", maxeval=5000)`` for a built ``B200Backend``."""
    fun, x0, unpack, names, positive = make_objective(gp, start)
    options = dict(kwargs.pop("options", {}) or {})
    options.setdefault("maxfun", maxeval)
    res = optimize.minimize(fun, x0, jac=True, method=method, options=options, **kwargs)
    point = unpack(res.x)
    MAP = {}
    for i, n in enumerate(names):
        val = np.asarray(point[n], dtype=np.float64)
        MAP[n] = val
        if positive[i]:
            MAP[n + "_log__"] = np.log(val)
    gp._factor_key = None
    gp.map_result = res
    gp.map_evals = fun.n_eval
    if return_raw:
        return MAP, res
    return MAP
