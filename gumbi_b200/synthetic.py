"""Synthetic workloads of SURVEY 8(d) / BASELINE.json configs 2-5 -- inputs only, shared by bench.py and the tests."""
from __future__ import annotations

import numpy as np


def synthetic_problem(n, d, P=1, M_res=100, kind="ExpQuad", seed=2021, sigma=0.1, Q=1):
    """BASELINE configs 2-5 inputs: X ~ N(0,1), y_p = sum_j sin(w_pj x_j) + 0.1 eps (z-scored per output),
    ls_j = (1 + 0.25 j) sqrt(d), eta = 1, W = default_rng(seed).standard_normal((P,2)) (GP.py:459), kappa = 1,
    noise coregion W_n = 0, kappa_n = 1; grid: gumbi-style prepare_grid over the first two dims at resolution
    M_res (limits [-2.2, 2.2]*1.1-padded as base.py:649-655 gives for data within +-2), other dims pinned at 0,
    tiled once per output.  Returns (spec, X, y, Xnew)."""
    rng = np.random.default_rng(seed)
    Xc = rng.standard_normal((n, d))
    w = rng.uniform(0.5, 2.0, size=(P, d))
    ys = []
    for p in range(P):
        yp = np.sin(Xc * w[p]).sum(1) + 0.1 * rng.standard_normal(n)
        ys.append((yp - yp.mean()) / yp.std(ddof=1))
    ls = (1.0 + 0.25 * np.arange(d)) * np.sqrt(d)
    lo = np.minimum(Xc.min(0), -2.0)
    hi = np.maximum(Xc.max(0), 2.0)
    pad = (hi - lo) * 0.1
    lo, hi = lo - pad, hi + pad
    if d >= 2:
        g0 = np.linspace(lo[0], hi[0], M_res)
        g1 = np.linspace(lo[1], hi[1], M_res)
        G0, G1 = np.meshgrid(g0, g1, indexing="ij")
        grid = np.zeros((M_res * M_res, d))
        grid[:, 0] = G0.ravel()
        grid[:, 1] = G1.ravel()
    else:
        grid = np.linspace(lo[0], hi[0], M_res)[:, None]
    terms = []
    for q in range(Q):
        term = {"kind": kind, "cont_idx": list(range(d)), "ls": list(ls * (1.0 + 0.5 * q)), "eta": 1.0 / np.sqrt(Q),
                "lin_idx": [], "c": [], "tau": 0.0, "coreg": []}
        if P > 1:
            W = np.random.default_rng(seed + q).standard_normal((P, 2))
            term["coreg"] = [{"col": d, "W": W.tolist(), "kappa": [1.0] * P}]
        terms.append(term)
    spec = {"terms": terms, "sigma": sigma, "noise_coreg": None, "jitter": 1e-6}
    if P > 1:
        X = np.vstack([np.column_stack([Xc, np.full(n, float(p))]) for p in range(P)])
        y = np.hstack(ys)
        Xnew = np.vstack([np.column_stack([grid, np.full(len(grid), float(p))]) for p in range(P)])
        spec["noise_coreg"] = {"col": d, "W": np.zeros((P, 2)).tolist(), "kappa": [1.0] * P}
    else:
        X, y, Xnew = Xc, ys[0], grid
    return spec, np.ascontiguousarray(X), np.ascontiguousarray(y), np.ascontiguousarray(Xnew)
