"""Host side of the sparse FITC path: inducing points for ``sparse=True`` (gumbi/regression/pymc/GP.py:571-578).

The reference calls ``pm.gp.util.kmeans_inducing_points(n_u, X)`` on the full shaped ``X`` (continuous AND categorical columns)
and hands the centroids to ``pm.gp.MarginalSparse(approx="FITC")``.  k-means is O(N n_u) host work per ``build_model``; everything
O(N n_u^2) (K(X,Xu), its triangular solve, the n_u x n_u systems, prediction) runs on the device through ``gb2_fitc_*``.
"""
from __future__ import annotations

import numpy as np


def kmeans_inducing_points(n_inducing, X, seed=None):
    """``pm.gp.util.kmeans_inducing_points``: whiten every column by its standard deviation (columns with std <= 1e-6 are left
    alone), run ``scipy.cluster.vq.kmeans`` with ``k_or_guess=n_inducing``, scale the centroids back.

    The reference leaves SciPy's k-means unseeded (a different ``Xu`` on every ``build_model``); here ``seed`` (the model seed) makes it
    repeatable.  Like SciPy, fewer than ``n_inducing`` centroids come back when clusters run empty."""
    from scipy.cluster.vq import kmeans

    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    scaling = np.std(X, 0)
    scaling[scaling <= 1e-6] = 1.0
    Xu, _ = kmeans(X / scaling, int(n_inducing), seed=seed)
    return np.ascontiguousarray(Xu * scaling)
