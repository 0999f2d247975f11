"""Sparse FITC path (gumbi/regression/pymc/GP.py:571-578, :585-602).

CPU: the oracle's Woodbury restatement of ``pm.gp.MarginalSparse(approx="FITC")`` against the dense definition of the FITC model
(y ~ N(0, Qff + diag(Kff - Qff) + sigma^2 I), standard FITC predictive), k-means inducing points, and the backend's host logic
(``sparse=True``: layout, MAP keys, finite-difference objective) on the oracle-backed test double.
GPU: ``gb2_fitc_*`` through the C ABI against the oracle."""
import warnings

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import gp_oracle as orc
from test_backend_host import HostGP


def problem(n=500, d=3, kind="Matern52", m=40, seed=3, P=1):
    spec, X, y, Xs = orc.synthetic_problem(n, d, P=P, M_res=9, kind=kind)
    spec["noise_coreg"] = None
    Xu = orc.kmeans_inducing_points(m, X, seed=seed)
    return spec, X, y, Xs, Xu


def dense_fitc(spec, X, y, Xu, Xs, pred_noise=True):
    """Textbook FITC (Snelson & Ghahramani 2006) written with dense N x N algebra -- independent of the Woodbury form."""
    jit = spec.get("jitter", 1e-6)
    Kuu = orc.cov_full(spec, Xu) + jit * np.eye(len(Xu))
    Kuf = orc.cov_full(spec, Xu, X)
    Qff = Kuf.T @ np.linalg.solve(Kuu, Kuf)
    lam = np.clip(orc.cov_diag(spec, X) - np.diag(Qff), 0, np.inf) + spec["sigma"] ** 2
    C = Qff - np.diag(np.diag(Qff)) + np.diag(np.diag(Qff) + lam)
    L = sla.cholesky(C, lower=True)
    a = sla.solve_triangular(L, y, lower=True)
    mll = -0.5 * len(y) * np.log(2 * np.pi) - np.sum(np.log(np.diag(L))) - 0.5 * a @ a
    Kus = orc.cov_full(spec, Xu, Xs)
    Qsf = Kus.T @ np.linalg.solve(Kuu, Kuf)
    mu = Qsf @ sla.cho_solve((L, True), y)
    B = sla.solve_triangular(L, Qsf.T, lower=True)
    var = orc.cov_diag(spec, Xs) - np.sum(B * B, 0)
    if pred_noise:
        var = var + spec["sigma"] ** 2
    return mll, mu, var


@pytest.mark.parametrize("kind,P", [("ExpQuad", 1), ("Matern52", 1), ("Matern32", 2)])
def test_oracle_woodbury_form_equals_dense_fitc(kind, P):
    spec, X, y, Xs, Xu = problem(n=300, d=2, kind=kind, m=25, P=P)
    mll0, mu0, var0 = dense_fitc(spec, X, y, Xu, Xs)
    assert orc.fitc_mll(spec, X, y, Xu) == pytest.approx(mll0, rel=1e-9)
    mu, var = orc.fitc_predict(spec, X, y, Xu, Xs, True)
    np.testing.assert_allclose(mu, mu0, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, var0, rtol=1e-7, atol=1e-9)


def test_fitc_with_all_training_points_as_inducing_points_is_the_exact_gp():
    spec, X, y, Xs, _ = problem(n=200, d=2, kind="ExpQuad")
    spec["sigma"] = 0.3
    mu, var = orc.fitc_predict(spec, X, y, X, Xs, True)
    mu0, var0 = orc.predict(spec, X, y, Xs, True)
    np.testing.assert_allclose(mu, mu0, atol=2e-3)       # equal up to the jitter on Kuu (amplified by the conditioning of Kuu)
    np.testing.assert_allclose(var, var0, atol=2e-3)


def test_kmeans_inducing_points_follow_pymc_whitening():
    from gumbi_b200.sparse import kmeans_inducing_points

    rng = np.random.default_rng(0)
    X = np.hstack([rng.standard_normal((400, 2)) * [1.0, 50.0], np.full((400, 1), 3.0)])   # a constant column is left unscaled
    Xu = kmeans_inducing_points(20, X, seed=5)
    assert Xu.shape[1] == 3 and 1 <= len(Xu) <= 20
    np.testing.assert_allclose(Xu[:, 2], 3.0)
    np.testing.assert_array_equal(Xu, orc.kmeans_inducing_points(20, X, seed=5))
    np.testing.assert_array_equal(Xu, kmeans_inducing_points(20, X, seed=5))
    # whitening matters: without it the second column would dominate the clustering
    assert np.std(Xu[:, 0]) > 0.3


def host_gp(sparse=True, P=1, n_u=30, **kw):
    spec, X, y, Xs = orc.synthetic_problem(240, 2, P=P, M_res=6, kind="ExpQuad")
    cat = {}
    if P > 1:
        cat = dict(categorical_dims=["Variable"], categorical_levels={"Variable": [f"y{p}" for p in range(P)]}, outputs=[f"y{p}" for p in range(P)])
    gp = HostGP(X, y, ["x0", "x1"], **cat)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        gp.build_model(sparse=sparse, n_u=n_u, **kw)
    return gp, spec, X, y, Xs, w


def test_backend_sparse_layout_and_predict_dispatch():
    gp, spec, X, y, Xs, w = host_gp(P=2)
    assert any("Reverting to scalar-valued noise" in str(x.message) for x in w)       # GP.py:573-577
    assert gp.model_specs["sparse"] is True and gp.model_specs["n_u"] == 30
    assert "W_Output_noise" not in gp.param_shapes() and "W_Variable" in gp.param_shapes()
    assert gp._Xu.shape[1] == X.shape[1] and len(gp._Xu) <= 30
    point = {"ls_total": spec["terms"][0]["ls"], "η_total": 1.0, "σ": 0.2, "W_Variable": spec["terms"][0]["coreg"][0]["W"],
             "κ_Variable": spec["terms"][0]["coreg"][0]["kappa"]}
    gp.find_MAP(point=point)
    mu, var = gp.predict(Xs, with_noise=True)
    s2 = gp.spec_from_point(gp.MAP)
    assert s2["noise_coreg"] is None
    mu0, var0 = orc.fitc_predict(s2, X, y, gp._Xu, Xs, True)
    np.testing.assert_allclose(mu, mu0)
    np.testing.assert_allclose(var, var0)
    assert gp.marginal_log_likelihood() == pytest.approx(orc.fitc_mll(s2, X, y, gp._Xu))
    with pytest.raises(NotImplementedError):
        gp.conditional(Xs[:4])


def test_backend_sparse_find_map_improves_the_fitc_objective():
    from gumbi_b200.map import make_objective

    gp, spec, X, y, Xs, _ = host_gp(P=1, n_u=20)
    fun, x0, unpack, names, positive = make_objective(gp)
    f0, g0 = fun(x0)
    # central-difference gradient of the objective agrees with a coarser independent difference of the value
    k = 1
    e = np.zeros_like(x0); e[k] = 1e-4
    fd = (fun(x0 + e)[0] - fun(x0 - e)[0]) / 2e-4
    assert g0[k] == pytest.approx(fd, rel=1e-4, abs=1e-6)
    MAP = gp.find_MAP(options={"maxiter": 15})
    fun2, *_ = make_objective(gp)
    x1 = np.concatenate([np.atleast_1d(MAP[n + "_log__"] if p else MAP[n]).reshape(-1) for n, p in zip(names, positive)])
    assert fun2(x1)[0] < f0 - 1.0
    mu, var = gp.predict(Xs)
    assert np.all(np.isfinite(mu)) and np.all(var > 0)


# ---------------------------------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,d,kind,m,P", [(500, 3, "Matern52", 40, 1), (1500, 2, "ExpQuad", 100, 1), (900, 4, "Matern32", 127, 1), (700, 2, "ExpQuad", 128, 1),
                                          (1100, 3, "ExpQuad", 300, 1), (600, 2, "Matern52", 50, 2)])
def test_fitc_device_against_oracle(lib_built, n, d, kind, m, P):
    from gumbi_b200 import GPEngine

    spec, X, y, Xs, Xu = problem(n=n, d=d, kind=kind, m=m, P=P)
    eng = GPEngine(0)
    eng.set_train(X, y)
    eng.set_kernel(spec)
    eng.fitc_factorize(Xu)
    assert eng.fitc_mll() == pytest.approx(orc.fitc_mll(spec, X, y, Xu), rel=1e-9)
    for noise in (True, False):
        mu, var = eng.fitc_predict(Xs, noise)
        mu0, var0 = orc.fitc_predict(spec, X, y, Xu, Xs, noise)
        np.testing.assert_allclose(mu, mu0, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(var, var0, rtol=1e-6, atol=1e-8)
    # a second kernel on the same handle re-uses the inner systems
    spec["terms"][0]["ls"] = [1.7 * v for v in spec["terms"][0]["ls"]]
    spec["sigma"] = 0.25
    eng.set_kernel(spec)
    with pytest.raises(ValueError):
        eng.fitc_mll()          # stale: set_kernel invalidates the FITC factor
    eng.fitc_factorize(Xu)
    assert eng.fitc_mll() == pytest.approx(orc.fitc_mll(spec, X, y, Xu), rel=1e-9)
    # the exact path of the same handle is untouched
    eng.factorize()
    assert eng.mll() == pytest.approx(orc.mll(spec, X, y), rel=1e-9)
    eng.close()


@pytest.mark.gpu
def test_fitc_device_refuses_noise_coregion_and_reports_non_pd(lib_built):
    from gumbi_b200 import GPEngine

    spec, X, y, Xs = orc.synthetic_problem(300, 2, P=2, M_res=5, kind="ExpQuad")
    eng = GPEngine(0)
    eng.set_train(X, y)
    eng.set_kernel(spec)                       # carries the Output_noise Coregion
    with pytest.raises(ValueError, match="scalar noise"):
        eng.fitc_factorize(X[:20])
    spec["noise_coreg"] = None
    spec["terms"][0]["lin_idx"] = [0]; spec["terms"][0]["c"] = [0.0]; spec["terms"][0]["tau"] = -50.0    # robustly indefinite Kuu
    eng.set_kernel(spec)
    with pytest.raises(np.linalg.LinAlgError):
        eng.fitc_factorize(X[:40])
    eng.close()


@pytest.mark.gpu
def test_fitc_backend_fit_and_predict_on_device(lib_built):
    """``build_model(sparse=True)`` -> ``find_MAP`` -> ``predict`` through ArrayGP on the CUDA engine; the MAP point is then replayed
    through the oracle."""
    from gumbi_b200 import ArrayGP

    spec, X, y, Xs = orc.synthetic_problem(800, 2, P=1, M_res=8, kind="ExpQuad")
    gp = ArrayGP(X, y, ["x0", "x1"])
    gp.build_model(sparse=True, n_u=60)
    MAP = gp.find_MAP(options={"maxiter": 25})
    mu, var = gp.predict(Xs, with_noise=True)
    s2 = gp.spec_from_point(MAP)
    mu0, var0 = orc.fitc_predict(s2, X, y, gp._Xu, Xs, True)
    np.testing.assert_allclose(mu, mu0, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(var, var0, rtol=1e-6, atol=1e-8)
    assert gp.marginal_log_likelihood() == pytest.approx(orc.fitc_mll(s2, X, y, gp._Xu), rel=1e-9)
    gp.engine.close()
