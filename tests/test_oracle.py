"""CPU tests of the oracle itself: it must agree with an independent implementation (scikit-learn), with explicit
Kronecker algebra for the Coregion/ICM expansion, with analytic identities, and with the committed golden vectors."""
import numpy as np
import pytest
import scipy.linalg as sla

from conftest import GOLDEN_CASES, load_golden
from oracle import gp_oracle as orc


@pytest.mark.parametrize("kind,nu", [("ExpQuad", None), ("Matern52", 2.5), ("Matern32", 1.5), ("Matern12", 0.5)])
def test_oracle_matches_sklearn(kind, nu):
    from sklearn.gaussian_process import GaussianProcessRegressor as GPR
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel as C, Matern

    rng = np.random.default_rng(0)
    n, d, M = 300, 3, 50
    X = rng.standard_normal((n, d))
    y = np.sin(X).sum(1) + 0.1 * rng.standard_normal(n)
    Xs = rng.standard_normal((M, d))
    ls = np.array([0.9, 1.7, 2.4])
    eta, sigma = 1.3, 0.2
    spec = {"terms": [{"kind": kind, "cont_idx": [0, 1, 2], "ls": ls.tolist(), "eta": eta}], "sigma": sigma, "jitter": 1e-6}
    mu, var = orc.predict(spec, X, y, Xs, pred_noise=False)
    base = RBF(ls) if nu is None else Matern(ls, nu=nu)
    gpr = GPR(kernel=C(eta ** 2, "fixed") * base, alpha=sigma ** 2 + 1e-6, optimizer=None).fit(X, y)
    m2, s2 = gpr.predict(Xs, return_std=True)
    # the Matern kernels differ from sklearn by PyMC's sqrt(r2 + 1e-12) (Stationary.euclidean_dist)
    tol = 1e-9 if kind == "ExpQuad" else 2e-5
    np.testing.assert_allclose(mu, m2, rtol=tol, atol=tol)
    np.testing.assert_allclose(var, s2 ** 2, rtol=10 * tol, atol=tol)
    # log marginal likelihood too
    np.testing.assert_allclose(orc.mll(spec, X, y), gpr.log_marginal_likelihood_value_, rtol=1e-9 if kind == "ExpQuad" else 2e-5)


def test_coregion_is_kronecker_on_aligned_outputs():
    """ICM: stacking one copy of X per output (base.py:459-464) makes K = B (x) Kx."""
    spec, X, y, Xs = orc.synthetic_problem(40, 3, P=3, M_res=4)
    n = 40
    K = orc.cov_full(spec, X)
    cg = spec["terms"][0]["coreg"][0]
    B = orc.coregion_B(cg["W"], cg["kappa"])
    term = dict(spec["terms"][0], coreg=[])
    Kx = orc._term_full(term, X[:n], None)
    np.testing.assert_allclose(K, np.kron(B, Kx), rtol=1e-13, atol=1e-15)
    # LCM (Q=2): sum of Kronecker products
    spec2, X2, _, _ = orc.synthetic_problem(30, 2, P=2, M_res=3, Q=2)
    K2 = orc.cov_full(spec2, X2)
    ref = 0
    for t in spec2["terms"]:
        Bq = orc.coregion_B(t["coreg"][0]["W"], t["coreg"][0]["kappa"])
        ref = ref + np.kron(Bq, orc._term_full(dict(t, coreg=[]), X2[:30], None))
    np.testing.assert_allclose(K2, ref, rtol=1e-13, atol=1e-15)


def test_noise_coregion_diag_and_jitter():
    spec, X, y, _ = orc.synthetic_problem(25, 2, P=2, M_res=3)
    spec["noise_coreg"]["W"] = [[0.3, -0.1], [0.2, 0.5]]
    spec["noise_coreg"]["kappa"] = [0.7, 1.9]
    K = orc.train_cov(spec, X)
    K0 = orc.cov_full(spec, X)
    Bn = orc.coregion_B(spec["noise_coreg"]["W"], spec["noise_coreg"]["kappa"])
    expect = spec["sigma"] ** 2 * np.diag(Bn)[X[:, -1].astype(int)] + 1e-6
    np.testing.assert_allclose(np.diag(K) - np.diag(K0), expect, rtol=1e-12)
    assert np.allclose(K - np.diag(np.diag(K)), K0 - np.diag(np.diag(K0)))


def test_linear_kernel_and_diag():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((20, 3))
    spec = {"terms": [{"kind": "ExpQuad", "cont_idx": [0, 1, 2], "ls": [1.0, 2.0, 3.0], "eta": 0.9, "lin_idx": [0, 2],
                       "c": [0.2, -0.4], "tau": 0.3}], "sigma": 0.1}
    K = orc.cov_full(spec, X)
    Xl = X[:, [0, 2]] - np.array([0.2, -0.4])
    Kc = orc._term_full(dict(spec["terms"][0], lin_idx=[]), X, None)
    np.testing.assert_allclose(K, Kc + 0.3 * Xl @ Xl.T, rtol=1e-13)
    np.testing.assert_allclose(orc.cov_diag(spec, X), np.diag(K), rtol=1e-12)


def test_posterior_identities():
    """sigma -> 0 reproduces y at training points; variance >= 0 and <= prior; noise adds exactly sigma^2."""
    spec, X, y, Xs = orc.synthetic_problem(60, 2, M_res=5)
    spec["sigma"] = 1e-4
    spec["terms"][0]["ls"] = [0.25, 0.25]  # short lengthscale: K well conditioned, the GP interpolates
    mu, var = orc.predict(spec, X, y, X, pred_noise=False)
    np.testing.assert_allclose(mu, y, atol=5e-3)
    assert np.all(np.abs(var) < 1e-4)
    spec["terms"][0]["ls"] = [1.5, 2.0]
    spec["sigma"] = 0.2
    mu, var = orc.predict(spec, X, y, Xs, pred_noise=False)
    mu2, var2 = orc.predict(spec, X, y, Xs, pred_noise=True)
    assert np.all(var > 0) and np.all(var <= orc.cov_diag(spec, Xs) + 1e-12)
    np.testing.assert_allclose(var2 - var, 0.04, rtol=1e-9)
    np.testing.assert_allclose(mu, mu2)
    # direct formula through the explicit inverse
    K = orc.train_cov(spec, X)
    Ks = orc.cov_full(spec, X, Xs)
    np.testing.assert_allclose(mu, Ks.T @ np.linalg.solve(K, y), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, orc.cov_diag(spec, Xs) - np.einsum("ij,ij->j", Ks, np.linalg.solve(K, Ks)), rtol=1e-7, atol=1e-10)


def test_clip_and_matern_epsilon():
    X = np.array([[0.0], [0.0], [1e-9]])
    r2 = orc.square_dist(X, None, np.array([1.0]))
    assert np.all(r2 >= 0)
    k = orc.stationary_full("Matern52", X, None, np.array([1.0]))
    assert k[0, 1] == pytest.approx((1 + np.sqrt(5) * 1e-6 + 5 / 3 * 1e-12) * np.exp(-np.sqrt(5) * 1e-6), rel=1e-15)


def test_not_positive_definite_raises():
    spec, X, y, _ = orc.synthetic_problem(30, 2)
    spec["sigma"] = 0.0
    spec["jitter"] = 0.0
    X[5] = X[4]
    with pytest.raises(np.linalg.LinAlgError):
        orc.factorize(spec, X, y)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_reproduces_golden(case):
    g = load_golden(case)
    spec = g["meta"]["spec"]
    mu, var = orc.predict(spec, g["X"], g["y"], g["points"], pred_noise=True)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g["var"], rtol=1e-8, atol=1e-11)
    mu, var = orc.predict(spec, g["X"], g["y"], g["points"], pred_noise=False)
    np.testing.assert_allclose(var, g["var_noisefree"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(orc.mll(spec, g["X"], g["y"]), float(g["mll"]), rtol=1e-10)


def test_golden_cases_present():
    assert {"simple_regression_ExpQuad", "multioutput_regression", "categorical_additive"} <= set(GOLDEN_CASES)


def test_oracle_mll_gradient_against_central_differences():
    import copy

    spec, X, y, _ = orc.synthetic_problem(30, 3, P=2, M_res=3, kind="Matern32", Q=2)
    for t in spec["terms"]:
        t["lin_idx"], t["c"], t["tau"] = [0, 1], [0.3, -0.2], 0.05
    spec["noise_coreg"]["W"] = (0.3 * np.random.default_rng(1).standard_normal((2, 2))).tolist()
    val, g = orc.mll_grad(spec, X, y)
    assert val == pytest.approx(orc.mll(spec, X, y), rel=1e-12)

    def fd(mut, eps=1e-6):
        a, b = copy.deepcopy(spec), copy.deepcopy(spec)
        mut(a, eps), mut(b, -eps)
        return (orc.mll(a, X, y) - orc.mll(b, X, y)) / (2 * eps)

    def bump(path):
        def mut(s, e):
            obj = s
            for k in path[:-1]:
                obj = obj[k]
            obj[path[-1]] += e
        return mut

    checks = [(g["terms"][1]["ls"][2], ["terms", 1, "ls", 2]), (g["terms"][0]["eta"], ["terms", 0, "eta"]),
              (g["terms"][0]["tau"], ["terms", 0, "tau"]), (g["terms"][1]["c"][0], ["terms", 1, "c", 0]),
              (g["terms"][1]["coreg"][0]["W"][1, 0], ["terms", 1, "coreg", 0, "W", 1, 0]),
              (g["terms"][0]["coreg"][0]["kappa"][1], ["terms", 0, "coreg", 0, "kappa", 1]),
              (g["sigma"], ["sigma"]), (g["noise_coreg"]["W"][0, 1], ["noise_coreg", "W", 0, 1]),
              (g["noise_coreg"]["kappa"][0], ["noise_coreg", "kappa", 0])]
    for got, path in checks:
        assert got == pytest.approx(fd(bump(path)), rel=2e-6, abs=1e-7), path


def test_blocked_oracle_of_the_full_size_gpu_tests_matches_the_plain_one():
    """tests/test_gpu_fullsize.py assembles K by row blocks to bound host memory at N = 32768; same numbers as orc.factorize/conditional."""
    from test_gpu_fullsize import oracle_subsample

    spec, X, y, Xs = orc.synthetic_problem(350, 4, P=2, M_res=9, kind="Matern52")
    ref = oracle_subsample(spec, X, y, Xs[:40], block=128)
    L, v = orc.factorize(spec, X, y)
    mu, var = orc.conditional(spec, X, L, v, Xs[:40], True)
    np.testing.assert_allclose(ref["mu"], mu, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(ref["var"], var, rtol=1e-10, atol=1e-12)
    assert ref["mll"] == pytest.approx(orc.mll(spec, X, y), rel=1e-12)
