"""CPU tests of the host-side backend logic (gumbi_b200/backend.py): kernel-structure lowering, MAP keys, argument and
error behaviour mirrored from PymcGP -- with a test-only engine double so that no GPU is needed."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from gumbi_b200.backend import ArrayRegressor, B200Backend
from oracle import gp_oracle as orc


class OracleEngine:
    """Test double standing where GPEngine stands (tests/ only)."""

    def set_train(self, X, y):
        self.X, self.y = X, y

    def set_kernel(self, spec):
        self.spec = spec
        self.n_set = getattr(self, "n_set", 0) + 1

    def factorize(self):
        self.L, self.v = orc.factorize(self.spec, self.X, self.y)
        self.n_fact = getattr(self, "n_fact", 0) + 1

    def predict(self, Xs, pred_noise=True):
        return orc.conditional(self.spec, self.X, self.L, self.v, Xs, pred_noise)

    def predict_full(self, Xs, pred_noise=False):
        return orc.conditional_full(self.spec, self.X, self.L, self.v, Xs, pred_noise)

    def mll(self):
        return orc.mll(self.spec, self.X, self.y)

    def mll_grad(self, spec):
        return orc.mll_grad(spec, self.X, self.y)

    def get_alpha(self):
        from scipy.linalg import solve_triangular

        return solve_triangular(self.L, self.v, lower=True, trans="T")

    def fitc_factorize(self, Xu):
        self.Xu = np.asarray(Xu, dtype=np.float64)
        orc.fitc_factorize(self.spec, self.X, self.y, self.Xu)   # raises LinAlgError like the device path
        self.n_fitc = getattr(self, "n_fitc", 0) + 1

    def fitc_mll(self):
        return orc.fitc_mll(self.spec, self.X, self.y, self.Xu)

    def fitc_predict(self, Xs, pred_noise=True):
        return orc.fitc_predict(self.spec, self.X, self.y, self.Xu, Xs, pred_noise)

    def set_option(self, name, value):
        pass

    def close(self):
        pass


class HostGP(B200Backend, ArrayRegressor):
    def __init__(self, *a, **k):
        ArrayRegressor.__init__(self, *a, **k)
        self._init_backend()
        self.engine = OracleEngine()


def gp_from_golden(g, cls=HostGP, **kw):
    m = g["meta"]
    levels = {k: v for k, v in m["categorical_levels"].items()}
    gp = cls(g["X"], g["y"], m["continuous_dims"], linear_dims=m["linear_dims"], categorical_dims=m["categorical_dims"],
             categorical_levels=levels, out_col=m["out_col"], outputs=m["outputs"], additive=m["additive"], **kw)
    gp.build_model(continuous_kernel=m["continuous_kernel"], ARD=m["ARD"])
    return gp


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_spec_lowering_matches_reference_structure(case):
    """The spec our backend derives from (dims, MAP) equals the one generated under the real gumbi Regressor."""
    g = load_golden(case)
    gp = gp_from_golden(g)
    gp.find_MAP(point=g["meta"]["point"])
    spec = gp.spec_from_point(gp.MAP)
    want = g["meta"]["spec"]
    assert len(spec["terms"]) == len(want["terms"])
    for a, b in zip(spec["terms"], want["terms"]):
        assert a["kind"] == b["kind"] and a["cont_idx"] == b["cont_idx"] and a["lin_idx"] == b["lin_idx"]
        np.testing.assert_allclose(a["ls"], b["ls"])
        assert [c["col"] for c in a["coreg"]] == [c["col"] for c in b["coreg"]]
        for ca, cb in zip(a["coreg"], b["coreg"]):
            np.testing.assert_allclose(ca["W"], cb["W"])
            np.testing.assert_allclose(ca["kappa"], cb["kappa"])
    assert (spec["noise_coreg"] is None) == (want["noise_coreg"] is None)
    mu, var = gp.predict(g["points"], with_noise=True)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g["var"], rtol=1e-8, atol=1e-11)


def test_map_keys_and_base_class_contract():
    """base.py:592-593 reads MAP['W_<out_col>'] and MAP['kappa_<out_col>']; pm.find_MAP also returns *_log__ entries."""
    g = load_golden("multioutput_regression")
    gp = gp_from_golden(g)
    MAP = gp.find_MAP(point=g["meta"]["point"])
    out_col = g["meta"]["out_col"]
    assert MAP[f"W_{out_col}"].shape == (5, 2) and MAP[f"κ_{out_col}"].shape == (5,)
    assert MAP["W_Output_noise"].shape == (5, 2)
    for k in ("ls_total", "η_total", "τ_total", "σ", f"κ_{out_col}"):
        np.testing.assert_allclose(np.exp(MAP[k + "_log__"]), MAP[k])
    assert set(gp.model_specs) == {"seed", "continuous_kernel", "heteroskedastic_inputs", "heteroskedastic_outputs", "sparse", "n_u"}


def test_factor_is_cached_between_predicts_and_reset_by_find_MAP():
    g = load_golden("simple_regression_ExpQuad")
    gp = gp_from_golden(g)
    gp.find_MAP(point=g["meta"]["point"])
    gp.predict(g["points"])
    gp.predict(g["points"], with_noise=False)
    assert gp.engine.n_fact == 1
    gp.predict_cold(g["points"])
    assert gp.engine.n_fact == 2
    gp.find_MAP(point=g["meta"]["point"])
    gp.predict(g["points"])
    assert gp.engine.n_fact == 3


def test_error_behaviour_mirrors_pymcgp():
    g = load_golden("simple_regression_ExpQuad")
    m = g["meta"]
    gp = HostGP(g["X"], g["y"], m["continuous_dims"], linear_dims=m["linear_dims"])
    with pytest.raises(NotImplementedError, match="Heteroskedasticity over inputs"):
        gp.build_model(heteroskedastic_inputs=True)  # GP.py:518-519
    with pytest.raises(ValueError, match="Continuous kernel must be one of"):
        gp.build_model(continuous_kernel="RatQuad")  # assert_in, GP.py:674
    gp.precision = "tf32"
    with pytest.raises(NotImplementedError, match="FITC"):
        gp.build_model(sparse=True)          # the sparse path is fp64, single GPU
    gp.precision = "fp64"
    with pytest.raises(AssertionError):
        HostGP(g["X"], g["y"], m["continuous_dims"]).find_MAP()  # GP.py:808 `assert self.model is not None`
    gp.build_model()
    with pytest.raises(RuntimeError, match="before find_MAP"):
        gp.predict(g["points"])
    gp.find_MAP(point=m["point"])
    with pytest.raises(NotImplementedError, match="additive sublevels"):
        gp.predict(g["points"], additive_level="global")  # GP.py:840-841
    with pytest.raises(KeyError, match="missing from point"):
        gp.find_MAP(point={"ls_total": [1.0, 1.0, 1.0]})
    bad = dict(m["point"], σ=-1.0)
    with pytest.raises(ValueError, match="must be positive"):
        gp.find_MAP(point=bad)


def test_nan_rows_dropped_like_get_shaped_data():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((30, 2))
    y = rng.standard_normal(30)
    y[[3, 7]] = np.nan
    gp = HostGP(X, y, ["u", "v"])
    gp.build_model()
    assert gp._X.shape == (28, 2) and not np.isnan(gp._y).any()  # base.py:469-471


def test_ard_false_uses_single_lengthscale():
    g = load_golden("test_dataset_filtered")
    gp = gp_from_golden(g)
    assert gp.param_shapes()["ls_total"] == (1,)
    gp.find_MAP(point=g["meta"]["point"])
    mu, var = gp.predict(g["points"])
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)


# ---------------------------------------------------------------------------------------------------------------------
# find_MAP host logic (gumbi_b200/map.py) with the oracle standing in for the device objective
# ---------------------------------------------------------------------------------------------------------------------
def test_ls_prior_matches_pdist_definition_and_mass():
    from scipy import stats
    from scipy.spatial.distance import pdist

    from gumbi_b200.map import find_constrained_invgamma, parse_ls_limits

    rng = np.random.default_rng(3)
    X = np.round(rng.standard_normal((200, 3)), 1)  # many duplicated coordinates
    lowers, uppers = parse_ls_limits(X, ARD=True)
    for j in range(3):
        d = pdist(X[:, [j]])  # gp_utils.py:34-43
        d = d[d != 0]
        assert lowers[j] == pytest.approx(max(d.min(), 0.01)) and uppers[j] == pytest.approx(d.max())
    lo1, hi1 = parse_ls_limits(X, ARD=False)
    d = pdist(X)
    d = d[d != 0]
    assert lo1 == [pytest.approx(max(d.min(), 0.01))] and hi1 == [pytest.approx(d.max())]
    # default = what the reference gets from PyMC: SciPy's SLSQP meets the mass constraint and stops (objective <= 1e-4), so the
    # lower tail is NOT 1 %; exact=True solves both conditions
    p = find_constrained_invgamma(0.1, 5.0, mass=0.98)
    cdf = lambda x: stats.invgamma.cdf(x, p["alpha"], scale=p["beta"])
    assert cdf(5.0) - cdf(0.1) == pytest.approx(0.98, abs=1e-4)
    assert cdf(0.1) < 1e-6 and p["alpha"] == pytest.approx(3.8958, rel=1e-3) and p["beta"] == pytest.approx(4.8302, rel=1e-3)
    p = find_constrained_invgamma(0.1, 5.0, mass=0.98, exact=True)
    cdf = lambda x: stats.invgamma.cdf(x, p["alpha"], scale=p["beta"])
    assert cdf(5.0) - cdf(0.1) == pytest.approx(0.98, abs=1e-6)
    assert cdf(0.1) == pytest.approx(0.01, abs=1e-6)


def test_find_map_objective_gradient_and_improvement():
    """The packed objective's gradient (device gradient + priors + log transforms) agrees with central differences, and
    L-BFGS-B increases the log-posterior from the initial point; MAP carries PyMC's keys."""
    from gumbi_b200 import map as gmap

    g = load_golden("multioutput_regression")
    gp = gp_from_golden(g)
    shapes = gp.param_shapes()
    MAP, res = gmap.find_map(gp, return_raw=True, options={"maxiter": 25})
    for k, shp in shapes.items():
        assert np.shape(MAP[k]) == tuple(shp)
        if k.split("_")[0] in gmap.POSITIVE:
            np.testing.assert_allclose(np.exp(MAP[k + "_log__"]), MAP[k])
    assert gp.map_evals >= 2 and np.isfinite(res.fun)
    # objective at the start vs at the optimum
    pri = gmap.build_priors(gp)

    def logpost(point):
        spec = gp.spec_from_point(gp._complete_point(point))
        return orc.mll(spec, gp._X, gp._y) + sum(float(pri[n][0](np.asarray(point[n], dtype=float))) for n in shapes)

    start = {n: pri[n][2](shapes[n] if shapes[n] != () else (1,)).reshape(shapes[n]) for n in shapes}
    assert logpost({k: MAP[k] for k in shapes}) > logpost(start) + 1.0
    # gradient check of the packed objective by central differences at the (non-stationary) start point
    fun, x0, unpack, names, _ = gmap.make_objective(gp)
    f0, g0 = fun(x0)
    assert f0 == pytest.approx(-logpost(unpack(x0)), rel=1e-10)
    rng = np.random.default_rng(0)
    for _ in range(4):
        dvec = rng.standard_normal(x0.size)
        dvec /= np.linalg.norm(dvec)
        fd = (fun(x0 + 1e-6 * dvec)[0] - fun(x0 - 1e-6 * dvec)[0]) / 2e-6
        assert g0 @ dvec == pytest.approx(fd, rel=2e-5, abs=1e-6)


def test_fit_runs_end_to_end_on_the_engine_double():
    g = load_golden("simple_regression_Matern52")
    gp = gp_from_golden(g)
    gp.find_MAP(options={"maxiter": 10})
    assert isinstance(gp.MAP, dict)  # tests/test_regression.py:172-182 asserts exactly this of PymcGP
    mu, var = gp.predict(g["points"])
    assert mu.shape == var.shape == (len(g["points"]),) and np.all(var > 0)


def test_conditional_and_joint_samples():
    """conditional(): diag of the full covariance equals predict()'s variance; joint draws have the conditional's moments."""
    g = load_golden("simple_regression_ExpQuad")
    gp = gp_from_golden(g)
    gp.find_MAP(point=g["meta"]["point"])
    pts = g["points"][:40]
    mu, cov = gp.conditional(pts)
    mu1, var1 = gp.predict(pts, with_noise=False)
    np.testing.assert_allclose(mu, mu1, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(np.diag(cov), var1, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(cov, cov.T, rtol=0, atol=1e-12)
    _, cov_n = gp.conditional(pts, pred_noise=True)
    np.testing.assert_allclose(np.diag(cov_n) - np.diag(cov), g["meta"]["point"]["σ"] ** 2, rtol=1e-9)
    draws = gp.sample_conditional(pts, size=4000, random_seed=1)
    assert draws.shape == (4000, 40)
    np.testing.assert_allclose(draws.mean(0), mu, atol=5 * np.sqrt(np.diag(cov).max() / 4000) + 1e-3)
    np.testing.assert_allclose(np.cov(draws.T), cov + 1e-6 * np.eye(40), atol=0.1 * np.abs(cov).max())


def test_predict_cold_fused_uses_the_one_pass_entry_point():
    """predict_cold(fused=True) -> engine.factorize_predict (gb2_factorize_predict); afterwards the factor counts as current."""
    g = load_golden("simple_regression_ExpQuad")
    gp = gp_from_golden(g)
    gp.find_MAP(point=g["meta"]["point"])
    with pytest.raises(NotImplementedError):                   # the plain double has no one-pass entry point
        gp.predict_cold(g["points"], fused=True)

    class Fused(OracleEngine):
        def factorize_predict(self, Xs, pred_noise=True):
            self.factorize()
            return self.predict(Xs, pred_noise)

    gp.engine = Fused()
    gp.engine.set_train(gp._X, gp._y)
    mu, var = gp.predict_cold(g["points"], with_noise=True, fused=True)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    n = gp.engine.n_fact
    gp.predict(g["points"])                                     # no second factorisation: the fused pass left the factor behind
    assert gp.engine.n_fact == n


def test_ls_bounds_follow_prepare_lengthscales():
    """fit(ls_bounds=...) (GP.py:630-646): bounds of the named dimension replace the pdist defaults (lower never below the smallest
    distance / 0.01), NaN = unbounded, and the reference's single-dimension check is kept as written."""
    from gumbi_b200.map import build_priors, find_constrained_invgamma, get_ls_prior, ls_bounds_z

    rng = np.random.default_rng(0)
    X = rng.standard_normal((40, 1))
    y = np.sin(X[:, 0])
    gp = HostGP(X, y, ["x0"])
    gp.build_model(ls_bounds={"x0": (0.5, 2.0)})
    assert ls_bounds_z(gp) == ((0.5,), (2.0,))
    want = find_constrained_invgamma(0.5, 2.0, mass=0.98)
    pri = build_priors(gp)["ls_total"]
    x = np.array([0.9])
    a, b = want["alpha"], want["beta"]
    assert pri[1](x)[0] == pytest.approx(-(a + 1) / x[0] + b / x[0] ** 2, rel=1e-12)
    gp.build_model(ls_bounds={"x0": (np.nan, 2.0)})             # NaN lower bound -> the pdist default
    free = get_ls_prior(X, ARD=True, lower=None, upper=(2.0,), mass=0.98)
    assert build_priors(gp)["ls_total"][1](x)[0] == pytest.approx(-(free["alpha"][0] + 1) / x[0] + free["beta"][0] / x[0] ** 2, rel=1e-12)
    gp.build_model(ls_bounds=None)
    assert ls_bounds_z(gp) == (None, None)
    X2 = rng.standard_normal((40, 2))
    gp2 = HostGP(X2, y, ["x0", "x1"])
    gp2.build_model(ls_bounds={"x0": (0.5, 2.0), "x1": (0.5, 2.0)})
    with pytest.raises(ValueError, match="single dimension"):   # `... or len(upper) != 1` as written in the reference
        build_priors(gp2)
    gp2.build_model(ls_bounds={"x1": (0.5, 2.0)})               # one bounded dimension is broadcast to all (parse_ls_limits)
    p2 = build_priors(gp2)["ls_total"]
    assert np.all(np.isfinite(p2[1](np.array([0.9, 1.1]))))


def reference_form(gp, spec, mode, zp):
    """The spec the reference would build for a periodic model: original columns, pm.gp.cov.Periodic / WarpedInput semantics."""
    import copy

    ref = copy.deepcopy(spec)
    t = ref["terms"][0]
    t["cont_idx"] = list(gp._layout["idx_s"])
    t["ls"] = np.atleast_1d(np.asarray(gp.MAP["ls_total"], dtype=np.float64)).tolist()
    if mode == "Periodic":
        t["kind"], t["period"] = "Periodic", list(zp)
    else:
        t["warp_period"] = list(zp)
    return ref


@pytest.mark.parametrize("kernel,d,ARD", [("Periodic", 1, True), ("Periodic", 3, True), ("Periodic", 2, False), ("Matern52+Periodic", 1, True),
                                          ("ExpQuad+Periodic", 1, True)])
def test_periodic_kernels_are_lowered_onto_the_stationary_path(kernel, d, ARD):
    """continuous_kernel="Periodic" / "<K>+Periodic" (GP.py:389-447, :664-689): the engine sees a stationary kernel on the warped
    coordinates [sin(2 pi x/T), cos(2 pi x/T)]; posterior, likelihood and ls-gradient equal the reference formulation
    (pm.gp.cov.Periodic / WarpedInput, restated in oracle/gp_oracle.py) on the original columns."""
    from gumbi_b200.map import make_objective

    rng = np.random.default_rng(4)
    X = rng.standard_normal((60, d))
    y = np.sin(3 * X[:, 0]) + 0.1 * rng.standard_normal(60)
    dims = [f"x{j}" for j in range(d)]
    zp = [1.7 + 0.4 * j for j in range(d)]
    gp = HostGP(X, y, dims, linear_dims=dims[:1])
    gp.build_model(continuous_kernel=kernel, period=dict(zip(dims, zp)), ARD=ARD)
    assert gp.engine.X.shape[1] == 3 * d                        # original columns + sin + cos
    pt = {"ls_total": rng.uniform(0.5, 1.5, size=d if ARD else 1), "η_total": 1.3, "σ": 0.2, "c_total": [0.1], "τ_total": 0.3}
    gp.find_MAP(point=pt)
    spec = gp.spec_from_point(gp.MAP)
    assert spec["terms"][0]["kind"] == ("ExpQuad" if kernel == "Periodic" else kernel.split("+")[0])
    ref = reference_form(gp, spec, "Periodic" if kernel == "Periodic" else "warped", zp)
    Xs = rng.standard_normal((25, d))
    mu, var = gp.predict(Xs, with_noise=True)
    mu0, var0 = orc.predict(ref, X, y, Xs, True)
    np.testing.assert_allclose(mu, mu0, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, var0, rtol=1e-8, atol=1e-11)
    assert gp.marginal_log_likelihood() == pytest.approx(orc.mll(ref, X, y), rel=1e-11)
    # the objective's gradient w.r.t. log(ls) against central differences of the reference-form likelihood + prior
    fun, x0, unpack, names, positive = make_objective(gp)
    x = x0.copy()
    f0, g0 = fun(x)
    i0 = 0
    for n in names:
        size = int(np.prod(gp.param_shapes()[n])) if gp.param_shapes()[n] != () else 1
        if n == "ls_total":
            for k in range(size):
                e = np.zeros_like(x); e[i0 + k] = 1e-6
                num = (fun(x + e)[0] - fun(x - e)[0]) / 2e-6
                assert g0[i0 + k] == pytest.approx(num, rel=2e-5, abs=1e-7)
        i0 += size


def test_periodic_argument_errors():
    X = np.random.default_rng(0).standard_normal((20, 2))
    y = X[:, 0]
    gp = HostGP(X, y, ["x0", "x1"])
    with pytest.raises(ValueError, match="Period must be specified"):
        gp.build_model(continuous_kernel="ExpQuad+Periodic")
    with pytest.raises(ValueError, match="Period must be specified"):
        gp.build_model(continuous_kernel="Periodic")
    with pytest.raises(NotImplementedError, match="one continuous dimension"):
        gp.build_model(continuous_kernel="ExpQuad+Periodic", period={"x0": 1.0, "x1": 2.0})
    with pytest.raises(ValueError, match="Continuous kernel"):
        gp.build_model(continuous_kernel="Periodic+Periodic", period={"x0": 1.0, "x1": 2.0})
    with pytest.raises(ValueError, match="non-zero"):
        gp.build_model(continuous_kernel="Periodic", period={"x0": 0.0, "x1": 2.0})


def test_inference_modes_left_to_pymc_raise():
    g = load_golden("simple_regression_ExpQuad")
    gp = gp_from_golden(g)
    with pytest.raises(NotImplementedError):
        gp.build_latent()
    with pytest.raises(NotImplementedError):
        gp.sample(100)
    gp.precision = "tf32"
    with pytest.raises(NotImplementedError, match="FITC"):
        gp.build_model(sparse=True)          # the sparse path is fp64, single GPU
    gp.precision = "fp64"
    with pytest.raises(NotImplementedError):
        gp.build_model(heteroskedastic_inputs=True)
