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

    def mll(self):
        return orc.mll(self.spec, self.X, self.y)


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
    with pytest.raises(NotImplementedError):
        gp.build_model(sparse=True)
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
