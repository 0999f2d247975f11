"""Drop-in check against the REAL reference wrappers (only where /root/reference exists, i.e. the build container; the GPU box
replays the committed goldens instead).  ``make_backend(gumbi.regression.base.Regressor)`` must give a class that the
reference's own DataSet / specify_model / prepare_grid / predict_grid / cross-validation-style re-construction drive without
modification -- with the oracle standing in for the GPU engine (tests/ only)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gumbi")), reason="reference tree not present (the GPU box runs tests/test_dropin_gpu.py against baseline/_ref instead)")


@pytest.fixture(scope="module")
def ref_env():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_golden

    gmb = gen_golden.import_reference()
    from gumbi.regression.base import Regressor

    from gumbi_b200 import make_backend
    from test_backend_host import OracleEngine

    B200GP = make_backend(Regressor)

    class HostB200GP(B200GP):
        """The drop-in class with the engine double injected (no GPU in this container)."""

        def build_model(self, *a, **k):
            self.engine = OracleEngine()
            return super().build_model(*a, **k)

    import pandas as pd

    return gmb, HostB200GP, pd


def test_fit_predict_grid_through_the_reference_wrappers(ref_env):
    gmb, GP, pd = ref_env
    g = load_golden("simple_regression_ExpQuad")
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    gp = GP(ds, outputs=["d"])
    # README.md:28-36 flow: fit -> prepare_grid -> predict_grid
    gp.fit(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"], MAP_kwargs={"options": {"maxiter": 15}})
    assert isinstance(gp.MAP, dict) and {"ls_total", "η_total", "σ", "c_total", "τ_total"} <= set(gp.MAP)
    # the shaped arrays crossing the boundary are exactly the committed golden inputs (same wrappers, same data)
    np.testing.assert_array_equal(gp._X, g["X"])
    np.testing.assert_array_equal(gp._y, g["y"])
    gp.prepare_grid(at=gp.parray(lg10_Z=8, X=0.5))
    up = gp.predict_grid()
    assert type(up).__name__ == "UncertainParameterArray" and up.shape == (100,)
    assert np.all(np.isfinite(up.μ)) and np.all(up.σ2 > 0)
    # pinning the hyper-parameters reproduces the golden posterior through the reference's own point preparation
    gp.find_MAP(point=g["meta"]["point"])
    pts_grid, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=gp._parse_prediction_output(None))
    mu, var = gp.predict(pts_grid, with_noise=True)
    np.testing.assert_allclose(mu, g["mean"][1:], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g["var"][1:], rtol=1e-8, atol=1e-11)


def test_multioutput_predict_points_reads_W_and_kappa_from_MAP(ref_env):
    """base.py:585-599: the base class splits by the output column and builds the correlation from MAP['W_*'], MAP['κ_*']."""
    gmb, GP, pd = ref_env
    g = load_golden("multioutput_regression")
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    gp = GP(ds, outputs=["a", "b", "c", "d", "e"])
    gp.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp.build_model()
    gp.find_MAP(point=g["meta"]["point"])
    np.testing.assert_array_equal(gp._X, g["X"])
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=17)
    mv = gp.predict_grid()
    assert type(mv).__name__ == "MVUncertainParameterArray" and mv.shape == (17,)
    pts, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=gp._parse_prediction_output(None))
    np.testing.assert_array_equal(pts, g["points"])
    mu, var = gp.predict(pts)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    # cross_validate-style reconstruction (base.py:1060-1068): same class, model_specs round-trip
    gp2 = GP(ds, outputs=["a", "b", "c", "d", "e"])
    gp2.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp2.build_model(**gp.model_specs)
    assert gp2.model_specs == gp.model_specs


def test_draw_point_and_grid_samples_through_the_reference_wrappers(ref_env):
    """GP.py:861-979: draw_point_samples / draw_grid_samples return a ParameterArray of joint posterior draws."""
    gmb, GP, pd = ref_env
    g = load_golden("simple_regression_ExpQuad")
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    gp = GP(ds, outputs=["d"])
    gp.specify_model(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"])
    gp.build_model()
    gp.find_MAP(point=g["meta"]["point"])
    gp.prepare_grid(at=gp.parray(lg10_Z=8, X=0.5), resolution=25)
    samples = gp.draw_grid_samples(size=6, random_seed=3)
    assert type(samples).__name__ == "ParameterArray" and samples.shape == (6, 25)
    z = samples.z.values() if hasattr(samples.z, "values") else np.asarray(samples.z["d_z"])
    assert np.all(np.isfinite(np.asarray(z, dtype=float)))
    # same seed -> same draws; the draws scatter around the posterior mean of predict()
    again = gp.draw_grid_samples(size=6, random_seed=3)
    np.testing.assert_array_equal(np.asarray(again.z.values(), dtype=float), np.asarray(z, dtype=float))


def test_kron_solver_behind_the_reference_wrappers(ref_env):
    """multioutput="kron" (gumbi_b200/kron.py) under the real DataSet / get_shaped_data / predict_grid: the stacked arrays the
    reference builds for the multi-output notebook are aligned, and the block-wise solve returns the dense golden posterior."""
    gmb, GP, pd = ref_env
    from gumbi_b200 import kron
    from test_backend_host import OracleEngine

    class KronGP(GP):
        def build_model(self, *a, **k):
            self.engine = None
            return super(GP, self).build_model(*a, **k)      # skip the dense double injected by HostB200GP

        def _make_block_engine(self):
            return OracleEngine()

    g = load_golden("multioutput_regression")
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    gp = KronGP(ds, outputs=["a", "b", "c", "d", "e"], multioutput="kron")
    gp.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp.build_model()
    assert isinstance(gp.engine, kron.KronEngine) and gp.engine.n == 14 and gp.engine.P == 5
    gp.find_MAP(point=g["meta"]["point"])
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=17)
    mv = gp.predict_grid()
    assert type(mv).__name__ == "MVUncertainParameterArray" and mv.shape == (17,)
    pts, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=gp._parse_prediction_output(None))
    mu, var = gp.predict(pts)
    np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, g["var"], rtol=1e-8, atol=1e-11)
    # one output only (predict_grid(output="c")): still every block, one solve per distinct grid point
    up = gp.predict_grid(output="c")
    assert type(up).__name__ == "UncertainParameterArray" and up.shape == (17,)
    # fit() end to end through the block objective
    gp.find_MAP(options={"maxiter": 10})
    assert np.isfinite(gp.marginal_log_likelihood())


def test_class_defaults_survive_reconstruction(ref_env):
    """cross_validate re-creates the regressor with ``self.__class__(train_ds, outputs=..., seed=...)`` (base.py:1060): options given
    to make_backend are class defaults and therefore survive; unknown options are rejected like any unexpected keyword."""
    gmb, GP, pd = ref_env
    from gumbi.regression.base import Regressor

    from gumbi_b200 import make_backend

    Cls = make_backend(Regressor, device=3, precision="tf32", multioutput="auto")
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    gp = Cls(ds, outputs=["d"])
    again = gp.__class__(ds, outputs=gp.outputs, seed=7)
    assert (again.device, again.precision, again.multioutput, again.seed) == (3, "tf32", "auto", 7)
    assert Cls(ds, outputs=["d"], device=1).device == 1
    with pytest.raises(TypeError):
        Cls(ds, outputs=["d"], gpu=1)


def test_cross_validate_runs_unmodified_on_the_backend(ref_env):
    """Regressor.cross_validate (base.py:844-1109) re-creates the class on a training subset, calls build_model(**model_specs),
    find_MAP(**MAP_kws), predict_points on train and test rows, and scores them -- all inherited, nothing backend-specific."""
    gmb, GP, pd = ref_env
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    gp = GP(ds, outputs=["d"])
    gp.fit(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"], MAP_kwargs={"options": {"maxiter": 10}})
    cv = gp.cross_validate(n_train=50, seed=1, options={"maxiter": 10})
    assert set(cv) == {"train", "test"}
    for part in cv.values():
        assert np.all(np.isfinite(np.asarray(part["NLPDs"], dtype=float))) and np.all(np.isfinite(np.asarray(part["errors"], dtype=float)))


def test_ls_bounds_as_a_parameter_array(ref_env):
    """fit(ls_bounds=parray) (GP.py:630-646): the bounds are read through ``ls_bounds[dim].z.values()`` exactly as the reference does."""
    gmb, GP, pd = ref_env
    from gumbi_b200.map import ls_bounds_z

    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    gp = GP(ds, outputs=["d"])
    gp.specify_model(continuous_dims="lg10_Z")
    bounds = gp.parray(lg10_Z=[4.0, 9.0])
    gp.build_model(ls_bounds=bounds)
    lower, upper = ls_bounds_z(gp)
    np.testing.assert_allclose([lower[0], upper[0]], np.asarray(bounds["lg10_Z"].z.values()).squeeze())
    gp.find_MAP(options={"maxiter": 5})
    assert np.isfinite(gp.MAP["ls_total"]).all()


def test_periodic_kernel_with_a_parameter_array_period(ref_env):
    """build_model(continuous_kernel="ExpQuad+Periodic", period=parray) (GP.py:403, :429): the period is read as
    ``period.z[dim + "_z"]`` like the reference does, and predict_grid runs through the warped-coordinate lowering."""
    gmb, GP, pd = ref_env
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    for kernel in ("ExpQuad+Periodic", "Periodic"):
        gp = GP(ds, outputs=["d"])
        gp.specify_model(continuous_dims="lg10_Z")
        period = gp.parray(lg10_Z=12.0)
        gp.build_model(continuous_kernel=kernel, period=period)
        zp = float(np.asarray(period.z["lg10_Z_z"].values()).reshape(-1)[0])
        np.testing.assert_allclose(gp._layout["warp"]["c"], [2 * np.pi / zp])
        gp.find_MAP(options={"maxiter": 8})
        gp.prepare_grid(resolution=11)
        up = gp.predict_grid()
        assert up.shape == (11,) and np.all(np.isfinite(up.μ)) and np.all(up.σ2 > 0)
