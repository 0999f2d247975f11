"""The REAL ``gumbi.regression.base.Regressor`` driving the backend end to end through the reference's public API --
``DataSet`` -> ``gp.fit`` -> ``predict_points`` / ``prepare_grid`` / ``predict_grid`` -> ``UncertainParameterArray`` -- and landing on the
numbers of the reference's own executed notebooks (Simple_Regression.ipynb:192,230-234; Multioutput_Regression.ipynb:270-274).

Two engines behind the same class: the numpy oracle (CPU, wherever the reference is importable) and the CUDA engine (``-m gpu``:
Regressor + C ABI + kernels in ONE process).  The reference is imported unmodified from /root/reference (build container) or from
its dependency-less install ``baseline/_ref`` (travels to the GPU box; created by
``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>``), with the absent
third-party imports (PyMC, matplotlib, ...) stubbed exactly as oracle/gen_golden.py does.  Skipped where neither exists."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_golden  # noqa: E402

REFROOT = gen_golden.reference_root()
pytestmark = pytest.mark.skipif(REFROOT is None, reason="reference package not importable here (no /root/reference, no baseline/_ref)")

ENGINES = [pytest.param("oracle", id="oracle-engine"), pytest.param("cuda", id="cuda-engine", marks=pytest.mark.gpu)]


def backend_class(engine):
    gmb = gen_golden.import_reference(REFROOT)
    from gumbi.regression.base import Regressor

    from gumbi_b200 import make_backend

    B200GP = make_backend(Regressor)
    if engine == "cuda":
        return gmb, B200GP
    from test_backend_host import OracleEngine

    class HostB200GP(B200GP):
        def build_model(self, *a, **k):
            self.engine = OracleEngine()
            return super().build_model(*a, **k)

    return gmb, HostB200GP


def example_frame():
    import pandas as pd

    return pd.read_pickle(os.path.join(REFROOT, "gumbi", "data", "Example_DataSet.pkl"))


@pytest.mark.parametrize("engine", ENGINES)
def test_simple_regression_notebook_through_the_public_api(engine, request):
    if engine == "cuda":
        request.getfixturevalue("lib_built")
    gmb, GP = backend_class(engine)
    df = example_frame().query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]     # Simple_Regression.pct.py:34-45
    gp = GP(ds, outputs=["d"])
    gp.fit(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"])
    up = gp.predict_points(gp.parray(lg10_Z=8, X=0.5, Y=88))
    assert type(up).__name__ == "UncertainParameterArray"
    np.testing.assert_allclose(float(np.asarray(up.μ).squeeze()), 0.7526282, rtol=2e-5)        # Simple_Regression.ipynb:192
    np.testing.assert_allclose(float(np.asarray(up.σ2).squeeze()), 0.00204789, rtol=2e-4)
    gp.prepare_grid(at=gp.parray(lg10_Z=8, X=0.5))
    grid = gp.predict_grid()
    expected = np.array([[0.95353955, 0.02777067], [0.94923129, 0.02648205], [0.94544874, 0.02492182], [0.94220088, 0.02307904],
                         [0.93948256, 0.02096868], [0.93727268, 0.01863859], [0.93553307, 0.01617249], [0.93420812, 0.01368664],
                         [0.93322533, 0.01131927], [0.93249681, 0.009213]])                    # Simple_Regression.ipynb:230-234
    np.testing.assert_allclose(np.asarray(grid.μ, dtype=float)[:10], expected[:, 0], rtol=2e-5)
    np.testing.assert_allclose(np.asarray(grid.σ2, dtype=float)[:10], expected[:, 1], rtol=2e-4)
    if engine == "cuda":
        assert type(gp.engine).__name__ == "GPEngine" and gp.engine.timings()["launches_factorize"] > 0
        gp.engine.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_multioutput_notebook_through_the_public_api(engine, request):
    if engine == "cuda":
        request.getfixturevalue("lib_built")
    gmb, GP = backend_class(engine)
    df = example_frame()
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]                  # Multioutput_Regression.pct.py:40-60
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    fit_params = ["a", "b", "c", "d", "e"]
    gp = GP(ds, outputs=fit_params)
    gp.ls_prior = "Gamma(2,1)"            # the lengthscale prior the executed cell was produced with (GP.py:408; tests/test_notebook_parity.py)
    gp.fit(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=5)
    mv = gp.predict_grid()
    assert type(mv).__name__ == "MVUncertainParameterArray" and mv.shape == (5,)
    raw = np.asarray(mv)
    mu = np.column_stack([raw["μ"][p] for p in fit_params])
    s2 = np.column_stack([raw["σ2"][p] for p in fit_params])
    from test_notebook_parity import MO_RTOL_MEAN, MO_RTOL_VAR, load_golden

    g = load_golden("notebook_multioutput_regression")
    np.testing.assert_allclose(mu, g["expected_mu"], rtol=MO_RTOL_MEAN)                          # Multioutput_Regression.ipynb:270-274
    np.testing.assert_allclose(s2, g["expected_s2"], rtol=MO_RTOL_VAR)
    assert mv.cor.shape[-2:] == (5, 5) if hasattr(mv, "cor") else True
    if engine == "cuda":
        gp.engine.close()


@pytest.mark.parametrize("engine", ENGINES)
def test_cross_validate_and_sparse_fit_through_the_public_api(engine, request):
    if engine == "cuda":
        request.getfixturevalue("lib_built")
    gmb, GP = backend_class(engine)
    df = example_frame().query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    gp = GP(ds, outputs=["d"])
    gp.fit(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"], MAP_kwargs={"options": {"maxiter": 10}})
    cv = gp.cross_validate(n_train=50, seed=1, options={"maxiter": 10})                         # base.py:844-1109, inherited unchanged
    assert set(cv) == {"train", "test"}
    for part in cv.values():
        assert np.all(np.isfinite(np.asarray(part["NLPDs"], dtype=float)))
    # sparse=True (GP.py:571-578): FITC behind the same public calls
    gps = GP(ds, outputs=["d"])
    gps.fit(continuous_dims=["X", "Y", "lg10_Z"], sparse=True, n_u=40, MAP_kwargs={"options": {"maxiter": 8}})
    up = gps.predict_points(gps.parray(lg10_Z=8, X=0.5, Y=88))
    assert np.isfinite(float(np.asarray(up.μ).squeeze())) and float(np.asarray(up.σ2).squeeze()) > 0
    if engine == "cuda":
        gp.engine.close(); gps.engine.close()
