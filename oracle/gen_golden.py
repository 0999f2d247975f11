"""Generate tests/golden/*.npz  --  TEST INFRASTRUCTURE ONLY; run in the build container (needs /root/reference).

For each case the reference's OWN wrapper layers (DataSet / Standardizer / Regressor.specify_model /
get_shaped_data / prepare_grid / _prepare_points_for_prediction, imported unmodified from /root/reference with the
plotting + PyMC imports stubbed) produce the standardized arrays that cross the backend boundary
(gumbi/regression/base.py:574), the numpy oracle (oracle/gp_oracle.py) produces the posterior at a fixed,
seed-derived hyper-parameter point, and the reference's predict_points post-processing (base.py:578-601) turns that
into un-standardized uparray fields.  Everything is stored so that the GPU box (no reference, no gumbi) can replay it.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz
"""
from __future__ import annotations

import json
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def reference_root():
    """Where the UNMODIFIED reference package can be imported from: the read-only source tree in the build container, else the
    dependency-less install ``baseline/_ref`` (``pip install --no-deps --target baseline/_ref <copy of /root/reference>``; git-ignored,
    travels to the GPU box with the snapshot).  None if neither exists."""
    for cand in (REF, os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "gumbi")):
            return cand
    return None


def import_reference(ref=None):
    """Stub the absent third-party imports, then import gumbi from the reference tree (or its install, see reference_root)."""
    if "gumbi" in sys.modules:
        return sys.modules["gumbi"]
    ref = ref or reference_root() or REF

    def stub(name):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__getattr__ = lambda attr: MagicMock()
        sys.modules[name] = m
        return m

    for n in ["matplotlib", "matplotlib.pyplot", "seaborn", "uncertainties", "uncertainties.unumpy", "pymc", "pytensor",
              "pytensor.tensor", "gpytorch", "gpytorch.priors", "gpytorch.priors.prior", "gpytorch.priors.utils",
              "botorch"]:
        try:
            __import__(n)
        except Exception:
            stub(n)
    if isinstance(getattr(sys.modules.get("gpytorch.priors.prior"), "Prior", None), MagicMock):
        sys.modules["gpytorch.priors.prior"].Prior = type("Prior", (), {})
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gumbi
    return gumbi


def fixed_point(shapes, seed):
    """Deterministic hyper-parameters of plausible magnitude (not a MAP: the oracle is evaluated AT this point)."""
    rng = np.random.default_rng(seed)
    pt = {}
    for name, shape in shapes.items():
        base = name.split("_")[0]
        if base == "ls":
            pt[name] = rng.uniform(0.7, 2.5, size=shape)
        elif base == "η":
            pt[name] = np.asarray(rng.uniform(0.8, 1.6))
        elif base == "c":
            pt[name] = rng.normal(0, 0.5, size=shape)
        elif base == "τ":
            pt[name] = np.asarray(rng.uniform(0.05, 0.5))
        elif base == "W":
            pt[name] = rng.standard_normal(size=shape) * (0.3 if "noise" in name else 1.0)
        elif base == "κ":
            pt[name] = rng.uniform(0.5, 1.5, size=shape)
        elif base == "σ":
            pt[name] = np.asarray(rng.uniform(0.08, 0.3))
        else:
            raise KeyError(name)
    return pt


def main():
    sys.path.insert(0, ROOT)
    gmb = import_reference()
    import pandas as pd
    from gumbi.regression.base import Regressor

    from gumbi_b200.backend import B200Backend
    from oracle import gp_oracle as orc

    class OracleEngine:
        """Stands where GPEngine stands, answers with the numpy oracle (generation only -- never shipped)."""

        def set_train(self, X, y):
            self.X, self.y = X, y

        def set_kernel(self, spec):
            self.spec = spec

        def factorize(self):
            self.L, self.v = orc.factorize(self.spec, self.X, self.y)

        def predict(self, Xs, pred_noise=True):
            return orc.conditional(self.spec, self.X, self.L, self.v, Xs, pred_noise)

        def mll(self):
            return orc.mll(self.spec, self.X, self.y)

    class GoldenGP(B200Backend, Regressor):
        def __init__(self, dataset, outputs=None, seed=2021):
            Regressor.__init__(self, dataset, outputs, seed)
            self._init_backend()
            self.engine = OracleEngine()

    os.makedirs(OUT, exist_ok=True)

    def dump(name, gp, point, points_array, extra):
        spec = gp.spec_from_point(gp.MAP)
        mu, var = gp.predict(points_array, with_noise=True)
        mu_nf, var_nf = gp.predict(points_array, with_noise=False)
        meta = {
            "continuous_dims": gp.continuous_dims, "linear_dims": gp.linear_dims, "categorical_dims": gp.categorical_dims,
            "categorical_levels": {k: list(map(str, v)) for k, v in gp.categorical_levels.items()},
            "categorical_coords": {k: {str(a): int(b) for a, b in v.items()} for k, v in gp.categorical_coords.items()},
            "out_col": gp.out_col, "outputs": gp.outputs, "additive": bool(gp.additive),
            "continuous_kernel": gp.continuous_kernel, "ARD": bool(gp.ARD), "spec": spec,
            "point": {k: np.asarray(v).tolist() for k, v in point.items()},
            "generator": "oracle/gen_golden.py", "reference_commit": "27bdbee",
        }
        np.savez_compressed(os.path.join(OUT, name + ".npz"), X=gp._X, y=gp._y, points=points_array, mean=mu, var=var,
                            mean_noisefree=mu_nf, var_noisefree=var_nf, mll=np.asarray(gp.marginal_log_likelihood()),
                            meta=np.asarray(json.dumps(meta, ensure_ascii=False)), **extra)
        print(f"{name}: X{gp._X.shape} points{points_array.shape} mll={gp.marginal_log_likelihood():.6f}")

    # ---- 1. Simple_Regression notebook (docs/source/notebooks/examples/Simple_Regression.pct.py:34-63) -------------
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    outputs = ["a", "b", "c", "d", "e", "f"]
    log_vars = ["Y", "b", "c", "d", "f"]
    logit_vars = ["X", "e"]
    ds = gmb.DataSet(df, outputs=outputs, log_vars=log_vars, logit_vars=logit_vars)
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    for kern in ("ExpQuad", "Matern52"):
        gp = GoldenGP(ds, outputs=["d"])
        gp.specify_model(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"])
        gp.build_model(continuous_kernel=kern)
        point = fixed_point(gp.param_shapes(), 11)
        gp.find_MAP(point=point)
        gp.prepare_grid(at=gp.parray(lg10_Z=8, X=0.5))
        pts_grid, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=gp._parse_prediction_output(None))
        pts_one, _, _ = gp._prepare_points_for_prediction(gp.parray(lg10_Z=8, X=0.5, Y=88), output=gp._parse_prediction_output(None))
        points_array = np.vstack([pts_one, pts_grid])
        # reference post-processing of the same predictions (base.py:578-601): natural-space mean and z-space fields
        up = gp.predict_points(gp.grid_points)
        extra = {"post_mu_z": np.asarray(up.z["μ"] if "μ" in up.z.dtype.names else up["μ"]),
                 "post_mu_natural": np.asarray(up["μ"]), "post_sigma2": np.asarray(up["σ2"])}
        dump(f"simple_regression_{kern}", gp, point, points_array, extra)

    # ---- 2. Multioutput_Regression notebook (Multioutput_Regression.pct.py:40-100) ---------------------------------
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=outputs, log_vars=log_vars, logit_vars=logit_vars)
    fit_params = ["a", "b", "c", "d", "e"]
    gp = GoldenGP(ds, outputs=fit_params)
    gp.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp.build_model()
    point = fixed_point(gp.param_shapes(), 12)
    gp.find_MAP(point=point)
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=17)
    out = gp._parse_prediction_output(None)
    points_array, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=out)
    dump("multioutput_regression", gp, point, points_array, {})

    # ---- 3. reference test fixture (tests/test_regression.py:13-45, :94-112, :158-166) ------------------------------
    example_stdzr = {
        "a": {"μ": -0.762, "σ2": 1.258 ** 2}, "b": {"μ": -0.0368, "σ2": 0.351 ** 2}, "c": {"μ": -5.30, "σ2": 0.582 ** 2},
        "d": {"μ": -0.307, "σ2": 0.158 ** 2}, "e": {"μ": -1.056, "σ2": 0.398 ** 2}, "f": {"μ": 3.34, "σ2": 0.1501 ** 2},
        "X": {"μ": -0.282, "σ2": 1 ** 2}, "Y": {"μ": 4.48, "σ2": 0.75 ** 2}, "lg10_Z": {"μ": 5, "σ2": 2 ** 2},
    }
    es = pd.read_pickle(os.path.join(REF, "tests", "test_data", "test_dataset.pkl"))
    stdzr = gmb.Standardizer(**example_stdzr, log_vars=["d", "f", "b", "c", "Y"], logit_vars=["e", "X"])
    ds = gmb.DataSet.from_tidy(es, names_column="Parameter", stdzr=stdzr)
    # Joint and additive categorical structure with two outputs (the shape of tests/test_regression.py:158-166).  The
    # reference test uses the NUMERIC column lg10_Z as the categorical dim; its coords are then the level values
    # themselves, get z-scored by get_shaped_data (base.py:464) and truncated by Coregion's int32 cast into indices
    # {1,0,0,-1}: the test only asserts "runs".  A string-valued categorical dim (Color) gives the 0..P-1 coords the
    # Coregion kernel is meant to see, so that is what the golden cases pin.
    dfc = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    dfc = dfc[(dfc.Pair == "burrata+merlot") & (dfc.Metric == "mean")]
    dsc = gmb.DataSet(dfc, outputs=outputs, log_vars=log_vars, logit_vars=logit_vars)
    for additive in (False, True):
        gp = GoldenGP(dsc, outputs=["d", "c"])
        gp.specify_model(outputs=["d", "c"], continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["Y"] if additive else None,
                         categorical_dims="Color", additive=additive)
        gp.build_model(continuous_kernel="Matern32" if additive else "ExpQuad")
        point = fixed_point(gp.param_shapes(), 13)
        gp.find_MAP(point=point)
        ygrid = np.geomspace(12.0, 160.0, 20)
        blocks = []
        for lvl, coord in gp.categorical_coords["Color"].items():
            pts = gp.parray(X=np.full(20, 0.5), Y=ygrid, lg10_Z=np.full(20, 8.0), Color=np.full(20, float(coord)), stdzd=False)
            pa, _, _ = gp._prepare_points_for_prediction(pts, output=gp._parse_prediction_output(None))
            blocks.append(pa)
        points_array = np.vstack(blocks)
        dump("categorical_additive" if additive else "categorical_joint", gp, point, points_array, {})

    # single-level filter case: tests/test_regression.py:108-112 (continuous_levels filters lg10_Z to one level)
    gp = GoldenGP(ds, outputs="d")
    gp.specify_model(continuous_dims=["X", "Y", "lg10_Z"], continuous_levels={"lg10_Z": [8]})
    gp.build_model(ARD=False)
    point = fixed_point(gp.param_shapes(), 14)
    gp.find_MAP(point=point)
    gp.prepare_grid(resolution=15)
    points_array, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=gp._parse_prediction_output(None))
    dump("test_dataset_filtered", gp, point, points_array, {})


def notebook_fixture():
    """tests/golden/notebook_simple_regression.npz: everything needed to replay docs/source/notebooks/examples/Simple_Regression
    (script Simple_Regression.pct.py:33-63) WITHOUT gumbi: the shaped training arrays, the standardized prediction points and the
    constants of the output transform -- plus the numbers the reference's own executed notebook shows (PyMC run by the author):
    Simple_Regression.ipynb:192 (point prediction) and :230-234 (first 10 grid predictions)."""
    sys.path.insert(0, ROOT)
    gmb = import_reference()
    import pandas as pd
    from gumbi.regression.base import Regressor

    from gumbi_b200.backend import B200Backend

    class ShapeOnly(B200Backend, Regressor):
        def __init__(self, dataset, outputs=None, seed=2021):
            Regressor.__init__(self, dataset, outputs, seed)
            self._init_backend()

    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl")).query('Metric=="mean"')
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    ds.tidy = ds.tidy[ds.tidy.Color.isin(["cyan", "magenta"]) & (ds.tidy.Pair == "burrata+barbaresco")]
    gp = ShapeOnly(ds, outputs=["d"])
    gp.specify_model(continuous_dims=["X", "Y", "lg10_Z"], linear_dims=["X", "Y", "lg10_Z"])
    X, y = gp.get_shaped_data("mean")
    out = gp._parse_prediction_output(None)
    pt, _, _ = gp._prepare_points_for_prediction(gp.parray(lg10_Z=8, X=0.5, Y=88), output=out)
    gp.prepare_grid(at=gp.parray(lg10_Z=8, X=0.5))
    grid, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=out)
    mu_d, s2_d = float(gp.stdzr["d"]["μ"]), float(gp.stdzr["d"]["σ2"])
    expected_point = np.array([0.7526282, 0.00204789])                                      # Simple_Regression.ipynb:192
    expected_grid = np.array([[0.95353955, 0.02777067], [0.94923129, 0.02648205], [0.94544874, 0.02492182], [0.94220088, 0.02307904],
                              [0.93948256, 0.02096868], [0.93727268, 0.01863859], [0.93553307, 0.01617249], [0.93420812, 0.01368664],
                              [0.93322533, 0.01131927], [0.93249681, 0.009213]])           # Simple_Regression.ipynb:230-234
    meta = {"continuous_dims": ["X", "Y", "lg10_Z"], "linear_dims": ["X", "Y", "lg10_Z"], "output": "d", "output_transform": "log",
            "stdzr_mu": mu_d, "stdzr_sigma2": s2_d, "seed": 2021,
            "source": "docs/source/notebooks/examples/Simple_Regression.ipynb:192,230-234 (reference commit 27bdbee)",
            "post_processing": "mu_natural = exp(mu_z*sqrt(sigma2) + mu); sigma2_reported = var_z*sigma2  (base.py:580, aggregation.py:469-485)"}
    np.savez_compressed(os.path.join(OUT, "notebook_simple_regression.npz"), X=X, y=y, point=pt, grid=grid[:10], expected_point=expected_point,
                        expected_grid=expected_grid, meta=np.asarray(json.dumps(meta, ensure_ascii=False)))
    print("notebook_simple_regression: X", X.shape, "stdzr d", mu_d, s2_d)


def notebook_multioutput_fixture():
    """tests/golden/notebook_multioutput_regression.npz: shaped arrays + output transforms to replay
    docs/source/notebooks/examples/Multioutput_Regression (script Multioutput_Regression.pct.py:40-100) WITHOUT gumbi, plus the
    5-output x 5-point ``mvuparray`` its executed cell shows (Multioutput_Regression.ipynb:270-274): the only Coregion output the
    reference tree holds.  That cell was executed with the lengthscale prior ``pm.Gamma("ls", alpha=2, beta=1)`` -- the line that
    is still in the source as a comment (gumbi/regression/pymc/GP.py:408), one line below the InverseGamma prior of today's code:
    with the current prior the replay misses the cell by 0.5 % (mean) / 4 % (variance), with Gamma(2, 1) and everything else as in
    today's code it reproduces it to 1e-5 / 2.5e-4 (tests/test_notebook_parity.py; bisect recorded in DESIGN.md)."""
    sys.path.insert(0, ROOT)
    gmb = import_reference()
    import pandas as pd
    from gumbi.regression.base import Regressor

    from gumbi_b200.backend import B200Backend

    class ShapeOnly(B200Backend, Regressor):
        def __init__(self, dataset, outputs=None, seed=2021):
            Regressor.__init__(self, dataset, outputs, seed)
            self._init_backend()

    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    fit_params = ["a", "b", "c", "d", "e"]
    gp = ShapeOnly(ds, outputs=fit_params)
    gp.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    X, y = gp.get_shaped_data("mean")
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=5)
    out = gp._parse_prediction_output(None)
    grid, _, _ = gp._prepare_points_for_prediction(gp.grid_points, output=out)     # 5 grid points x 5 outputs, output-major (base.py:533-536)
    mu = [float(gp.stdzr[p]["μ"]) for p in fit_params]
    s2 = [float(gp.stdzr[p]["σ2"]) for p in fit_params]
    transforms = ["log" if p in ds.stdzr.log_vars else ("logit" if p in ds.stdzr.logit_vars else "none") for p in fit_params]
    expected_mu = np.array([[-9.59442479, 0.65605058, 0.00646403, 0.81416271, 0.15214448],
                            [-8.05656298, 0.66609041, 0.00635764, 0.81267686, 0.16440518],
                            [-6.40414117, 0.67787309, 0.00620618, 0.8105507, 0.17662809],
                            [-4.75033515, 0.68729924, 0.00617787, 0.81008143, 0.19510875],
                            [-2.94787273, 0.69658766, 0.00619329, 0.81021742, 0.21940875]])
    expected_s2 = np.array([[0.01676639, 1.22016392e-04, 0.00134064, 1.22105406e-05, 2.80013388e-04],
                            [0.00462947, 1.03825281e-04, 0.00114775, 1.03319592e-05, 8.16338238e-05],
                            [0.00455407, 9.92785623e-05, 0.00110227, 9.91092000e-06, 8.03754540e-05],
                            [0.00462973, 1.04073081e-04, 0.00115023, 1.03548470e-05, 8.16411508e-05],
                            [0.01685376, 1.22594426e-04, 0.00134647, 1.22658105e-05, 2.81471085e-04]])
    meta = {"continuous_dims": ["lg10_Z"], "linear_dims": ["lg10_Z"], "categorical_dims": [gp.out_col], "out_col": gp.out_col,
            "categorical_levels": {gp.out_col: list(gp.categorical_levels[gp.out_col])}, "outputs": fit_params, "seed": 2021,
            "output_order_of_points": list(out), "stdzr_mu": mu, "stdzr_sigma2": s2, "output_transforms": transforms,
            "source": "docs/source/notebooks/examples/Multioutput_Regression.ipynb:270-274 (reference commit 27bdbee)",
            "ls_prior_of_the_executed_cell": "Gamma(alpha=2, beta=1) -- gumbi/regression/pymc/GP.py:408 (commented out in today's code)",
            "post_processing": "mu_natural = T^-1(mu_z*sqrt(sigma2) + mu), T = log / logit / identity; sigma2_reported = var_z*sigma2 (base.py:580-599)"}
    np.savez_compressed(os.path.join(OUT, "notebook_multioutput_regression.npz"), X=X, y=y, grid=grid, expected_mu=expected_mu,
                        expected_s2=expected_s2, meta=np.asarray(json.dumps(meta, ensure_ascii=False)))
    print("notebook_multioutput_regression: X", X.shape, "grid", grid.shape, "transforms", transforms)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "notebook":
        notebook_fixture()
        notebook_multioutput_fixture()
    else:
        main()
        notebook_fixture()
        notebook_multioutput_fixture()
