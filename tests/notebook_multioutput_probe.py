"""Probe (container only, needs /root/reference): replay docs/source/notebooks/examples/Multioutput_Regression.ipynb through the
reference's wrappers with the numpy oracle as the engine and compare with the executed cell output (cell 16)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gen_golden import REF, import_reference  # noqa: E402

gmb = import_reference()
import pandas as pd  # noqa: E402
from gumbi.regression.base import Regressor  # noqa: E402
from test_backend_host import OracleEngine  # noqa: E402

from gumbi_b200.backend import B200Backend  # noqa: E402


class HostRefGP(B200Backend, Regressor):
    def __init__(self, dataset, outputs=None, seed=2021):
        Regressor.__init__(self, dataset, outputs, seed)
        self._init_backend()
        self.engine = OracleEngine()


NB_MU = np.array([[-9.59442479, 0.65605058, 0.00646403, 0.81416271, 0.15214448],
                  [-8.05656298, 0.66609041, 0.00635764, 0.81267686, 0.16440518],
                  [-6.40414117, 0.67787309, 0.00620618, 0.8105507, 0.17662809],
                  [-4.75033515, 0.68729924, 0.00617787, 0.81008143, 0.19510875],
                  [-2.94787273, 0.69658766, 0.00619329, 0.81021742, 0.21940875]])
NB_S2 = np.array([[0.01676639, 1.22016392e-04, 0.00134064, 1.22105406e-05, 2.80013388e-04],
                  [0.00462947, 1.03825281e-04, 0.00114775, 1.03319592e-05, 8.16338238e-05],
                  [0.00455407, 9.92785623e-05, 0.00110227, 9.91092000e-06, 8.03754540e-05],
                  [0.00462973, 1.04073081e-04, 0.00115023, 1.03548470e-05, 8.16411508e-05],
                  [0.01685376, 1.22594426e-04, 0.00134647, 1.22658105e-05, 2.81471085e-04]])


def main(**fit_kw):
    df = pd.read_pickle(os.path.join(REF, "gumbi", "data", "Example_DataSet.pkl"))
    df = df[(df.Name == "binary-pollen") & (df.Color == "cyan") & (df.Metric == "mean")]
    ds = gmb.DataSet(df, outputs=["a", "b", "c", "d", "e", "f"], log_vars=["Y", "b", "c", "d", "f"], logit_vars=["X", "e"])
    fit_params = ["a", "b", "c", "d", "e"]
    gp = HostRefGP(ds, outputs=fit_params)
    gp.specify_model(continuous_dims="lg10_Z", linear_dims="lg10_Z")
    gp.build_model()
    gp.find_MAP(**fit_kw)
    gp.prepare_grid(limits=gp.parray(lg10_Z=[1, 9]), resolution=5)
    gp.predict_grid()
    mv = gp.predictions
    raw = np.asarray(mv)
    mu = np.column_stack([raw["μ"][p] for p in fit_params])
    s2 = np.column_stack([raw["σ2"][p] for p in fit_params])
    return gp, mu, s2


if __name__ == "__main__":
    gp, mu, s2 = main()
    np.set_printoptions(precision=6, linewidth=160)
    print("MAP:", {k: np.round(np.asarray(v), 5).tolist() for k, v in gp.MAP.items() if not k.endswith("__")})
    print("mu rel err:\n", mu / NB_MU - 1)
    print("s2 rel err:\n", s2 / NB_S2 - 1)
