"""Cold-predict timing of the Kronecker-aware multi-output solve against the dense stacked solve, through the plugin class with
host buffers (the e2e arm of bench.py): BASELINE config 3 (n=16384, P=2, d=4, 2 x 10k-pt grid) by default.

    python tools/kron_timing.py [n] [P] [d] [steps]        # one GPU; prints one JSON line per solver + the max deviation
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import ArrayGP  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    spec, X, y, Xs = synthetic_problem(n, d, P=P)
    t = spec["terms"][0]
    point = {"ls_total": t["ls"], "η_total": t["eta"], "σ": spec["sigma"], "W_Variable": t["coreg"][0]["W"], "κ_Variable": t["coreg"][0]["kappa"],
             "W_Output_noise": spec["noise_coreg"]["W"], "κ_Output_noise": spec["noise_coreg"]["kappa"]}
    kw = dict(categorical_dims=["Variable"], categorical_levels={"Variable": [f"y{p}" for p in range(P)]}, outputs=[f"y{p}" for p in range(P)])
    out = {}
    for mode in ("kron", "kron-threads", "kron-fused", "kron-fused-threads", "dense-fused", "dense"):
        gp = ArrayGP(X, y, [f"x{j}" for j in range(d)], multioutput="dense" if mode.startswith("dense") else "kron", **kw)
        gp.build_model()
        fused = "fused" in mode                                        # one-pass factorise+predict per (block) engine
        if mode.endswith("threads"):
            gp.engine.threads = P                                      # all blocks in flight at once on this GPU
        gp.find_MAP(point=point)
        gp.predict_cold(Xs, fused=fused)                               # warm-up: allocations, module load
        t0 = time.perf_counter()
        for _ in range(steps):
            mu, var = gp.predict_cold(Xs, fused=fused)                 # K-build + Cholesky + solve + H2D/D2H, every call
        ms = (time.perf_counter() - t0) * 1e3 / steps
        t0 = time.perf_counter()
        gp.predict(Xs)
        warm = (time.perf_counter() - t0) * 1e3
        out[mode] = (mu, var)
        print(json.dumps({"solver": mode, "n": n, "P": P, "N": n * P, "d": d, "M": len(Xs), "cold_ms_per_step": ms, "warm_predict_ms": warm,
                          "predictions_per_s": len(Xs) / (ms * 1e-3)}), flush=True)
        gp.engine.close()
    dm = np.max(np.abs(out["kron"][0] - out["dense"][0]) / (1e-12 + np.abs(out["dense"][0])))
    dv = np.max(np.abs(out["kron"][1] - out["dense"][1]) / (1e-12 + np.abs(out["dense"][1])))
    print(json.dumps({"max_rel_dev_mean": dm, "max_rel_dev_var": dv}))


if __name__ == "__main__":
    main()
