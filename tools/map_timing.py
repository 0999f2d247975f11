"""find_MAP on the device at a non-trivial size (dev tool): wall time per objective evaluation and the optimum reached."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gumbi_b200 import ArrayGP
from gumbi_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec, X, y, Xs = synthetic_problem(n, 4, M_res=30)
gp = ArrayGP(X, y, [f"x{j}" for j in range(4)])
gp.build_model()
t0 = time.perf_counter()
MAP = gp.find_MAP(options={"maxiter": 60})
dt = time.perf_counter() - t0
print({"N": n, "evals": gp.map_evals, "s_total": round(dt, 3), "ms_per_eval": round(1e3 * dt / gp.map_evals, 2), "fun": float(gp.map_result.fun),
       "ls": MAP["ls_total"].round(3).tolist(), "eta": float(MAP["η_total"]), "sigma": float(MAP["σ"]), "msg": str(gp.map_result.message)})
mu, var = gp.predict(Xs)
print("rmse vs truth-free check: mean|mu|", float(np.abs(mu).mean()), "var range", float(var.min()), float(var.max()))
