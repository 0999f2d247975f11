"""Cost of one find_MAP objective evaluation (factorise + value + gradient) at a given size (dev tool): python tools/grad_timing.py [N] [d] [kind]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from gumbi_b200 import GPEngine
from gumbi_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = sys.argv[3] if len(sys.argv) > 3 else "Matern52"
spec, X, y, Xs = synthetic_problem(n, d, kind=kind, M_res=10)
e = GPEngine()
e.set_train(X, y)
e.set_kernel(spec)
out = []
for it in range(3):
    t0 = time.perf_counter(); e.factorize(); t1 = time.perf_counter(); mll, g = e.mll_grad(spec); t2 = time.perf_counter()
    out.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
print({"N": n, "d": d, "kind": kind, "factorize_ms": [round(a, 1) for a, _ in out], "mll_grad_ms": [round(b, 1) for _, b in out], "mll": float(mll),
       "d_ls": np.asarray(g["terms"][0]["ls"]).round(4).tolist(), "d_eta": g["terms"][0]["eta"], "d_sigma": g["sigma"]})
