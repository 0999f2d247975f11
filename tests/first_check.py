"""First GPU correctness run: CUDA path vs oracle on small/medium problems (dev tool, not a test)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from oracle import gp_oracle as orc
from gumbi_b200 import GPEngine

def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
def relmax(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))

eng = GPEngine(0)
for (n, d, P, kind, Q) in [(100, 1, 1, "ExpQuad", 1), (392, 3, 1, "Matern52", 1), (300, 2, 3, "ExpQuad", 1), (1000, 8, 1, "ExpQuad", 1),
                            (700, 4, 2, "Matern32", 2), (2048, 8, 1, "Matern12", 1), (1500, 5, 1, "Exponential", 1)]:
    spec, X, y, Xs = orc.synthetic_problem(n, d, P=P, M_res=20, kind=kind, Q=Q)
    if d >= 2:
        spec["terms"][0]["lin_idx"] = [0, 1]; spec["terms"][0]["c"] = [0.3, -0.2]; spec["terms"][0]["tau"] = 0.05
    if P > 1:
        spec["noise_coreg"]["W"] = (0.3 * np.random.default_rng(1).standard_normal((P, 2))).tolist()
    eng.set_train(X, y); eng.set_kernel(spec)
    K = eng.get_K()
    K0 = orc.train_cov(spec, X)
    eng.factorize()
    L = eng.get_L(); v = eng.get_v()
    L0, v0 = orc.factorize(spec, X, y)
    mu, var = eng.predict(Xs, True)
    mu0, var0 = orc.conditional(spec, X, L0, v0, Xs, True)
    print(json.dumps({"n": n, "d": d, "P": P, "kind": kind, "Q": Q, "K": relmax(K, K0), "L": relmax(L, L0), "v": relmax(v, v0),
                      "mean": relmax(mu, mu0), "var_rel": rel(var, var0), "mll": [eng.mll(), orc.mll(spec, X, y)], "t": eng.timings()}))
# non-PD
spec, X, y, Xs = orc.synthetic_problem(200, 2)
spec["sigma"] = 0.0; spec["jitter"] = 0.0
X[5] = X[4]
eng.set_train(X, y); eng.set_kernel(spec)
try:
    eng.factorize(); print("non-PD: no error?!")
except np.linalg.LinAlgError as e:
    print("non-PD ok:", e)
# medium timing
for n in (4096, 8192):
    spec, X, y, Xs = orc.synthetic_problem(n, 8, M_res=100)
    eng.set_train(X, y); eng.set_kernel(spec)
    for it in range(3):
        t0 = time.time(); eng.factorize(); t1 = time.time(); mu, var = eng.predict(Xs, True); t2 = time.time()
    tm = eng.timings()
    print(json.dumps({"n": n, "wall_fact_ms": (t1 - t0) * 1e3, "wall_pred_ms": (t2 - t1) * 1e3, "chol_tflops": n**3 / 3 / tm["cholesky_ms"] / 1e9,
                      "solve_tflops": n * n * len(Xs) / tm["solve_ms"] / 1e9, "t": tm}))
    if n == 4096:
        t0 = time.time(); mu0, var0 = orc.predict(spec, X, y, Xs, True); t1 = time.time()
        print(json.dumps({"n": n, "cpu_s": t1 - t0, "mean": relmax(mu, mu0), "var_rel": rel(var, var0)}))
eng.set_option("lookahead", 0)
eng.factorize(); print("no-lookahead", eng.timings())
