"""Cold predict, two ways, on BASELINE config 2 (N=8192, d=8, 10k-point grid, fp64, host buffers):
    (a) gb2_factorize + gb2_predict             -- recursive triangular solve after the factorisation
    (b) gb2_factorize_predict                   -- prediction points carried through the factorisation as extra rows
plus the solve_streams variants of (a).  Prints one JSON line per variant.   python tools/fused_timing.py [n] [d] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
spec, X, y, Xs = synthetic_problem(n, d)
e = GPEngine()
e.set_train(X, y)


def two_calls():
    e.set_kernel(spec)
    e.factorize()
    return e.predict(Xs, True)


def one_call():
    e.set_kernel(spec)
    return e.factorize_predict(Xs, True)


ref = None
for name, fn, streams, group in (("factorize+predict", two_calls, 1, 4), ("factorize+predict solve_streams=2", two_calls, 2, 4),
                                 ("factorize+predict solve_streams=4", two_calls, 4, 4), ("factorize_predict fused_group=1", one_call, 1, 1),
                                 ("factorize_predict fused_group=2", one_call, 1, 2), ("factorize_predict fused_group=4", one_call, 1, 4),
                                 ("factorize_predict fused_group=8", one_call, 1, 8)):
    e.set_option("solve_streams", streams)
    e.set_option("fused_group", group)
    for _ in range(3):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    tm = e.timings()
    if ref is None:
        ref = out
    print(json.dumps({"variant": name, "N": n, "d": d, "M": len(Xs), "ms_per_step": ms, "predictions_per_s": len(Xs) / (ms * 1e-3),
                      "phases_ms": {k: round(v, 3) for k, v in tm.items() if k.endswith("_ms")},
                      "max_rel_dev_mean": float(np.max(np.abs(out[0] - ref[0]) / (1e-12 + np.abs(ref[0])))),
                      "max_rel_dev_var": float(np.max(np.abs(out[1] - ref[1]) / (1e-12 + np.abs(ref[1]))))}), flush=True)
e.close()
