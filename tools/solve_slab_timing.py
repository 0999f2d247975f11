"""Dev tool: warm predict time at N = 32768 for small prediction batches (what one rank of an 8- or 4-GPU run serves) as a function of
solve_streams -- decides the slab rule of predict_common.   python tools/solve_slab_timing.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402

spec, X, y, Xs = synthetic_problem(32768, 8, M_res=100, kind="Matern52")
e = GPEngine()
e.set_train(X, y)
e.set_kernel(spec)
e.factorize()
for M in (1250, 2500, 5000, 10000):
    row = {"M": M}
    for streams in (1, 2, 4):
        e.set_option("solve_streams", streams)
        e.predict(Xs[:M], True)
        t0 = time.perf_counter()
        for _ in range(3):
            e.predict(Xs[:M], True)
        row[f"streams={streams}"] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    print(json.dumps(row), flush=True)
e.close()
