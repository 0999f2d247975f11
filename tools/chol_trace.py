"""Per-block-step timeline of the blocked Cholesky (set_option("trace", 1) / gb2_get_trace): does the look-ahead really overlap the
chain  diagonal kernel -> panel solve -> next-column update  with the bulk trailing update of the previous step?

Hypothesis to test: the diagonal-panel kernel needs 222 KB of shared memory, i.e. an EMPTY SM, while the bulk update
keeps two 92 KB CTAs on every SM; a freed 92 KB slot is refilled by the next bulk CTA, so the diagonal kernel of step k starts only
when the bulk update of step k-1 has drained -- the factorisation then costs sum(bulk) + steps x (chain), which is what C2 measures
(9.3 ms ~ 6.0 + 64 x 0.05).  Evidence: the stamp before the diagonal kernel (eligible) against the stamp after it (done) minus its
stand-alone duration = time spent waiting for an SM.

    python tools/chol_trace.py [n] [d]          # one GPU; prints a summary JSON line per configuration and the first steps
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
spec, X, y, Xs = synthetic_problem(n, d)
e = GPEngine()
e.set_train(X, y)
e.set_kernel(spec)
for name, opts, fused in (("look-ahead (default)", {}, False), ("look-ahead off", {"lookahead": 0}, False),
                          ("fused cold predict, group 4", {}, True), ("fused cold predict, group 1", {"fused_group": 1}, True),
                          ("two-level blocking, panel of 4", {"fp64_panel": 4}, False)):
    try:
        for k, v in {"lookahead": 1, "fused_group": 4, "fp64_panel": 0, **opts}.items():
            e.set_option(k, v)
    except RuntimeError as err:
        print(json.dumps({"config": name, "unavailable": str(err)}), flush=True)
        continue
    e.set_option("trace", 0)
    for _ in range(2):
        e.factorize_predict(Xs, True) if fused else e.factorize()
    e.set_option("trace", 1)
    e.factorize_predict(Xs, True) if fused else e.factorize()
    t = e.get_trace().astype(np.int64)
    tm = e.timings()
    steps = len(t)
    t0 = t[0, 0]
    us = (t - t0) / 1e3
    us[t == 0] = np.nan
    diag = us[:, 1] - us[:, 0]                       # eligible -> done: stand-alone duration + time waiting for an SM
    diag_alone = np.nanmin(diag)
    wait = diag - diag_alone
    panel = us[:, 2] - us[:, 1]
    nextcol = us[:, 3] - us[:, 2]
    bulk = us[:, 5] - us[:, 4]
    # overlap: how long before the previous step's bulk update finished did this step's diagonal kernel finish (positive = overlapped)
    lead = np.r_[np.nan, us[:-1, 5] - us[1:, 1]]
    out = {"config": name, "N": n, "steps": steps, "cholesky_ms": tm["cholesky_ms"], "span_ms": float(np.nanmax(us) / 1e3),
           "diag_alone_us": float(diag_alone), "diag_wait_total_ms": float(np.nansum(wait) / 1e3), "diag_wait_median_us": float(np.nanmedian(wait)),
           "panel_solve_total_ms": float(np.nansum(panel) / 1e3), "next_column_total_ms": float(np.nansum(nextcol) / 1e3),
           "bulk_total_ms": float(np.nansum(bulk) / 1e3), "steps_with_diag_done_before_prev_bulk_done": int(np.nansum(lead > 0)),
           "median_lead_us": float(np.nanmedian(lead))}
    print(json.dumps(out), flush=True)
    if name.startswith("look-ahead (default)"):
        print("step  diag_elig  diag_done  panel_done  nextcol_done  bulk_elig  bulk_done   (us since the first stamp)")
        for k in range(min(8, steps)):
            print("%4d " % k + " ".join("%10.1f" % v for v in us[k]))
e.close()
