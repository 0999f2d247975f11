"""BASELINE config 5 as an ICM (one Coregion term, Q = 1) through the Kronecker-aware multi-output solve (gumbi_b200/kron.py):
4 outputs, n = 65536 shared locations -> stacked N = 262144, d = 16, fp64, 4 x 10k-point grid.  The stacked system (275 GB lower
triangle) is never formed: it is rotated into 4 independent 65536 x 65536 problems (34 GB each), one per GPU, with NO exchange on
the data path.  Run under torchrun with 1, 2 or 4 ranks (blocks are dealt round-robin):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 tools/c5_kron_demo.py [n] [passes]

At n <= 4096 the result is also compared with the dense stacked solve on rank 0's GPU."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import ArrayGP  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
P, d = 4, 16
spec, X, y, Xs = synthetic_problem(n, d, P=P, M_res=100, kind="ExpQuad", Q=1)
t = spec["terms"][0]
point = {"ls_total": t["ls"], "η_total": t["eta"], "σ": spec["sigma"], "W_Variable": t["coreg"][0]["W"], "κ_Variable": t["coreg"][0]["kappa"],
         "W_Output_noise": spec["noise_coreg"]["W"], "κ_Output_noise": spec["noise_coreg"]["kappa"]}
kw = dict(categorical_dims=["Variable"], categorical_levels={"Variable": [f"y{p}" for p in range(P)]}, outputs=[f"y{p}" for p in range(P)])
gp = ArrayGP(X, y, [f"x{j}" for j in range(d)], device=local_rank, distributed=world > 1, multioutput="kron", **kw)
gp.build_model()
gp.find_MAP(point=point)
N, M = len(y), len(Xs)
for p in range(passes):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mu, var = gp.predict_cold(Xs)
    t1 = time.perf_counter()
    free, total = torch.cuda.mem_get_info()
    tm = {q: e.timings() for q, e in gp.engine.blocks.items()}
    out = {"config": "c5 as ICM (Q=1) via the Kronecker solve: 4 outputs, n=%d (stacked N=%d), d=16, M=%d, fp64, %d GPUs" % (n, N, M, world),
           "pass": p, "cold_step_s": t1 - t0, "predictions_per_s": M / (t1 - t0), "blocks_on_this_rank": sorted(tm),
           "cholesky_ms_per_block": {q: v["cholesky_ms"] for q, v in tm.items()},
           "block_cholesky_tflops": {q: n ** 3 / 3 / (v["cholesky_ms"] * 1e-3) / 1e12 for q, v in tm.items()},
           "equivalent_dense_cholesky_flop": N ** 3 / 3, "block_cholesky_flop_total": P * n ** 3 / 3,
           "hbm_used_gb_this_rank": (total - free) / 1e9, "mll": gp.marginal_log_likelihood(),
           "checks": {"all_finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var))), "var_min": float(var.min()),
                      "var_max": float(var.max()), "mean_abs_max": float(np.abs(mu).max())}}
    if rank == 0:
        print(json.dumps(out), flush=True)
if n <= 4096 and rank == 0:
    dense = ArrayGP(X, y, [f"x{j}" for j in range(d)], device=local_rank, **kw)
    dense.build_model()
    dense.find_MAP(point=point)
    mu_d, var_d = dense.predict(Xs)
    print(json.dumps({"max_rel_dev_mean_vs_dense": float(np.max(np.abs(mu - mu_d) / (1e-12 + np.abs(mu_d)))),
                      "max_rel_dev_var_vs_dense": float(np.max(np.abs(var - var_d) / (1e-12 + np.abs(var_d))))}), flush=True)
    dense.engine.close()
gp.engine.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
