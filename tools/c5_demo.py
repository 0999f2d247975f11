"""BASELINE config 5 on the storage-sharded path (run under torchrun, one rank per GPU):
4-output LCM (Q = 2 terms), n = 65536 locations -> stacked N = 262144, d = 16, fp64, 4 x 10k-point grid.
The lower triangle of K is 275 GB: it only exists distributed (69 GB of row blocks per GPU on 8 GPUs).
Two cold passes (K-build + block Cholesky with NVLink panel pushes + distributed solve); the second is reported.
No CPU / single-GPU comparison is possible at this size; bit-identity of the sharded factor with the single-GPU one is
established at N <= 16384 by tests/dist_check.py.  Reports size-independent checks instead."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine
from gumbi_b200 import dist as gdist
from gumbi_b200.synthetic import synthetic_problem

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
precision = sys.argv[3] if len(sys.argv) > 3 else "fp64"   # "tf32": split-TF32 tcgen05 trailing updates (the sharded solve stays fp64)
spec, X, y, Xs = synthetic_problem(n, 16, P=4, M_res=100, kind="ExpQuad", Q=2)
N, M = len(y), len(Xs)
eng = GPEngine(local_rank, precision)
eng.set_option("shard_storage", 1)
gdist.init_engine(eng)
eng.set_train(X, y)
eng.set_kernel(spec)
out = {}
for p in range(passes):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.factorize()
    t1 = time.perf_counter()
    mu, var = eng.predict(Xs, True)
    t2 = time.perf_counter()
    tm = eng.timings()
    free, total = torch.cuda.mem_get_info()
    out = {"config": "c5: 4-output LCM (Q=2), n=%d, stacked N=%d, d=16, M=%d, %s, %d GPUs, storage-sharded" % (n, N, M, precision, world),
           "pass": p, "factorize_s": t1 - t0, "predict_s": t2 - t1, "cold_step_s": t2 - t0, "predictions_per_s": M / (t2 - t0),
           "phases_ms": {k: v for k, v in tm.items() if k.endswith("_ms")},
           "cholesky_tflops_aggregate": N ** 3 / 3 / (tm["cholesky_ms"] * 1e-3) / 1e12,
           "solve_tflops_aggregate": float(N) * N * M / (tm["solve_ms"] * 1e-3) / 1e12,
           "hbm_used_gb_this_rank": (total - free) / 1e9, "lower_triangle_gb": 8.0 * N * (N + 1) / 2 / 1e9, "mll": eng.mll(),
           "checks": {"all_finite": bool(np.all(np.isfinite(mu)) and np.all(np.isfinite(var))),
                      "var_min": float(var.min()), "var_max": float(var.max()), "sigma2": spec["sigma"] ** 2,
                      "mean_abs_max": float(np.abs(mu).max())}}
    if rank == 0:
        print(json.dumps(out), flush=True)
# every rank computed the same full result
h = torch.tensor([float(np.sum(mu)), float(np.sum(var))], dtype=torch.float64, device="cuda")
hs = [torch.zeros_like(h) for _ in range(world)]
dist.all_gather(hs, h)
if rank == 0:
    print("RESULT_IDENTICAL_ON_ALL_RANKS", all(torch.equal(hs[0], t) for t in hs), flush=True)
eng.close()
dist.destroy_process_group()
