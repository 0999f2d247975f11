"""Multi-GPU check (run under torchrun, one rank per GPU): row-block-sharded Cholesky over NCCL vs the single-GPU factor,
sliced prediction + gather vs the oracle, and factorisation timings.  Used by tests/test_dist_gpu.py and by hand:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine  # noqa: E402
from gumbi_b200 import dist as gdist  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 3000, 8192]
    with_oracle = os.environ.get("GB2_DIST_ORACLE", "1") == "1"
    prec = os.environ.get("GB2_DIST_PRECISION", "fp64")
    eng = GPEngine(local_rank, prec)
    if os.environ.get("GB2_DIST_P2P", "1") == "0":
        eng.set_option("p2p", 0)     # ablation: NCCL broadcast + all-gather instead of NVLink peer stores
    shard = os.environ.get("GB2_DIST_SHARD", "0") == "1"
    if shard:
        eng.set_option("shard_storage", 1)   # every rank stores only its own row blocks; predict is a collective over all points
    r, w = gdist.init_engine(eng)
    assert (r, w) == (rank, world)
    ref = GPEngine(local_rank, prec)  # same GPU, not sharded
    ok = True
    out = []
    for n in sizes:
        P = 2 if n == 3000 else 1
        spec, X, y, Xs = synthetic_problem(n // P, 4, P=P, M_res=20, kind="Matern52" if n == 3000 else "ExpQuad")
        eng.set_train(X, y); eng.set_kernel(spec)
        ref.set_train(X, y); ref.set_kernel(spec)
        for _ in range(2):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter(); eng.factorize(); torch.cuda.synchronize(); t_shard = time.perf_counter() - t0
        tm = eng.timings()
        for _ in range(2):
            t0 = time.perf_counter(); ref.factorize(); t_one = time.perf_counter() - t0
        tm1 = ref.timings()
        rec = {"N": len(y), "world": world, "rank": rank, "precision": prec, "p2p": os.environ.get("GB2_DIST_P2P", "1"), "shard_storage": shard, "chol_ms_sharded": tm["cholesky_ms"], "chol_ms_single": tm1["cholesky_ms"],
               "wall_ms_sharded": t_shard * 1e3, "wall_ms_single": t_one * 1e3}
        if len(y) <= 8192:
            L, L1 = eng.get_L(), ref.get_L()
            if shard:   # each rank returns its own row blocks (zeros elsewhere): sum over the ranks
                Lt = torch.from_numpy(L).cuda()
                dist.all_reduce(Lt)
                L = Lt.cpu().numpy()
            rec["L_bit_identical"] = bool(np.array_equal(L, L1))
            rec["L_max_abs_diff"] = float(np.max(np.abs(L - L1)))
            rec["v_max_abs_diff"] = float(np.max(np.abs(eng.get_v() - ref.get_v())))
            rec["mll_diff"] = abs(eng.mll() - ref.mll())
            ok &= rec["L_max_abs_diff"] < 1e-9 and rec["v_max_abs_diff"] < 1e-9
        if shard:
            mu, var = eng.predict(Xs, True)
        else:
            lo, hi = gdist.grid_slice(len(Xs), rank, world)
            mu_l, var_l = eng.predict(Xs[lo:hi], True)
            mu, var = gdist.gather_grid(mu_l, var_l, len(Xs))
        mu1, var1 = ref.predict(Xs, True)
        rec["pred_max_abs_diff"] = float(max(np.max(np.abs(mu - mu1)), np.max(np.abs(var - var1))))
        # (tf32 + shard_storage: the sharded solve is fp64 while the single-GPU reference solve is split-TF32)
        ok &= rec["pred_max_abs_diff"] < (1e-8 if not (shard and prec == "tf32") else 2e-2)
        rec["pred_ms"] = eng.timings()["solve_ms"] + eng.timings()["kstar_ms"] + eng.timings()["reduce_ms"]
        if not shard and prec == "fp64" and len(y) <= 4096:
            # find_MAP's objective on a sharded factorisation: every rank holds the complete factor, so the gradient is computed
            # redundantly and must equal the single-GPU one
            val_s, g_s = eng.mll_grad(spec)
            val_1, g_1 = ref.mll_grad(spec)
            rec["mll_grad_max_abs_diff"] = float(max(abs(val_s - val_1), abs(g_s["sigma"] - g_1["sigma"]),
                                                     np.max(np.abs(np.asarray(g_s["terms"][0]["ls"]) - np.asarray(g_1["terms"][0]["ls"])))))
            ok &= rec["mll_grad_max_abs_diff"] < 1e-9 * max(1.0, abs(g_1["sigma"]))
        if with_oracle and rank == 0 and len(y) <= 4096:
            from oracle import gp_oracle as orc

            mu0, var0 = orc.predict(spec, X, y, Xs, True)
            rec["oracle_max_rel"] = float(max(np.max(np.abs(mu - mu0)) / np.abs(mu0).max(), np.max(np.abs(var - var0) / np.abs(var0))))
            ok &= rec["oracle_max_rel"] < (1e-6 if prec == "fp64" else 1e-2)
        out.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    # a non-positive-definite matrix: every rank must raise, and report the SAME pivot (the owner of the failing block knows it,
    # the verdict is agreed collectively inside gb2_factorize)
    spec, X, y, Xs = synthetic_problem(600, 2)
    spec["sigma"] = 0.0
    spec["jitter"] = 0.0
    X[450] = X[140]
    eng.set_train(X, y); eng.set_kernel(spec)
    try:
        eng.factorize()
        msg = "no error"
    except np.linalg.LinAlgError as e:
        msg = str(e)
    msgs = [None] * world
    dist.all_gather_object(msgs, msg)
    same = all(m == msgs[0] for m in msgs) and "not positive definite" in msgs[0]
    if rank == 0:
        print(json.dumps({"non_pd_verdict_identical_on_all_ranks": same, "message": msgs[0]}), flush=True)
    ok &= same
    # a full distributed find_MAP through the plugin class: every rank drives the same L-BFGS-B over the collective objective and must
    # end at the SAME point after the SAME number of evaluations (rank 0's value / gradient are broadcast, gumbi_b200/map.py)
    eng.close(); ref.close()
    if prec == "fp64" and not shard and os.environ.get("GB2_DIST_P2P", "1") == "1":
        from gumbi_b200 import ArrayGP

        spec, X, y, Xs = synthetic_problem(700, 2)
        gp = ArrayGP(X, y, ["x0", "x1"], device=local_rank, distributed=True)
        gp.build_model()
        MAP = gp.find_MAP(options={"maxiter": 12})
        key = (tuple(np.round(np.asarray(MAP["ls_total"], dtype=np.float64), 14).tolist()), round(float(MAP["σ"]), 14), int(gp.map_evals))
        keys = [None] * world
        dist.all_gather_object(keys, key)
        same_map = all(k == keys[0] for k in keys)
        mu_m, _ = gp.predict(Xs[:64], with_noise=True)
        if rank == 0:
            print(json.dumps({"find_MAP_identical_on_all_ranks": same_map, "n_eval": keys[0][2], "ls": keys[0][0]}), flush=True)
        ok &= same_map and bool(np.all(np.isfinite(mu_m)))
        gp.engine.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
