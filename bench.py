#!/usr/bin/env python
"""bench.py -- posterior predictions/sec on an M-point grid (BASELINE.json metric), B200 CUDA core vs the CPU path.

A "step" is ONE reference-equivalent ``predict`` call over the M-point grid: the reference rebuilds K(X,X), re-factorises
and solves on every call (gumbi/regression/pymc/GP.py:843-847), so a step here is K-build -> jittered
Cholesky (+ v = L^-1 y) -> K(X*,X) build -> triangular solve -> posterior mean/variance.  Nothing is cached between
steps ("cold").  The warm figure (factor resident, what this backend does for repeated predicts) is reported beside it
under "warm".

  value : steps timed with CUDA events on the handle's stream, training inputs + grid already resident in HBM.
  e2e   : the same step through the plugin class (ArrayGP.build_model -> find_MAP(point=) -> predict) with HOST numpy
          buffers: H2D of X, y and the grid and D2H of mean/var happen inside the timed region every step.
  roofline      : the dominant kernel, dgemm_nt_kernel (DMMA fp64), measured on the predict triangular solve which
                  consists of that kernel only: N^2 M' flop / solve_ms.
  roofline_cholesky / roofline_kbuild: N^3/3 flop over the whole factorisation; lower-triangle bytes over the K-build.
  cpu_baseline  : oracle/gp_oracle.py (numpy/scipy restatement of the PyMC path) on the host cores, same workload.

Workloads (BASELINE.json configs): c2 (default; N=8192 d=8 ExpQuad, M=10^4, fp64 -- the config the metric is quoted on),
c1 (N=392 d=1, M=200), c3 (2-output ICM, n=16384 -> N=32768, d=4), c4 (Matern52 N=32768 d=8).
N>1 GPUs: ONE problem, strong scaling -- the factorisation is row-block-cyclic sharded over the ranks (per block step an NCCL
broadcast of the diagonal block and an all-gather of the panel), every rank then holds the factor and serves 1/N of the grid.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n, d, P, kind, M_res, Q, description)
    "c1": (392, 1, 1, "ExpQuad", 200, 1, "single-output RBF, N=392 d=1, 200-pt grid (synthetic stand-in for mpg~horsepower)"),
    "c2": (8192, 8, 1, "ExpQuad", 100, 1, "single-output ARD RBF, synthetic N=8192 d=8, 10k-pt grid, fp64"),
    "c3": (16384, 4, 2, "ExpQuad", 100, 1, "2-output ICM coregion, n=16384 (stacked N=32768) d=4, 2x10k-pt grid, fp64"),
    "c4": (32768, 8, 1, "Matern52", 100, 1, "Matern-5/2 ARD, N=32768 d=8, 10k-pt grid"),
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    """Roofline denominators: MEASURED_PEAKS.json (driver-written) when present, else the B200_PROFILING.md fallback; says which."""
    fb = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
        except Exception:
            d = {}

        def find(keys):
            for k in keys:
                v = d.get(k)
                if isinstance(v, dict):
                    v = v.get("value", v.get("burst"))
                if isinstance(v, (int, float)) and v > 0:
                    return float(v)
            return None

        hbm = find(["hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"])
        bf16 = find(["bf16_tflops", "bf16_tflops_burst", "bf16_tf"])
        if hbm and bf16:
            return {"hbm_gbs": hbm, "bf16_tflops": bf16, "source": "MEASURED_PEAKS.json (measured)"}
        return {"hbm_gbs": hbm or fb["hbm_gbs"], "bf16_tflops": bf16 or fb["bf16_tflops"],
                "source": "MEASURED_PEAKS.json where it has the entry, B200_PROFILING.md fallback otherwise"}
    return {**fb, "source": "B200_PROFILING.md fallback"}


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu pass (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


def cublas_dgemm_peak(torch, dev):
    """fp64 tensor roofline denominator: MEASURED_PEAKS.json carries no fp64 figure, so cuBLAS DGEMM 8192^3 is measured
    live (burst, best of 5) outside the timed region.  tcgen05 has no fp64 kind; DMMA is the fp64 tensor path."""
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best / 1e9


def live_copy_bandwidth(torch, dev):
    """STREAM-style device copy (read + write bytes / time), 2 GiB, best of 5 -- context for the HBM roofline when
    MEASURED_PEAKS.json is absent; never replaces the contract's denominator."""
    try:
        n = (1 << 31) // 8
        a = torch.empty(n, dtype=torch.float64, device=dev).normal_()
        b = torch.empty_like(a)
        b.copy_(a)
        torch.cuda.synchronize(dev)
        best = 1e30
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n * 8 / (best * 1e-3) / 1e9
    except Exception:
        return None


def make_workload(name):
    from gumbi_b200.synthetic import synthetic_problem

    n, d, P, kind, M_res, Q, desc = WORKLOADS[name]
    spec, X, y, Xs = synthetic_problem(n, d, P=P, M_res=M_res, kind=kind, Q=Q)
    return spec, X, y, Xs, desc


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU path (numpy/scipy restatement of PyMC's Marginal.predict) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from threadpoolctl import threadpool_info

    from oracle import gp_oracle as orc

    spec, X, y, Xs, desc = make_workload(args.workload)
    M, N = len(Xs), len(y)
    cores = os.cpu_count()
    t0 = time.perf_counter()
    orc.predict(spec, X, y, Xs, True)  # warm-up 1 (also calibrates)
    t_full = time.perf_counter() - t0
    warm = max(0, args.warmup - 1)
    budget = 240.0
    steps = args.steps
    if (steps + warm) * t_full <= budget:
        for _ in range(warm):
            orc.predict(spec, X, y, Xs, True)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.predict(spec, X, y, Xs, True)
        dt = (time.perf_counter() - t0) / steps
        sample = f"{steps} full cold predict calls, N={N}, grid M={M}"
    else:
        # Bounded sample: every step still pays the full per-call factorisation (K-build + dpotrf + v), but solves only
        # the first Ms grid points; the solve is linear in M, so the step time for the full grid is
        # t_fact + t_solve(Ms) * M / Ms.  If even the factorisations do not fit the budget, fewer of them are timed.
        Ms = max(64, M // 16)
        n_fact = int(max(1, min(steps, budget / max(t_full * 0.5, 1e-9))))
        t_fact = t_solve = 0.0
        for _ in range(n_fact):
            t0 = time.perf_counter()
            L, v = orc.factorize(spec, X, y)
            t1 = time.perf_counter()
            orc.conditional(spec, X, L, v, Xs[:Ms], True)
            t2 = time.perf_counter()
            t_fact += t1 - t0
            t_solve += t2 - t1
        dt = t_fact / n_fact + (t_solve / n_fact) * M / Ms
        sample = (f"{n_fact} cold calls with the full factorisation (N={N}) and the first {Ms} of {M} grid points each; "
                  f"step time = t_fact + t_solve*M/Ms (solve is linear in M)")
    val = M / dt
    blas = [f"{i.get('internal_api')}:{i.get('num_threads')}" for i in threadpool_info()]
    line = {
        "impl": "reference", "metric": "posterior predictions/sec on M-point grid (cold: K-build + Cholesky + solve per call)",
        "value": val, "unit": "predictions/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "N": N, "M": M, "d": X.shape[1]},
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}; numpy/scipy restatement of the PyMC path "
                                   f"(PyMC itself is not installable here), BLAS threads {blas}"},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cholesky_gflops": N ** 3 / 3 / max(1e-9, _time_potrf(orc, spec, X, y)) / 1e9,
    }
    print(json.dumps(line), flush=True)


def _time_potrf(orc, spec, X, y):
    import scipy.linalg as sla

    K = orc.train_cov(spec, X)
    t0 = time.perf_counter()
    sla.cholesky(K, lower=True, check_finite=False, overwrite_a=True)
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def ensure_built():
    """The CUDA library is built in-tree by __graft_entry__.build(); build it here only if it is missing (nvcc, sm_100a)."""
    from gumbi_b200 import _lib

    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__ as ge

        if int(os.environ.get("LOCAL_RANK", 0)) == 0:
            ge.build()
        else:   # another rank of the same node is compiling: wait for the file
            for _ in range(600):
                if os.path.exists(_lib.lib_path()):
                    break
                time.sleep(0.5)


def run_ours(args):
    import torch

    ensure_built()

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the gumbi_b200 core has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from gumbi_b200 import ArrayGP, GPEngine
    from gumbi_b200 import dist as gdist

    spec, X, y, Xs, desc = make_workload(args.workload)
    N, D_in, M = len(y), X.shape[1], len(Xs)
    precision = args.precision
    warmup = max(3, args.warmup)

    # ---- device-resident arm ("value") --------------------------------------------------------------------------
    # N GPUs: ONE problem, strong scaling -- the factorisation is row-block sharded over the ranks (NCCL broadcast of the
    # diagonal block + all-gather of the panel every block step), every rank then holds the factor and serves a contiguous
    # 1/N slice of the grid; the slices are all-gathered on the handle's stream inside the timed region.
    eng = GPEngine(local_rank, precision)
    for kv in args.opt:
        eng.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if use_dist:
        gdist.init_engine(eng)
    shard = use_dist and any(o.replace(" ", "") == "shard_storage=1" for o in args.opt)
    # storage-sharded mode: the factor lives distributed over the ranks, prediction is a collective over the whole grid
    lo, hi = (0, M) if shard else gdist.grid_slice(M, rank, world)
    slot = M if shard else -(-M // world)   # padded slice length (equal counts for the all-gather)
    dX = torch.from_numpy(X).to(dev)
    dy = torch.from_numpy(y).to(dev)
    dXs = torch.from_numpy(np.ascontiguousarray(Xs[lo:hi])).to(dev)
    dloc = torch.zeros(2 * slot, dtype=torch.float64, device=dev)          # [mean slice | var slice]
    dall = torch.zeros(2 * slot * world, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(dev)
    eng.set_train_device(dX.data_ptr(), N, D_in, dy.data_ptr())

    def step_dev():
        eng.set_kernel(spec)       # hyper-parameters arrive per call (point=MAP); tiny
        eng.factorize()            # K-build + Cholesky + v  (collective when world > 1)
        eng.predict_device(dXs.data_ptr(), hi - lo, True, dloc.data_ptr(), dloc.data_ptr() + 8 * slot)
        if use_dist and not shard:
            eng.allgather_device(dloc.data_ptr(), dall.data_ptr(), 2 * slot)

    def gathered():
        if not use_dist or shard:
            return dloc[:M].cpu().numpy(), dloc[slot:slot + M].cpu().numpy()
        a = dall.cpu().numpy().reshape(world, 2, slot)
        mu = np.concatenate([a[r, 0, : gdist.grid_slice(M, r, world)[1] - gdist.grid_slice(M, r, world)[0]] for r in range(world)])
        var = np.concatenate([a[r, 1, : gdist.grid_slice(M, r, world)[1] - gdist.grid_slice(M, r, world)[0]] for r in range(world)])
        return mu, var

    for _ in range(warmup):
        step_dev()
    phase = {k: 0.0 for k in ("prep_ms", "kbuild_ms", "cholesky_ms", "kstar_ms", "solve_ms", "reduce_ms")}
    launches = 0
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    eng.mark(0)
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        step_dev()
        tm = eng.timings()
        for k in phase:
            phase[k] += tm[k]
        launches += int(tm["launches_factorize"] + tm["launches_predict"])
    eng.mark(1)
    ms_total = eng.elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = M / (ms_step * 1e-3)
    for k in phase:
        phase[k] = max_over_ranks(phase[k] / args.steps)

    # warm predicts (factor resident)
    eng.mark(2)
    for _ in range(args.steps):
        eng.predict_device(dXs.data_ptr(), hi - lo, True, dloc.data_ptr(), dloc.data_ptr() + 8 * slot)
        if use_dist and not shard:
            eng.allgather_device(dloc.data_ptr(), dall.data_ptr(), 2 * slot)
    eng.mark(3)
    warm_ms = max_over_ranks(eng.elapsed_ms(2, 3) / args.steps)
    mu_dev, var_dev = gathered()

    # ---- end-to-end arm through the plugin class, host buffers ---------------------------------------------------
    n_, d_, P_, kind_, _, Q_, _ = WORKLOADS[args.workload]
    cat = {}
    if P_ > 1:
        cat = dict(categorical_dims=["Variable"], categorical_levels={"Variable": [f"y{p}" for p in range(P_)]},
                   outputs=[f"y{p}" for p in range(P_)])
    cont = [f"x{j}" for j in range(d_)]
    point = {"ls_total": spec["terms"][0]["ls"], "η_total": spec["terms"][0]["eta"], "σ": spec["sigma"]}
    if P_ > 1:
        point["W_Variable"] = spec["terms"][0]["coreg"][0]["W"]
        point["κ_Variable"] = spec["terms"][0]["coreg"][0]["kappa"]
        point["W_Output_noise"] = spec["noise_coreg"]["W"]
        point["κ_Output_noise"] = spec["noise_coreg"]["kappa"]
    e2e = None
    if Q_ == 1:
        eng.close()  # free the first handle's factor before the plugin allocates its own
        del eng
        gp = ArrayGP(X, y, cont, device=local_rank, precision=precision, distributed=use_dist, **cat)
        e2e_opts = list(args.opt)

        host_ms = {"build_model": 0.0, "find_MAP": 0.0, "predict": 0.0}

        def step_e2e():
            t0 = time.perf_counter()
            gp.build_model(continuous_kernel=kind_)   # H2D X, y
            while e2e_opts:
                kv = e2e_opts.pop()
                gp.engine.set_option(kv.split("=")[0], int(kv.split("=")[1]))
            t1 = time.perf_counter()
            gp.find_MAP(point=point)
            t2 = time.perf_counter()
            out = gp.predict(Xs, with_noise=True)       # K-build + Cholesky + solve; H2D grid (slice), D2H mean/var (+ gather)
            t3 = time.perf_counter()
            host_ms["build_model"] += (t1 - t0) * 1e3; host_ms["find_MAP"] += (t2 - t1) * 1e3; host_ms["predict"] += (t3 - t2) * 1e3
            return out

        for _ in range(3):
            mu_h, var_h = step_e2e()
        barrier()
        for k in host_ms:
            host_ms[k] = 0.0
        t0 = time.perf_counter()
        for _ in range(args.steps):
            mu_h, var_h = step_e2e()
        torch.cuda.synchronize(dev)
        dt = max_over_ranks((time.perf_counter() - t0) / args.steps)
        e2e = {"value": M / dt, "unit": "predictions/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(X.nbytes + y.nbytes + Xs[lo:hi].nbytes), "d2h_bytes_per_step": int(16 * (hi - lo)),
               "timing": "host wall clock around the public call (it returns host arrays), max over ranks"}
        tol = 1e-6 if precision == "fp64" else 1e-2
        assert np.max(np.abs(mu_h - mu_dev)) <= tol * np.max(np.abs(mu_dev)), "e2e and device arms disagree"
        e2e["device_phases_ms_last_step"] = {k: v for k, v in gp.engine.timings().items() if k.endswith("_ms")}
        e2e["host_call_ms_per_step"] = {k: v / args.steps for k, v in host_ms.items()}
        gp.engine.close()

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    # ---- rooflines ---------------------------------------------------------------------------------------------------
    peaks = measured_peaks()
    dgemm_peak = cublas_dgemm_peak(torch, dev)
    copy_gbs = live_copy_bandwidth(torch, dev)
    Ml = hi - lo
    kb_bytes = 8.0 * N * (N + 1) / 2 + 8.0 * N * D_in
    kb_gbs = kb_bytes / world / (phase["kbuild_ms"] * 1e-3) / 1e9
    chol_tflops = N ** 3 / 3 / (phase["cholesky_ms"] * 1e-3) / 1e12
    solve_tflops = float(N) * N * Ml / (phase["solve_ms"] * 1e-3) / 1e12
    nblk = (N + 1 + 127) // 128
    if precision == "fp64":
        solve_launches = 2 * nblk - 1
        roofline = {
            "kernel": "dgemm_nt_kernel (DMMA m8n8k4 fp64) in the predict triangular solve L^-1 K(X,X*)",
            "bound": "tensor", "achieved": solve_tflops, "peak": dgemm_peak, "unit": "TFLOP/s", "frac": solve_tflops / dgemm_peak,
            "peak_source": "cuBLAS DGEMM 8192^3 burst measured live in this run (MEASURED_PEAKS.json has no fp64 entry; tcgen05 has no fp64 kind)",
            "algorithmic_flop_per_step": float(N) * N * Ml, "launches_per_step": solve_launches,
            "avg_launch_ms": phase["solve_ms"] / solve_launches,
            "traffic": (measured_traffic(f"{args.workload}:fp64:solve") or {}).get("dram_bytes_per_launch") if world == 1 else None,
            "traffic_source": (measured_traffic(f"{args.workload}:fp64:solve") or {}).get("source"),
        }
    else:
        tf32_peak = peaks["bf16_tflops"] / 2.0
        roofline = {
            "kernel": "gemm_tf32x3_kernel (tcgen05.mma kind::tf32, 3 MMAs per product) in the predict triangular solve L^-1 K(X,X*)",
            "bound": "tensor", "achieved": 3.0 * solve_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": 3.0 * solve_tflops / tf32_peak,
            "peak_source": f"half of the bf16 GEMM peak of {peaks['source']} (tf32 dense rate = 1/2 bf16)",
            "fp64_equivalent_tflops": solve_tflops, "algorithmic_flop_per_step": 3.0 * float(N) * N * Ml,
            "note": "achieved counts the 3 tf32 MMAs issued per fp64-equivalent product; the fp64 leaf sub-solves (512 columns) are inside the timed phase",
            "traffic": None,
        }
    roofline_chol = {"kernel": "blocked Cholesky (potrf_diag + DMMA panel" + (" + tcgen05 split-TF32 trailing SYRK)" if precision != "fp64" else "/trailing update)"),
                     "bound": "tensor", "achieved": chol_tflops, "peak": dgemm_peak, "unit": "TFLOP/s (fp64-equivalent N^3/3)",
                     "frac": chol_tflops / dgemm_peak, "algorithmic_flop_per_step": N ** 3 / 3,
                     "peak_source": "cuBLAS DGEMM 8192^3 burst measured live in this run"}
    roofline_kb = {"kernel": "kbuild_strip_kernel<TRAIN> (single stationary term) / kbuild_dmma_kernel (Linear, Coregion, additive models)", "bound": "hbm", "achieved": kb_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": kb_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"], "algorithmic_bytes_per_launch": kb_bytes / world, "traffic": None,
                   "device_copy_gbs_measured_live": copy_gbs, "frac_of_live_copy": (kb_gbs / copy_gbs) if copy_gbs else None,
                   "note": "fp64-pipe bound on B200 (DESIGN.md section 3): DMMA Gram + table-driven exp need ~24 fp64 issue slots per entry"}

    # ---- CPU baseline on the host cores (bounded: one full cold call) ---------------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:
        from oracle import gp_oracle as orc

        t0 = time.perf_counter()
        mu_c, var_c = orc.predict(spec, X, y, Xs, True)
        t_cpu = time.perf_counter() - t0
        cpu = {"value": M / t_cpu, "unit": "predictions/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"1 full cold predict call (N={N}, M={M}) of the numpy/scipy restatement of the PyMC path, {t_cpu:.2f} s",
               "max_rel_err_mean_vs_gpu": float(np.max(np.abs(mu_dev - mu_c)) / np.max(np.abs(mu_c))),
               "max_rel_err_var_vs_gpu": float(np.max(np.abs(var_dev - var_c) / np.abs(var_c)))}

    line = {
        "metric": "posterior predictions/sec on M-point grid (cold: K-build + Cholesky + solve per call)",
        "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
        "dtype": "f64" if precision == "fp64" else "tf32x3+f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "N": N, "M": M, "d": d_, "outputs": P_, "kernel": kind_,
                   "parallelism": "single GPU" if world == 1 else (
                       f"factor stored row-block-sharded over {world} GPUs (NVLink peer pushes into panel rings), distributed column-sharded solve" if shard else
                       f"row-block-cyclic sharded Cholesky over {world} GPUs (NVLink peer exchange per block step), factor replicated, grid split {world} ways"),
                   "l2_policy": f"inputs larger than L2: the factor is {8.0 * N * N / 1e6:.0f} MB and is rewritten every step",
                   "engine_options": list(args.opt)},
        "phases_ms": phase, "wall_ms_per_step": wall_ms / args.steps,
        "warm": {"value": M / (warm_ms * 1e-3), "unit": "predictions/s", "ms_per_step": warm_ms},
        "cholesky_tflops": chol_tflops,
        "roofline": roofline, "roofline_cholesky": roofline_chol, "roofline_kbuild": roofline_kb,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp64", choices=["fp64", "tf32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="engine tunable name=value (e.g. tf32_nb=8, kbuild_v1=1); ablations only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
