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
  roofline      : the dominant kernel, dgemm_tma_kernel (TMA-staged DMMA fp64), measured on the predict triangular solve which
                  consists of that kernel only: N^2 M' flop / solve_ms.
  roofline_cholesky / roofline_kbuild: N^3/3 flop over the whole factorisation (per GPU); lower-triangle bytes over the K-build.
  cpu_baseline  : oracle/gp_oracle.py (numpy/scipy restatement of the PyMC path) on the host cores, bounded sample.
  also          : short runs of the other single-GPU configurations in the same process (C2 fp64, C4 split-TF32), N=1 only.

Workloads (BASELINE.json configs): c4 (default; N=32768 d=8 Matern-5/2, M=10^4 -- the size north_star quotes its targets on;
fp64 unless --precision tf32), c2 (N=8192 d=8 ExpQuad, M=10^4), c1 (N=392 d=1, M=200), c3 (2-output ICM, n=16384 -> N=32768, d=4).
N>1 GPUs: ONE problem, strong scaling -- the factorisation is row-block-cyclic sharded over the ranks (per block step the
diagonal block and the panel travel over NVLink), every rank then holds the factor and serves 1/N of the grid.
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv and os.environ.get("GB2_BENCH_REEXEC") != "1":
    # The CPU arm must use every host core.  torchrun exports OMP_NUM_THREADS=1 to its workers, which pins OpenBLAS to one
    # thread at import time (round 1: the N>1 reference runs were 4x slower than the N=1 run for this reason).  Re-execute
    # once with the BLAS thread variables set to the core count BEFORE numpy/scipy are imported.
    env = dict(os.environ)
    n = str(os.cpu_count() or 1)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env[k] = n
    env["GB2_BENCH_REEXEC"] = "1"
    os.execve(sys.executable, [sys.executable] + sys.argv, env)

import argparse
import json
import subprocess
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
# shared: the config dict both arms print (identical by construction)
# ------------------------------------------------------------------------------------------------------------------
METRIC = "posterior predictions/sec on M-point grid (cold: K-build + Cholesky + solve per call)"


def config_dict(args, world, N, M, D_in):
    n_, d_, P_, kind_, _, Q_, desc = WORKLOADS[args.workload]
    shard = world > 1 and any(o.replace(" ", "") == "shard_storage=1" for o in args.opt)
    return {"workload": f"{args.workload}: {desc}", "precision": args.precision, "N": N, "M": M, "d": d_, "outputs": P_, "kernel": kind_,
            "n_gpus": world,
            "parallelism": "single GPU" if world == 1 else (
                f"factor stored row-block-sharded over {world} GPUs (NVLink peer pushes into panel rings), distributed column-sharded solve" if shard else
                f"row-block-cyclic sharded Cholesky over {world} GPUs (NVLink peer exchange per block step), factor replicated, grid split {world} ways"),
            "l2_policy": f"inputs larger than L2: the factor is {8.0 * N * N / 1e6:.0f} MB and is rewritten every step",
            "engine_options": list(args.opt)}


def blas_threads():
    from threadpoolctl import threadpool_info

    return [f"{i.get('internal_api')}:{i.get('num_threads')}" for i in threadpool_info()]


def cpu_sample(orc, spec, X, y, Xs, n_fact, Ms):
    """The reference's per-call work on the host cores, bounded: ``n_fact`` complete factorisations (K-build + dpotrf + v on
    the full N) and, for each, the conditional on the first ``Ms`` grid points.  The solve is linear in the number of grid
    points, so the time of one full cold call is  t_fact + t_solve(Ms) * M / Ms."""
    M = len(Xs)
    t_fact = t_solve = 0.0
    out = None
    for _ in range(n_fact):
        t0 = time.perf_counter()
        L, v = orc.factorize(spec, X, y)
        t1 = time.perf_counter()
        out = orc.conditional(spec, X, L, v, Xs[:Ms], True)
        t2 = time.perf_counter()
        del L
        t_fact += t1 - t0
        t_solve += t2 - t1
    return t_fact / n_fact, t_solve / n_fact, t_fact / n_fact + (t_solve / n_fact) * M / Ms, out


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU path (numpy/scipy restatement of PyMC's Marginal.predict) on the host cores
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from threadpoolctl import threadpool_limits

    from oracle import gp_oracle as orc

    cores = os.cpu_count() or 1
    threadpool_limits(limits=cores)          # belt and braces on top of the re-exec at the top of this file
    spec, X, y, Xs, desc = make_workload(args.workload)
    M, N = len(Xs), len(y)
    # calibration = warm-up: one factorisation + a 1/16 grid slice
    Ms = max(64, M // 16)
    t_f, t_s, t_full, _ = cpu_sample(orc, spec, X, y, Xs, 1, Ms)
    budget = 150.0
    steps = args.steps
    if (steps + max(0, args.warmup - 1)) * t_full <= budget:
        for _ in range(max(0, args.warmup - 1)):
            orc.predict(spec, X, y, Xs, True)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.predict(spec, X, y, Xs, True)
        dt = (time.perf_counter() - t0) / steps
        sample = f"{steps} full cold predict calls, N={N}, grid M={M}"
    else:
        n_fact = int(max(1, min(steps, budget / max(t_f + t_s, 1e-9))))
        t_f, t_s, dt, _ = cpu_sample(orc, spec, X, y, Xs, n_fact, Ms)
        sample = (f"{n_fact} cold calls, each with the complete factorisation (K-build + dpotrf + v, N={N}: {t_f:.1f} s) and the conditional on the "
                  f"first {Ms} of {M} grid points ({t_s:.2f} s); step time = t_fact + t_solve*M/Ms (the solve is linear in M)")
    val = M / dt
    line = {
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "predictions/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args, world, N, M, X.shape[1]),
        "cpu_baseline": {"value": val, "unit": "predictions/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}; numpy/scipy restatement of the PyMC path (PyMC itself is not installable here), "
                                   f"BLAS threads {blas_threads()}, OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS')}"},
        "e2e": {"value": val, "unit": "predictions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "cholesky_gflops_cpu": N ** 3 / 3 / max(1e-9, t_f) / 1e9,
    }
    print(json.dumps(line), flush=True)


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


def cublas_dgemm_sustained(torch, dev, seconds=1.5):
    """cuBLAS DGEMM 8192^3 back to back for ~seconds (the figure a kernel timed inside a long step should be held against)."""
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64)
    b = torch.randn(n, n, device=dev, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize(dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(3, int(seconds / 0.031))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / ms / 1e9


def device_arm(torch, dev, args, workload, precision, steps, warmup, rank, local_rank, world, barrier, max_over_ranks, use_dist, sample_clocks):
    """Cold steps of one workload with everything resident in HBM.  Returns a dict (phases, value, warm, results)."""
    from gumbi_b200 import GPEngine
    from gumbi_b200 import dist as gdist

    spec, X, y, Xs, desc = make_workload(workload)
    N, D_in, M = len(y), X.shape[1], len(Xs)
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
        cnt = [gdist.grid_slice(M, r, world)[1] - gdist.grid_slice(M, r, world)[0] for r in range(world)]
        return (np.concatenate([a[r, 0, :cnt[r]] for r in range(world)]), np.concatenate([a[r, 1, :cnt[r]] for r in range(world)]))

    for _ in range(warmup):
        step_dev()
    phase = {k: 0.0 for k in ("prep_ms", "kbuild_ms", "cholesky_ms", "kstar_ms", "solve_ms", "reduce_ms")}
    launches = 0
    launches_predict = 0
    sampler = ClockSampler(local_rank) if sample_clocks else None
    barrier()
    if sampler and rank == 0:
        sampler.start()
    eng.mark(0)
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        step_dev()
        tm = eng.timings()
        for k in phase:
            phase[k] += tm[k]
        launches += int(tm["launches_factorize"] + tm["launches_predict"])
        launches_predict = int(tm["launches_predict"])
    eng.mark(1)
    ms_total = eng.elapsed_ms(0, 1)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if (sampler and rank == 0) else None
    ms_step = max_over_ranks(ms_total) / steps
    for k in phase:
        phase[k] = max_over_ranks(phase[k] / steps)
    # warm predicts (factor resident)
    eng.mark(2)
    for _ in range(steps):
        eng.predict_device(dXs.data_ptr(), hi - lo, True, dloc.data_ptr(), dloc.data_ptr() + 8 * slot)
        if use_dist and not shard:
            eng.allgather_device(dloc.data_ptr(), dall.data_ptr(), 2 * slot)
    eng.mark(3)
    warm_ms = max_over_ranks(eng.elapsed_ms(2, 3) / steps)
    mu_dev, var_dev = gathered()
    eng.close()
    del eng, dX, dy, dXs, dloc, dall
    torch.cuda.empty_cache()
    return {"spec": spec, "X": X, "y": y, "Xs": Xs, "desc": desc, "N": N, "D_in": D_in, "M": M, "lo": lo, "hi": hi, "shard": shard,
            "ms_step": ms_step, "value": M / (ms_step * 1e-3), "phase": phase, "launches": launches, "wall_ms": wall_ms / steps,
            "clocks": clocks, "warm_ms": warm_ms, "mu": mu_dev, "var": var_dev, "launches_predict": launches_predict}


def rooflines(r, precision, world, peaks, dgemm_peak, dgemm_sustained, copy_gbs, workload):
    N, D_in, phase = r["N"], r["D_in"], r["phase"]
    Ml = r["hi"] - r["lo"]
    kb_bytes = 8.0 * N * (N + 1) / 2 + 8.0 * N * D_in
    kb_gbs = kb_bytes / world / (phase["kbuild_ms"] * 1e-3) / 1e9
    chol_tflops = N ** 3 / 3 / (phase["cholesky_ms"] * 1e-3) / 1e12          # aggregate over the ranks
    solve_tflops = float(N) * N * Ml / (phase["solve_ms"] * 1e-3) / 1e12      # per GPU (each rank solves its own grid slice)
    nblk = (N + 1 + 127) // 128
    fp64_src = ("builder-measured cuBLAS DGEMM 8192^3 in this run, burst = best of 5 (MEASURED_PEAKS.json has no fp64 entry; tcgen05 has no "
                "fp64 kind, DMMA is the fp64 tensor path)")
    tf32_peak = peaks["bf16_tflops"] / 2.0
    if precision == "fp64":
        # GEMM launches of one predict solve: everything the predict call launches except prep, K*-build and the reduction
        # (4 row slabs in concurrent streams by default, each with its own recursion of 2 * nblk - 1 products)
        solve_launches = max(1, r.get("launches_predict", 2 * nblk + 2) - 3)
        tr = measured_traffic(f"{workload}:fp64:solve") or {}
        roofline = {
            "kernel": "dgemm_tma_kernel (TMA-staged DMMA m8n8k4 fp64) in the predict triangular solve L^-1 K(X,X*)",
            "bound": "tensor", "achieved": solve_tflops, "peak": dgemm_peak, "unit": "TFLOP/s", "frac": solve_tflops / dgemm_peak,
            "peak_source": fp64_src, "peak_sustained": dgemm_sustained, "frac_of_sustained": solve_tflops / dgemm_sustained,
            "algorithmic_flop_per_step": float(N) * N * Ml, "launches_per_step": solve_launches,
            "avg_launch_ms": phase["solve_ms"] / solve_launches,
            "traffic": tr.get("dram_bytes_per_launch") if world == 1 else None, "traffic_source": tr.get("source"),
        }
        roofline_chol = {"kernel": "blocked Cholesky (potrf_diag + DMMA panel/trailing update)", "bound": "tensor",
                         "achieved": chol_tflops / world, "peak": dgemm_peak, "unit": "TFLOP/s per GPU (N^3/3 over the whole factorisation)",
                         "frac": chol_tflops / world / dgemm_peak, "peak_sustained": dgemm_sustained,
                         "frac_of_sustained": chol_tflops / world / dgemm_sustained, "aggregate_tflops": chol_tflops,
                         "algorithmic_flop_per_step": N ** 3 / 3, "peak_source": fp64_src}
    else:
        roofline = {
            "kernel": "gemm_tf32x3_kernel (tcgen05.mma kind::tf32, 3 MMAs per product) in the predict triangular solve L^-1 K(X,X*)",
            "bound": "tensor", "achieved": 3.0 * solve_tflops, "peak": tf32_peak, "unit": "TFLOP/s", "frac": 3.0 * solve_tflops / tf32_peak,
            "peak_source": f"half of the bf16 GEMM peak of {peaks['source']} (tf32 dense rate = 1/2 bf16)",
            "fp64_equivalent_tflops": solve_tflops, "algorithmic_flop_per_step": 3.0 * float(N) * N * Ml,
            "note": "achieved counts the 3 tf32 MMAs issued per fp64-equivalent product; the fp64 leaf sub-solves (512 columns) are inside the timed phase",
            "traffic": None,
        }
        roofline_chol = {"kernel": "blocked Cholesky (fp64 DMMA panels + tcgen05 split-TF32 trailing SYRK)", "bound": "tensor",
                         "achieved": 3.0 * chol_tflops / world, "peak": tf32_peak, "unit": "TFLOP/s per GPU (3 tf32 MMAs per product of N^3/3)",
                         "frac": 3.0 * chol_tflops / world / tf32_peak, "fp64_equivalent_tflops_per_gpu": chol_tflops / world,
                         "aggregate_fp64_equivalent_tflops": chol_tflops, "algorithmic_flop_per_step": N ** 3,
                         "peak_source": f"half of the bf16 GEMM peak of {peaks['source']}",
                         "note": "upper bound on the tensor work: the fp64 panel factorisations (DMMA) are inside the timed phase and counted as tf32 products"}
    roofline_kb = {"kernel": "kbuild_persist_kernel<TRAIN> v6 (single-term models: C2, C3, C4) / kbuild_dmma_kernel (Linear, additive models)", "bound": "hbm",
                   "achieved": kb_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kb_gbs / peaks["hbm_gbs"], "peak_source": peaks["source"],
                   "algorithmic_bytes_per_launch": kb_bytes / world, "avg_launch_ms": phase["kbuild_ms"],
                   "traffic": (measured_traffic(f"{workload}:kbuild") or {}).get("dram_bytes_per_launch") if world == 1 else None,
                   "device_copy_gbs_measured_live": copy_gbs, "frac_of_live_copy": (kb_gbs / copy_gbs) if copy_gbs else None}
    return roofline, roofline_chol, roofline_kb, chol_tflops


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

    precision = args.precision
    warmup = max(3, args.warmup)

    # ---- device-resident arm ("value") --------------------------------------------------------------------------
    r = device_arm(torch, dev, args, args.workload, precision, args.steps, warmup, rank, local_rank, world, barrier, max_over_ranks, use_dist, True)
    spec, X, y, Xs = r["spec"], r["X"], r["y"], r["Xs"]
    N, D_in, M, lo, hi = r["N"], r["D_in"], r["M"], r["lo"], r["hi"]
    mu_dev, var_dev = r["mu"], r["var"]

    # ---- multi-GPU exactness: rank 0 repeats the step on ONE GPU and compares a 256-point subsample -------------------------
    dist_check = None
    if use_dist:
        if rank == 0:
            sel = np.random.default_rng(0).choice(M, min(256, M), replace=False)
            e1 = GPEngine(local_rank, precision)
            e1.set_train(X, y)
            e1.set_kernel(spec)
            e1.factorize()
            mu1, var1 = e1.predict(Xs[sel], True)
            e1.close()
            del e1
            torch.cuda.empty_cache()
            dm = float(np.max(np.abs(mu_dev[sel] - mu1)) / np.max(np.abs(mu1)))
            dv = float(np.max(np.abs(var_dev[sel] - var1) / np.abs(var1)))
            tol = 1e-9 if precision == "fp64" else 1e-2
            dist_check = {"points": int(len(sel)), "max_rel_dev_mean_vs_1gpu": dm, "max_rel_dev_var_vs_1gpu": dv, "tolerance": tol,
                          "what": f"{world}-GPU sharded result vs the single-GPU CUDA path on rank 0 (itself parity-tested against the oracle at this size in tests/)"}
            assert dm <= tol and dv <= tol, f"multi-GPU result deviates from the single-GPU path: {dist_check}"
        barrier()

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
        del gp
        torch.cuda.empty_cache()

    # ---- the other single-GPU configurations, short, same process (N = 1 only) ------------------------------------------------
    also = None
    if world == 1 and not args.no_also:
        also = {}
        for name, wl, prec in (("c2_fp64", "c2", "fp64"), ("c3_fp64", "c3", "fp64"), ("c4_tf32", "c4", "tf32"), ("c4_fp64", "c4", "fp64")):
            if wl == args.workload and prec == precision:
                continue
            try:
                ra = device_arm(torch, dev, args, wl, prec, max(3, min(args.steps, 5)), 3, 0, local_rank, 1, barrier, max_over_ranks, False, False)
                also[name] = {"value": ra["value"], "unit": "predictions/s", "ms_per_step": ra["ms_step"], "phases_ms": ra["phase"],
                              "warm_ms_per_step": ra["warm_ms"], "N": ra["N"], "M": ra["M"], "steps": max(3, min(args.steps, 5)),
                              "_res": ra}
            except Exception as ex:   # an extra line must never take the headline down
                also[name] = {"error": repr(ex)}

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

    # ---- rooflines ---------------------------------------------------------------------------------------------------
    peaks = measured_peaks()
    dgemm_peak = cublas_dgemm_peak(torch, dev)
    dgemm_sus = cublas_dgemm_sustained(torch, dev)
    copy_gbs = live_copy_bandwidth(torch, dev)
    roofline, roofline_chol, roofline_kb, chol_tflops = rooflines(r, precision, world, peaks, dgemm_peak, dgemm_sus, copy_gbs, args.workload)
    if also:
        for name, a in also.items():
            ra = a.pop("_res", None)
            if ra is None:
                continue
            prec = "tf32" if name.endswith("tf32") else "fp64"
            rf, rc, rk, ct = rooflines(ra, prec, 1, peaks, dgemm_peak, dgemm_sus, copy_gbs, name[:2])
            a["roofline"] = {k: rf[k] for k in ("achieved", "peak", "frac", "unit")}
            a["roofline_cholesky"] = {k: rc[k] for k in ("achieved", "peak", "frac", "unit")}
            a["roofline_kbuild"] = {k: rk[k] for k in ("achieved", "peak", "frac", "unit")}
            if prec == "tf32" and "c4_fp64" not in also and args.workload == "c4" and precision == "fp64":
                # split-TF32 against this run's fp64 result (north_star gate 1e-2)
                a["max_rel_dev_mean_vs_fp64"] = float(np.max(np.abs(ra["mu"] - mu_dev)) / np.max(np.abs(mu_dev)))
                a["max_rel_dev_var_vs_fp64"] = float(np.max(np.abs(ra["var"] - var_dev) / np.abs(var_dev)))

    # ---- CPU baseline on the host cores: bounded sample of the same workload ----------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:
        from threadpoolctl import threadpool_limits

        from oracle import gp_oracle as orc

        cores = os.cpu_count() or 1
        threadpool_limits(limits=cores)
        if 8.0 * N * N <= 2.5e9:
            t0 = time.perf_counter()
            mu_c, var_c = orc.predict(spec, X, y, Xs, True)
            t_cpu = time.perf_counter() - t0
            cpu = {"value": M / t_cpu, "unit": "predictions/s", "cores": cores, "kind": "port",
                   "sample": f"1 full cold predict call (N={N}, M={M}) of the numpy/scipy restatement of the PyMC path, {t_cpu:.2f} s, BLAS threads {blas_threads()}",
                   "max_rel_err_mean_vs_gpu": float(np.max(np.abs(mu_dev - mu_c)) / np.max(np.abs(mu_c))),
                   "max_rel_err_var_vs_gpu": float(np.max(np.abs(var_dev - var_c) / np.abs(var_c)))}
        else:
            # Bounded (about 10-30 s of CPU work): the reference's per-call work on the leading Ns = N/2 training points and
            # M/16 grid points, each phase timed and scaled by its own exact operation count to the full problem:
            #   K-build (Ns^2 entries) x (N/Ns)^2,  dpotrf (Ns^3/3) x (N/Ns)^3,  conditional (Ns^2 Ms) x (N/Ns)^2 (M/Ms).
            # The reference arm (--impl reference) runs complete factorisations at the full N and is the number the driver compares.
            import scipy.linalg as sla

            Ns, Ms = N // 2, max(64, M // 16)
            t0 = time.perf_counter()
            K = orc.train_cov(spec, X[:Ns])
            t1 = time.perf_counter()
            L = sla.cholesky(K, lower=True, check_finite=False, overwrite_a=True)
            v = sla.solve_triangular(L, y[:Ns], lower=True, check_finite=False)
            t2 = time.perf_counter()
            orc.conditional(spec, X[:Ns], L, v, Xs[:Ms], True)
            t3 = time.perf_counter()
            del K, L
            f = N / Ns
            t_cpu = (t1 - t0) * f ** 2 + (t2 - t1) * f ** 3 + (t3 - t2) * f ** 2 * (M / Ms)
            cpu = {"value": M / t_cpu, "unit": "predictions/s", "cores": cores, "kind": "port",
                   "sample": (f"bounded: the per-call work of the numpy/scipy restatement of the PyMC path on the leading {Ns} of {N} training points and "
                              f"{Ms} of {M} grid points (K-build {t1 - t0:.2f} s, dpotrf+v {t2 - t1:.2f} s, conditional {t3 - t2:.2f} s), each phase scaled by "
                              f"its operation count to the full problem -> {t_cpu:.1f} s per cold call; BLAS threads {blas_threads()}"),
                   "cholesky_gflops_cpu": Ns ** 3 / 3 / (t2 - t1) / 1e9}

    line = {
        "metric": METRIC,
        "value": r["value"], "unit": "predictions/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": r["ms_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if precision == "fp64" else "tf32x3+f64", "data": "synthetic",
        "config": config_dict(args, world, N, M, D_in),
        "phases_ms": r["phase"], "wall_ms_per_step": r["wall_ms"],
        "warm": {"value": M / (r["warm_ms"] * 1e-3), "unit": "predictions/s", "ms_per_step": r["warm_ms"]},
        "cholesky_tflops": chol_tflops,
        "roofline": roofline, "roofline_cholesky": roofline_chol, "roofline_kbuild": roofline_kb,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": r["launches"], "clocks": r["clocks"], "also": also, "multi_gpu_check": dist_check,
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
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp64", choices=["fp64", "tf32"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-also", action="store_true", help="skip the short runs of the other single-GPU configurations")
    ap.add_argument("--opt", action="append", default=[], help="engine tunable name=value (e.g. tf32_nb=8, kbuild_v1=1); ablations only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
