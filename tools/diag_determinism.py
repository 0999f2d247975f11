"""Dev tool: run-to-run determinism and ablation agreement of the factorisation (TMA-staged GEMM on/off, persistent K-build on/off).
Prints, per size, where (128-row block coordinates) the factor of a run differs from the cp.async / round-1-K-build run."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine
from gumbi_b200.synthetic import synthetic_problem


def blocks(D, thr):
    bad = np.argwhere(np.abs(D) > thr)
    if len(bad) == 0:
        return "none"
    b = np.unique(bad // 128, axis=0)
    return f"{len(bad)} entries in {len(b)} blocks, first {b[:6].tolist()} last {b[-3:].tolist()}"


def run(N, d, kind, reps, full):
    spec, X, y, Xs = synthetic_problem(N, d, P=1, M_res=20, kind=kind)
    e = GPEngine(0)
    e.set_train(X, y)
    e.set_kernel(spec)
    ref = {}
    base = {"dgemm_tma": 7, "dgemm_persistent": 1, "dgemm_fence": 1, "lookahead": 1, "chain_on_panel": 1, "kbuild_persist": 1, "fp64_panel": FP64_PANEL}
    for tag, delta in CONFIGS:
        opts = {**base, **delta}
        for k, v in opts.items():
            e.set_option(k, v)
        for rep in range(reps):
            e.set_kernel(spec)
            try:
                e.factorize()
            except np.linalg.LinAlgError as ex:
                print(f"N={N} {tag} rep{rep}: {str(ex)[:90]}", flush=True)
                continue
            v = e.get_v()
            mll = e.mll()
            mu, var = e.predict(Xs[:256], True)
            L = e.get_L() if full else None
            if not ref:
                ref = {"v": v, "mll": mll, "mu": mu, "L": L}
                print(f"N={N} {tag} rep{rep}: reference mll={mll:.9f}", flush=True)
                continue
            dv = np.abs(v - ref["v"])
            line = f"N={N} {tag} rep{rep}: dmll={mll - ref['mll']:+.3e} max|dv|={dv.max():.3e} first bad v idx={int(np.argmax(dv > 1e-9)) if (dv > 1e-9).any() else -1} max|dmu|={np.abs(mu - ref['mu']).max():.3e}"
            if full:
                line += " | L: " + blocks(L - ref["L"], 1e-9)
            print(line, flush=True)
    if full:
        for kb in (0, 1):
            e.set_option("kbuild_persist", kb)
            e.set_kernel(spec)
            K = e.get_K()
            if kb == 0:
                K0 = K
            else:
                print(f"N={N} K-build persist vs round-1: " + blocks(K - K0, 1e-12), flush=True)
    e.close()


# dgemm_fence = 0 drops the generic->async proxy fence at the release of a shared-memory stage (the round-2 bug, kept as a control)
CONFIGS = [("cp.async (reference)", {"dgemm_tma": 0}), ("tma, stage fence (product), plain algorithm", {"fp64_panel": 0}),
           ("tma, stage fence (product), two-level blocking", {"fp64_panel": 16}),
           ("tma WITHOUT the stage fence (control), plain algorithm", {"fp64_panel": 0, "dgemm_fence": 0})]

FP64_PANEL = int(os.environ.get("DIAG_FP64_PANEL", "0"))   # 0: the plain algorithm (k = 128 bulk updates, the regime that failed)

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16384]
    for n in sizes:
        run(n, 8, "Matern52", 8, False)
