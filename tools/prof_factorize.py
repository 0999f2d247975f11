"""Profiling driver (dev tool): a few factorize + predict calls at a given size, nothing else.
    python tools/prof_factorize.py [n] [reps] [kind] [precision] [region]
region = "predict": cudaProfilerStart/Stop bracket the LAST predict call only (use with ncu --profile-from-start off)."""
import sys
sys.path.insert(0, ".")
from gumbi_b200 import GPEngine
from gumbi_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "ExpQuad"
prec = sys.argv[4] if len(sys.argv) > 4 else "fp64"
region = sys.argv[5] if len(sys.argv) > 5 else ""
spec, X, y, Xs = synthetic_problem(n, 8, M_res=100, kind=kind)
eng = GPEngine(0, prec)
eng.set_train(X, y); eng.set_kernel(spec)
for i in range(reps):
    eng.factorize()
    last = region == "predict" and i == reps - 1
    if last:
        import torch
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    eng.predict(Xs, True)
    if last:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
print(eng.timings())
