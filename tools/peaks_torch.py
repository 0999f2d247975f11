"""Measure the library comparators on the GPU box: cuBLAS DGEMM / TF32 GEMM peak and cuSOLVER potrf.
Writes gpurun_out/peaks_fp64.json.  Not part of the product path."""
import json, os, time, torch
torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda:0")
res = {"gpu": torch.cuda.get_device_name(0)}
def ev_time(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n = 8192
a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64)
ms = ev_time(lambda: torch.matmul(a, b))
res["dgemm_8192_tflops_burst"] = 2 * n**3 / ms / 1e9
t0 = time.time(); k = 0
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    torch.matmul(a, b); k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
res["dgemm_8192_tflops_sustained"] = 2 * n**3 * k / e0.elapsed_time(e1) / 1e9
af = a.float(); bf = b.float()
ms = ev_time(lambda: torch.matmul(af, bf))
res["tf32gemm_8192_tflops_burst"] = 2 * n**3 / ms / 1e9
del af, bf, b
for N in (8192, 16384, 32768):
    x = torch.randn(N, 64, device=dev, dtype=torch.float64)
    K = torch.exp(-0.5 * torch.cdist(x, x) ** 2 / 64.0); K.diagonal().add_(0.01)
    ms = ev_time(lambda: torch.linalg.cholesky(K), reps=3)
    res[f"cusolver_potrf_{N}_tflops"] = N**3 / 3 / ms / 1e9
    res[f"cusolver_potrf_{N}_ms"] = ms
    del K, x
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/peaks_fp64.json", "w"), indent=1)
print(json.dumps(res))
