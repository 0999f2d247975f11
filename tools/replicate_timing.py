"""Small-N multi-GPU alternative to the sharded factorisation: every rank factorises the WHOLE K itself (no exchange at all) and
solves only its slice of the prediction grid; the slices are all-gathered (NCCL).  Expected on BASELINE config 2 from the
single-GPU phase times: 9.3 + 24.3 / G ms per cold step, against 21.3 / 15.6 / 14.4 ms of the sharded path at G = 2 / 4 / 8.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/replicate_timing.py [n] [d] [steps]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gumbi_b200 import GPEngine  # noqa: E402
from gumbi_b200 import dist as gdist  # noqa: E402
from gumbi_b200.synthetic import synthetic_problem  # noqa: E402

rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
spec, X, y, Xs = synthetic_problem(n, d)
M = len(Xs)
lo, hi = gdist.grid_slice(M, rank, world)
slot = -(-M // world)
eng = GPEngine(local_rank)                  # never joined to the process group: a plain single-GPU handle per rank
dX, dy = torch.from_numpy(X).to(dev), torch.from_numpy(y).to(dev)
dXs = torch.from_numpy(np.ascontiguousarray(Xs[lo:hi])).to(dev)
dloc = torch.zeros(2 * slot, dtype=torch.float64, device=dev)
dall = torch.zeros(2 * slot * world, dtype=torch.float64, device=dev)
eng.set_train_device(dX.data_ptr(), n, X.shape[1], dy.data_ptr())


def step():
    eng.set_kernel(spec)
    eng.factorize()
    eng.predict_device(dXs.data_ptr(), hi - lo, True, dloc.data_ptr(), dloc.data_ptr() + 8 * slot)
    if world > 1:
        dist.all_gather_into_tensor(dall, dloc)


for _ in range(3):
    step()
torch.cuda.synchronize(dev)
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(dev)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"mode": "replicated factorisation, grid split over ranks", "n_gpus": world, "N": n, "d": d, "M": M,
                      "ms_per_step": float(ms.item()), "predictions_per_s": M / (float(ms.item()) * 1e-3)}), flush=True)
eng.close()
if world > 1:
    dist.destroy_process_group()
