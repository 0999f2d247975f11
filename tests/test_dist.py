"""Multi-process tests of the N>1 path.

CPU (gloo, world_size 2): the host-side plumbing of gumbi_b200/dist.py -- ownership maps, the unique-id exchange, grid
slicing and the result gather.  GPU (>= 2 devices): the sharded factorisation itself, through torchrun + tests/dist_check.py.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from gumbi_b200 import dist as gdist


def test_block_cyclic_ownership_maps():
    for world in (1, 2, 3, 4, 8):
        for nb in (1, 5, 8, 17):
            seen = []
            for r in range(world):
                blocks = gdist.owned_blocks(nb, r, world)
                assert all(gdist.block_owner(b, world) == r for b in blocks)
                seen += blocks
            assert sorted(seen) == list(range(nb))
            for k in range(nb):
                cmax, per_rank = gdist.allgather_slots(nb, k, world)
                below = sorted(b for v in per_rank.values() for b in v)
                assert below == list(range(k + 1, nb))            # every panel block has exactly one owner/slot
                assert cmax == max(len(v) for v in per_rank.values())
                assert cmax == len(gdist.owned_blocks(nb, (k + 1) % world, world, after=k))  # the library's cmax rule
    assert gdist.padded_size(8192) == 8320 and gdist.padded_size(127) == 128 and gdist.padded_size(128) == 256


def test_grid_slices_partition_the_grid():
    for M in (0, 1, 7, 10000, 10001):
        for world in (1, 2, 3, 8):
            sl = [gdist.grid_slice(M, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [hi - lo for lo, hi in sl]
            assert max(sizes) - min(sizes) <= 1


WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r})
from gumbi_b200 import dist as gdist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
calls = []
def make_id():
    calls.append(1)
    return bytes(range(128))
uid = gdist.exchange_unique_id(make_id)
assert uid == bytes(range(128)) and len(calls) == (1 if rank == 0 else 0), (rank, len(calls))
M = 1001
lo, hi = gdist.grid_slice(M, rank, world)
full = np.arange(M, dtype=np.float64)
mean, var = gdist.gather_grid(full[lo:hi] * 2.0, full[lo:hi] + 0.5, M)
assert np.array_equal(mean, full * 2.0) and np.array_equal(var, full + 0.5)
dist.barrier()
dist.destroy_process_group()
print("WORKER_OK", rank)
"""


def test_unique_id_exchange_and_gather_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"WORKER_OK {r}" in o, o


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["p2p", "nccl", "shard_storage", "tf32"])
def test_sharded_cholesky_matches_single_gpu(mode):
    """Factor bit-identical to the single-GPU one; predictions equal (replicated storage) or within 1e-8 (storage-sharded:
    different summation order); through torchrun + tests/dist_check.py."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_dist.py -m gpu)")
    world = 2 if n < 4 else 4
    env = dict(os.environ)
    if mode == "nccl":
        env["GB2_DIST_P2P"] = "0"
    if mode == "shard_storage":
        env["GB2_DIST_SHARD"] = "1"
    if mode == "tf32":
        env["GB2_DIST_PRECISION"] = "tf32"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_check.py"), "1000", "3000"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and "DIST_CHECK OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


BACKEND_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from conftest import load_golden
from test_backend_host import HostGP, OracleEngine, gp_from_golden
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")

class DistDouble(OracleEngine):
    # engine double that looks like a rank of a replicated-factor job (every rank holds the factor, serves a grid slice)
    world = {world}
    rank = rank
    shard_storage = False
    calls = []
    def predict(self, Xs, pred_noise=True):
        self.calls.append(len(Xs))
        return super().predict(Xs, pred_noise)

g = load_golden("multioutput_regression")
gp = gp_from_golden(g)
gp.engine = DistDouble()
gp.engine.set_train(gp._X, gp._y)
gp.find_MAP(point=g["meta"]["point"])
mu, var = gp.predict(g["points"], with_noise=True)
# every rank gets the FULL result, computed from its own slice only
from gumbi_b200 import dist as gdist
lo, hi = gdist.grid_slice(len(g["points"]), rank, world)
assert gp.engine.calls == [hi - lo], (gp.engine.calls, lo, hi)
np.testing.assert_allclose(mu, g["mean"], rtol=1e-9, atol=1e-11)
np.testing.assert_allclose(var, g["var"], rtol=1e-8, atol=1e-11)
# storage-sharded engines predict collectively over all points: no slicing, no gather
gp.engine.shard_storage = True
gp.engine.calls.clear()
mu2, _ = gp.predict(g["points"], with_noise=True)
assert gp.engine.calls == [len(g["points"])]
np.testing.assert_allclose(mu2, g["mean"], rtol=1e-9, atol=1e-11)
dist.barrier(); dist.destroy_process_group()
print("BACKEND_WORKER_OK", rank)
"""


def test_backend_predict_slices_and_gathers_over_gloo(tmp_path):
    """B200Backend.predict on a 2-rank job: each rank solves its grid slice, the gathered result is the full golden posterior."""
    script = tmp_path / "bworker.py"
    script.write_text(BACKEND_WORKER.format(root=ROOT, world=2))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29535", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"BACKEND_WORKER_OK {r}" in o, o[-3000:]
