"""Host-side plumbing of the multi-GPU path: one process per GPU, ``torch.distributed`` only for rendezvous/result gathers.

The data path itself (row-block-cyclic sharded Cholesky with a per-step NCCL broadcast of the diagonal block and all-gather of
the panel) lives behind ``gb2_dist_init`` / ``gb2_factorize`` in the CUDA library; this module
  * ships rank 0's ``ncclUniqueId`` to the other ranks (any backend: gloo in the CPU tests, nccl on the GPU box),
  * states the ownership maps the library uses (so that tests can pin them), and
  * splits a prediction grid over ranks and gathers the posterior back.
The reference has no counterpart (single process).
"""
from __future__ import annotations

import numpy as np

TILE = 128  # row-block size of the factorisation (gb2::TILE)


def block_owner(block: int, world: int) -> int:
    """Rank owning 128-row block ``block`` (block-cyclic)."""
    return block % world


def owned_blocks(n_blocks: int, rank: int, world: int, after: int = -1):
    """Global indices of the row blocks ``rank`` owns, optionally only those > ``after`` (the panel below step ``after``)."""
    first = after + 1 + ((rank - (after + 1)) % world)
    return list(range(first, n_blocks, world))


def padded_size(N: int) -> int:
    """Rows of the augmented, padded system the library factorises: round_up(N + 1, 128)."""
    return (N + 1 + TILE - 1) // TILE * TILE


def allgather_slots(n_blocks: int, k: int, world: int):
    """Layout of the per-step panel all-gather: (slots_per_rank, {rank: [global blocks in slot order]})."""
    per_rank = {r: owned_blocks(n_blocks, r, world, after=k) for r in range(world)}
    return max((len(v) for v in per_rank.values()), default=0), per_rank


def grid_slice(M: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of an M-point prediction grid served by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(M, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_unique_id(make_id, group=None) -> bytes:
    """Rank 0 calls ``make_id()`` (-> 128 bytes) and broadcasts it; every rank returns the same bytes."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return make_id()
    payload = [make_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(payload, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return payload[0]


def gather_grid(local_mean: np.ndarray, local_var: np.ndarray, M: int, group=None):
    """All-gather the per-rank posterior slices (host arrays) into full (M,) arrays on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_mean, local_var
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, (np.asarray(local_mean), np.asarray(local_var)), group=group)
    mean = np.concatenate([p[0] for p in parts])
    var = np.concatenate([p[1] for p in parts])
    assert mean.shape == (M,) and var.shape == (M,)
    return mean, var


def init_engine(engine, group=None):
    """Make ``engine`` (a GPEngine) part of the process group: collective; afterwards ``engine.factorize()`` is sharded."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 1:
        uid = exchange_unique_id(engine.nccl_unique_id, group)
        engine.dist_init(rank, world, uid)
    return rank, world
