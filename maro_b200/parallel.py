"""Multi-GPU plumbing: replicas are independent (the reference runs one OS process per env, env_process.py:26-67), so
the batch is sharded across ranks with no data-path collective; the single collective collates per-replica episode
metrics.  Backend "nccl" on GPUs (NVLink/NVSwitch), "gloo" in the CPU tests."""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_replicas: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of global replica ids owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_replicas, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def owner_of(replica: int, n_replicas: int, world: int) -> int:
    base, rem = divmod(n_replicas, world)
    cut = rem * (base + 1)
    return replica // (base + 1) if replica < cut else rem + (replica - cut) // max(base, 1)


def gather_metrics(local_metrics, n_replicas: int, group=None):
    """all_gather of per-replica metric rows [local, 3] int64 -> [n_replicas, 3] on every rank, in global replica
    order.  `local_metrics` is a torch tensor on the device the process group communicates on."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_replicas, r, world) for r in range(world)]
    max_local = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    assert local_metrics.shape[0] == hi - lo
    pad = torch.zeros((max_local, local_metrics.shape[1]), dtype=local_metrics.dtype, device=local_metrics.device)
    pad[: hi - lo] = local_metrics
    out = torch.empty((world * max_local, local_metrics.shape[1]), dtype=local_metrics.dtype, device=local_metrics.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    rows = [out[r * max_local: r * max_local + (h - l)] for r, (l, h) in enumerate(sizes)]
    return torch.cat(rows, 0)


def scatter_actions(actions: np.ndarray, n_replicas: int, rank: int, world: int) -> np.ndarray:
    """Slice a global action array [n_replicas, ...] down to this rank's shard."""
    lo, hi = shard_range(n_replicas, rank, world)
    return actions[lo:hi]
