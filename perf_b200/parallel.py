"""Ray-sharded data parallelism: one process per GPU, full field replica per rank, ONE all-reduce
over NVLink per optimisation step on the flat gradient of the active network.

The reference has no distributed code at all; this is the only collective the framework adds.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def init(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank) from the torchrun environment; initialises the process group
    (NCCL on GPUs, gloo on CPU) when WORLD_SIZE > 1."""
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_slice(n: int, r: Optional[int] = None, w: Optional[int] = None) -> slice:
    """Contiguous shard of ``n`` units for rank ``r`` of ``w`` (the first n % w ranks get one more)."""
    r, w = rank() if r is None else r, world_size() if w is None else w
    base, rem = divmod(n, w)
    start = r * base + min(r, rem)
    return slice(start, start + base + (1 if r < rem else 0))


def allreduce_mean_(flat_grad: torch.Tensor) -> torch.Tensor:
    """In-place average of a flat gradient over all ranks: the single collective of a training step.
    The reference's losses are means over the local batch (`nerf.py:211,285`), so equal shards +
    gradient averaging reproduce the single-GPU full-batch gradient."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(w)
    return flat_grad
