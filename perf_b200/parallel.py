"""Ray-sharded data parallelism: one process per GPU, full fp16 replica of the field per rank, ONE gradient
exchange per optimisation step on the flat gradient of the active network.  Default (round 2):
reduce-scatter of the fp32 gradient -> Adam on the rank's 1/N shard of the parameters and moments -> all-gather
of the fp16 shadow the kernels read (SURVEY 8e's variant: the exchange moves 7/8 * (26.6 + 13.3) MB per rank
instead of an all-reduce's 2 * 7/8 * 26.6 MB, and the optimiser pass shrinks N-fold).  ``PERF_B200_DP=allreduce``
selects round 1's all-reduce + replicated Adam.

The reference has no distributed code at all; this is the only exchange the framework adds.
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
        import datetime
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        # failure detection: a rank that dies or stalls must not leave the others spinning in the step's
        # exchange for ever -- NCCL's watchdog aborts the communicator when a collective exceeds the timeout and the
        # error surfaces as an exception on every rank (async error handling); PERF_B200_COMM_TIMEOUT_S overrides 120 s
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        timeout = datetime.timedelta(seconds=float(os.environ.get("PERF_B200_COMM_TIMEOUT_S", "120")))
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timeout)
        else:
            dist.init_process_group(backend, timeout=timeout)
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


def dp_mode() -> str:
    """'sharded' (reduce-scatter / sharded Adam / all-gather of the fp16 shadow; default) or 'allreduce'."""
    return "allreduce" if os.environ.get("PERF_B200_DP", "sharded") == "allreduce" else "sharded"


def shard_len(n: int, w: Optional[int] = None, align: int = 8) -> int:
    """Elements per rank when ``n`` values are cut into ``w`` EQUAL contiguous shards, each a multiple of ``align``
    (NCCL reduce-scatter / all-gather need equal counts; 8 fp16 = 16 bytes keeps every shard vector-aligned).
    ``w * shard_len(n, w) >= n``: the caller pads the tail."""
    w = world_size() if w is None else w
    per = -(-n // w)
    return -(-per // align) * align


def reduce_scatter_sum_(flat: torch.Tensor, shard: int) -> torch.Tensor:
    """In-place reduce-scatter(SUM) of ``flat`` (numel == world * shard): returns this rank's reduced shard, a view of
    ``flat`` (NCCL's in-place form: recvbuf = sendbuf + rank * count)."""
    w, r = world_size(), rank()
    assert flat.numel() == w * shard
    out = flat[r * shard:(r + 1) * shard]
    if w > 1:
        dist.reduce_scatter_tensor(out, flat, op=dist.ReduceOp.SUM)
    return out


def all_gather_(flat: torch.Tensor, shard: int) -> torch.Tensor:
    """In-place all-gather: every rank contributes ``flat[rank*shard:(rank+1)*shard]`` and ends with all of ``flat``."""
    w, r = world_size(), rank()
    assert flat.numel() == w * shard
    if w > 1:
        dist.all_gather_into_tensor(flat, flat[r * shard:(r + 1) * shard])
    return flat


def gather_row_tiles(tile: torch.Tensor, total_rows: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather the row tiles of `shard_slice(total_rows)` ([rows_r, W, C] each) on rank ``dst`` -> [total_rows, W, C]
    there, None elsewhere.  The only collective of render_dense.  NCCL gather needs equal counts: when
    total_rows % world != 0 the tiles are padded to the largest one and cropped on arrival."""
    w, r = world_size(), rank()
    if w == 1:
        return tile
    sizes = [shard_slice(total_rows, i, w) for i in range(w)]
    rows_max = max(s.stop - s.start for s in sizes)
    send = tile.contiguous()
    if send.shape[0] != rows_max:
        pad = torch.zeros((rows_max,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        pad[: send.shape[0]] = send
        send = pad
    recv = [torch.empty_like(send) for _ in range(w)] if r == dst else None
    dist.gather(send, recv, dst=dst)
    if r != dst:
        return None
    return torch.cat([t[: s.stop - s.start] for t, s in zip(recv, sizes)], 0)
