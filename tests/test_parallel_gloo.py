"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: ray sharding and the single
gradient all-reduce of a data-parallel step (perf_b200/parallel.py), plus panorama row tiling."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from perf_b200 import parallel
    r, w, _ = parallel.init("gloo")
    assert (r, w) == (rank, world) and parallel.rank() == rank and parallel.world_size() == world
    # a "field" of 1000 params, a global batch of 8192 rays with per-ray targets; loss = mean over the batch
    g = torch.Generator().manual_seed(0)
    params = torch.randn(1000, generator=g, requires_grad=True)
    feats, target = torch.randn(8192, 1000, generator=g), torch.randn(8192, generator=g)
    sl = parallel.shard_slice(8192)
    loss = ((feats[sl] @ params - target[sl]) ** 2).mean()          # mean over the LOCAL batch (nerf.py:211)
    loss.backward()
    grad = parallel.allreduce_mean_(params.grad.clone())
    full = torch.autograd.grad(((feats @ params.detach().requires_grad_(True) - target) ** 2).mean(),
                               [p for p in [params.detach().requires_grad_(True)]], allow_unused=True)
    p2 = params.detach().clone().requires_grad_(True)
    ((feats @ p2 - target) ** 2).mean().backward()
    q.put((rank, float((grad - p2.grad).abs().max()), sl.start, sl.stop))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_allreduce_equals_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[2:] for r in res] == [(0, 4096), (4096, 8192)]
    assert all(r[1] < 1e-5 for r in res), res


def test_shard_slice_covers_everything_once():
    from perf_b200 import parallel
    for n in (0, 1, 7, 1024, 1025, 2097152):
        for w in (1, 2, 3, 8):
            sl = [parallel.shard_slice(n, r, w) for r in range(w)]
            assert sl[0].start == 0 and sl[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(sl, sl[1:]))
            sizes = [s.stop - s.start for s in sl]
            assert max(sizes) - min(sizes) <= 1
