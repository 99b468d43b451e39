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


def _torch_adam(params, grads, exp_avg, exp_avg_sq, step, lr, params_half=None, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    """perf_adam_step restated with torch ops (CPU stand-in for the CUDA kernel in the host-logic test below)."""
    g = grads * grad_scale
    exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, (1 - beta2 ** step) ** 0.5
    params.addcdiv_(exp_avg, exp_avg_sq.sqrt() / bc2 + eps, value=-lr / bc1)
    if params_half is not None:
        params_half.copy_(params.half())


class _FakeModule:
    def __init__(self, p):
        self.params, self.h = p, p.detach().half()

    def _half(self):
        return self.h


def _opt_worker(rank, world, port, n, mode, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      PERF_B200_DP=mode)
    from perf_b200 import ops, parallel
    from perf_b200.scene import FusedAdam
    parallel.init("gloo")
    ops.adam_step = _torch_adam
    g = torch.Generator().manual_seed(1)
    p = torch.nn.Parameter(torch.randn(n, generator=g))
    mod = _FakeModule(p)
    opt = FusedAdam(p, lr=1e-2, module=mod)
    assert opt.sharded == (mode == "sharded")
    grads = torch.randn(5, world, n, generator=g)                  # 5 steps, one local gradient per rank
    for t in range(5):
        p.grad = grads[t, rank].clone() if not (t == 3 and rank == 1) else None
        opt.step(valid=p.grad is not None)                          # step 3: rank 1 has no samples, must not dead-lock
    half_before = mod._half().clone()
    opt.sync_master()
    q.put((rank, p.detach().clone(), half_before))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,n", [("sharded", 1024), ("sharded", 1001), ("allreduce", 1001)])
def test_sharded_optimizer_equals_full_batch_adam(mode, n):
    """FusedAdam's distributed flows on gloo / CPU (the CUDA Adam kernel replaced by its torch restatement):
    reduce-scatter -> Adam on the 1/world shard -> all-gather of the fp16 shadow (incl. a length that needs a padded
    tail) == all-reduce + replicated Adam == single-process Adam on the rank-averaged gradient; a rank without samples
    joins with a zero gradient."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_opt_worker, args=(r, world, port, n, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1)
    p_ref = torch.randn(n, generator=g)
    grads = torch.randn(5, world, n, generator=g)
    m, v = torch.zeros(n), torch.zeros(n)
    for t in range(5):
        gm = grads[t].clone()
        if t == 3:
            gm[1] = 0
        _torch_adam(p_ref, gm.sum(0), m, v, t + 1, 1e-2, grad_scale=1.0 / world)
    for rank, p, half in res:
        assert torch.allclose(p, p_ref, rtol=1e-6, atol=1e-7), (rank, float((p - p_ref).abs().max()))
        assert torch.equal(half, p_ref.half()) or (half.float() - p_ref).abs().max() < 2e-3
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])      # replicas bit-identical


def _gather_worker(rank, world, port, height, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from perf_b200 import parallel
    parallel.init("gloo")
    sl = parallel.shard_slice(height)
    full = torch.arange(height * 5 * 4, dtype=torch.float32).reshape(height, 5, 4)
    got = parallel.gather_row_tiles(full[sl].clone(), height)
    q.put((rank, None if got is None else bool(torch.equal(got, full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [16, 17])
def test_row_tile_gather_handles_unequal_shards(height):
    """render_dense's tile gather (the only collective on that path): 17 rows on 2 ranks are unequal tiles."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, height, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: None}
