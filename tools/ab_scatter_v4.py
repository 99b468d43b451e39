"""A/B of the fine-level grid scatter: 8-byte atomics (PERF_B200_SCATTER_V4=0) vs 16-byte vector atomics for
x-neighbour pairs (the default since this measurement: 0.659 -> 0.586 ms on a B200).  Same inputs as one 8192 x 128 training step; prints ms per call and the max difference."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perf_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
R, S = 8192, 128
o = ((torch.rand(R, 3, generator=g) - 0.5) * 0.2).cuda()
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).cuda()
jit = torch.rand(R, generator=g).cuda()
dfeat = torch.randn(R * S, 32, generator=g).cuda()
out = torch.zeros(ops.PERF_GRID.n_entries, 2, device="cuda")


def run(v4, iters=20, coarse=False):
    os.environ["PERF_B200_SCATTER_V4"] = "1" if v4 else "0"
    os.environ["PERF_B200_SCATTER_V4_COARSE"] = "1" if coarse else "0"
    for _ in range(3):
        ops.hashgrid_bwd_rays(o, d, jit, S, 1e-2, 1.0, dfeat, out=out.zero_())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.hashgrid_bwd_rays(o, d, jit, S, 1e-2, 1.0, dfeat, out=out)
    e1.record()
    torch.cuda.synchronize()
    res = ops.hashgrid_bwd_rays(o, d, jit, S, 1e-2, 1.0, dfeat, out=out.zero_()).clone()
    return e0.elapsed_time(e1) / iters, res


os.environ["PERF_B200_SCATTER_OVERLAP"] = "0"
t0, r0 = run(False)
t1, r1 = run(True)
t2, r2 = run(True, coarse=True)
os.environ["PERF_B200_SCATTER_OVERLAP"] = "1"
t3, r3 = run(True, coarse=True)
print(f"coarse launch on a side stream (default): {t3:.3f} ms; max|diff| {(r0 - r3).abs().max().item():.3e}")
print(f"coarse+fine scatter: 8-byte atomics {t0:.3f} ms, pair atomics on the fine levels (default) {t1:.3f} ms, also on the coarse flush {t2:.3f} ms; "
      f"max|diff| {(r0 - r1).abs().max().item():.3e} / {(r0 - r2).abs().max().item():.3e} of {r0.abs().max().item():.3e}")
