"""Kernel microbenchmarks SURVEY.md 8(d) asks for: the stand-alone hash-grid encode (perf_hashgrid_fwd) and the fused
encode + MLP (perf_network_fwd) on N points that are (a) uniform in [0,1)^3 -- worst-case gather locality -- and
(b) ordered ray-major along the rays of the benchmark panorama -- the locality the modular path actually sees.
Prints Msamples/s and algorithmic GB/s (512 B per sample and field).

    python tools/encode_microbench.py [log2_N=22]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perf_b200 import ops  # noqa: E402
from perf_b200.config import APP_MLP, GEO_MLP, PERF_GRID  # noqa: E402

N = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
g = torch.Generator().manual_seed(0)


def params(mlp):
    n = mlp.n_params + 2 * PERF_GRID.n_entries
    return ((torch.rand(n, generator=g) * 2 - 1) * 0.3).cuda()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


uniform = torch.rand(N, 3, generator=g).cuda()
S = 128
o, d = ops.raygen_pano(torch.eye(4), 1024, 2048, row0=512 - N // (S * 2048) // 2, rows=max(1, N // (S * 2048)))
t = (1e-2 + (torch.arange(S, device="cuda", dtype=torch.float32) + 0.5) * (0.99 / S))
ray_major = ((o.reshape(-1, 1, 3) + d.reshape(-1, 1, 3) * t[None, :, None] + 1) / 2).reshape(-1, 3)[:N].clamp(0, 1).contiguous()
for name, mlp in (("density", GEO_MLP), ("colour", APP_MLP)):
    half = ops.params_to_half(params(mlp))
    table = half[mlp.n_params:].view(-1, 2)
    for pts_name, pts in (("uniform", uniform), ("ray-major", ray_major)):
        n = pts.shape[0]
        ms_e = timed(lambda: ops.hashgrid_fwd(table, pts, PERF_GRID))
        ms_n = timed(lambda: ops.network_fwd(half, pts, PERF_GRID, mlp))
        print(f"{name:8s} {pts_name:10s} N={n}: encode {n / ms_e / 1e3:8.1f} Msamples/s ({512e-9 * n / (ms_e * 1e-3):7.1f} GB/s)   "
              f"encode+MLP {n / ms_n / 1e3:8.1f} Msamples/s ({512e-9 * n / (ms_n * 1e-3):7.1f} GB/s)")
