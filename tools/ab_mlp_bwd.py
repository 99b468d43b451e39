"""A/B of the MLP backward for one 8192 x 128 training step: default path (cuBLAS fp16 GEMMs + two small
kernels, ops.mlp_backward_half) vs the single tcgen05 kernel (ops.mlp_backward_fused, csrc/mlp_bwd.cu).
Prints ms per call for both networks and the largest differences.  Run the correctness tests first:

    PERF_B200_EXPERIMENTAL=1 python -m pytest tests/test_gpu_train.py -k single_kernel_mlp_backward -x -q --timeout 60
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perf_b200 import ops  # noqa: E402
from perf_b200.config import APP_MLP, GEO_MLP  # noqa: E402

N = 8192 * 128
g = torch.Generator().manual_seed(0)


def timed(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


for name, mlp in (("density", GEO_MLP), ("colour", APP_MLP)):
    W = ((torch.rand(mlp.n_params, generator=g) * 2 - 1) * 0.3).half().cuda()
    feat = ((torch.rand(N, 32, generator=g) * 2 - 1) * 0.5).half().cuda()
    h1 = torch.relu(feat.float() @ W[:2048].view(64, 32).float().t()).half()
    h2 = torch.relu(h1.float() @ W[2048:6144].view(64, 64).float().t()).half() if mlp.n_hidden_layers == 2 else None
    dz = (torch.randn(N, mlp.n_out, generator=g) * 0.1).cuda()
    t_ref, (w_ref, f_ref) = timed(lambda: ops.mlp_backward_gemm(mlp, W, feat, h1, h2, dz))
    for simt in ((True, False) if "--simt" in sys.argv else (False,)):
        t_new, (w_new, f_new) = timed(lambda: ops.mlp_backward_fused(mlp, W, feat, h1, h2, dz, simt=simt), iters=3 if simt else 10)
        print(f"{name:8s} {'simt' if simt else 'tcgen05':8s}: gemm path {t_ref:.3f} ms, single kernel {t_new:.3f} ms; "
              f"max|d dW| {(w_new - w_ref).abs().max().item():.3e} of {w_ref.abs().max().item():.3e}, "
              f"max|d dfeat| {(f_new - f_ref).abs().max().item():.3e} of {f_ref.abs().max().item():.3e}")
