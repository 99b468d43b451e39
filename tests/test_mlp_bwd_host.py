"""The single-kernel MLP backward (perf_b200/csrc/mlp_bwd.cu, experimental): its CUDA-core twin -- the per-thread
phases the kernel runs between barriers, reading the same shared-memory operand images the tensor-core path
describes to tcgen05 -- emulated for one CTA on the host (tests/host_harness.py) against the plain matrix formulas.
This pins the image layout, the "transposed image" reads, the block layout of the weight-gradient accumulators and
the flush; the tcgen05 descriptors themselves can only be checked on a GPU (tests/test_gpu_train.py, gated)."""
import numpy as np
import pytest
import torch

import host_harness as hh
from perf_b200.config import APP_MLP, GEO_MLP


@pytest.mark.parametrize("two_hidden", [False, True], ids=["density", "colour"])
@pytest.mark.parametrize("N", [128, 128 * 3 + 41])
def test_cuda_core_twin_on_host_matches_matrix_formulas(two_hidden, N):
    mlp = APP_MLP if two_hidden else GEO_MLP
    g = torch.Generator().manual_seed(N + two_hidden)
    W = ((torch.rand(mlp.n_params, generator=g) * 2 - 1) * 0.3).half()
    feat = ((torch.rand(N, 32, generator=g) * 2 - 1) * 0.5).half()
    w1 = W[:2048].view(64, 32).double()
    h1 = torch.relu(feat.double() @ w1.t()).half()
    p = 2048
    w2 = h2 = None
    if two_hidden:
        w2 = W[p:p + 4096].view(64, 64).double()
        h2 = torch.relu(h1.double() @ w2.t()).half()
        p += 4096
    wout = W[p:p + 16 * 64].view(16, 64)[:mlp.n_out].double()
    dz = torch.randn(N, mlp.n_out, generator=g) * 0.1
    got_w, got_f = hh.mlp_backward(mlp, W.numpy(), feat.numpy(), h1.numpy(), None if h2 is None else h2.numpy(), dz.numpy())

    h_last = h2 if two_hidden else h1
    dz16 = dz.half().double()                                         # the weight-gradient product reads dz as an fp16 operand
    want_wout = dz16.t() @ h_last.double()
    dh_last = ((dz.double() @ wout) * (h_last > 0)).half()            # output-layer backward in fp32, stored as fp16
    want = {}
    if two_hidden:
        want["w2"] = dh_last.double().t() @ h1.double()
        dh1 = ((dh_last.double() @ w2) * (h1 > 0)).half()
    else:
        dh1 = dh_last
    want["w1"] = dh1.double().t() @ feat.double()
    want_f = dh1.double() @ w1

    def close(got, ref, name, rel=2e-3):
        err = (torch.as_tensor(got).double() - ref).abs().max().item()
        assert err <= rel * ref.abs().max().item() + 1e-7, f"{name}: {err:.3e} vs {ref.abs().max().item():.3e}"
    close(got_f, want_f, "dfeat")
    close(got_w[:2048].reshape(64, 32), want["w1"], "dW1")
    if two_hidden:
        close(got_w[2048:6144].reshape(64, 64), want["w2"], "dW2")
    out_block = got_w[p:p + 16 * 64].reshape(16, 64)
    close(out_block[:mlp.n_out], want_wout, "dWout")
    assert np.all(out_block[mlp.n_out:] == 0)                         # padded rows of the last matrix get no gradient
