"""composite_bwd_ray of perf_b200/csrc/train.cu (backward through background rule, accumulate, transmittance scan,
distortion loss and trunc_exp -- or, in the colour phase, through the detached-weight accumulate and the sigmoid)
compiled for the host (tests/host_harness.py) against torch autograd through the oracle's composite, with and
without ray splitting into segments."""
import numpy as np
import pytest
import torch

import host_harness as hh
import oracle


def _forward(raw, ts, te, noise4):
    """Training-mode composite of `nerf_renderer.py:170-194` from raw density logits [R,S] (fp32 torch)."""
    sigma = torch.exp(raw)
    w, T, _ = oracle.render_weight_from_density(ts, te, sigma)
    m, iv = (ts + te) * 0.5, te - ts
    op = w.sum(-1)
    dist_acc = (w * m).sum(-1)
    dist_out = torch.relu(dist_acc + (noise4[:, 3] * 2 - 1) * (1 - op))
    # distortion-loss numerator per ray (torch_efficient_distloss)
    w_ex, wm_ex = torch.cumsum(w, -1) - w, torch.cumsum(w * m, -1) - w * m
    dl = (iv * w * w).sum(-1) / 3.0 + 2.0 * (w * (m * w_ex - wm_ex)).sum(-1)
    return sigma, w, T, op, dist_acc, dist_out, dl


def _rows(x):                       # [R,S,...] -> sample-major rows [S*R,...]
    return x.transpose(0, 1).reshape((-1,) + tuple(x.shape[2:])).contiguous()


@pytest.mark.parametrize("S,segments", [(16, 1), (32, 4), (48, 16), (5, 1)])
@pytest.mark.parametrize("jittered", [False, True])
def test_density_phase_matches_autograd(S, segments, jittered):
    g = torch.Generator().manual_seed(S + segments)
    R, near, far = 40, 1e-2, 1.0
    jitter = torch.rand(R, generator=g) if jittered else torch.zeros(R)
    step = (torch.tensor(far) - torch.tensor(near)) / torch.tensor(float(S))
    k = torch.arange(S, dtype=torch.float32)[None, :]
    ts, te = near + (k + jitter[:, None]) * step, near + (k + 1 + jitter[:, None]) * step
    raw = (torch.randn(R, S, generator=g) * 1.5 + 1.0).requires_grad_(True)
    raw.data[0] = 8.0                                    # an opaque ray: transmittance collapses after a few samples
    raw.data[1] = -20.0                                  # an empty ray: opacity ~ 0, the background rule dominates
    raw.data[2, S // 2] = 17.0                           # a logit beyond the slope cap of trunc_exp (15)
    noise4 = torch.rand(R, 4, generator=g)
    sigma, w, T, op, dist_acc, dist_out, dl = _forward(raw, ts, te, noise4)
    g_dist, g_op, g_dl = torch.randn(R, generator=g), torch.randn(R, generator=g), torch.randn(R, generator=g) * 0.3
    # trunc_exp's backward caps the slope: d sigma / d raw := exp(min(raw, 15)); everything else is plain autograd
    (d_sigma,) = torch.autograd.grad((g_dist * dist_out + g_op * op + g_dl * dl).sum(), sigma)
    want = _rows((d_sigma * torch.exp(raw.detach().clamp(max=15.0)))).numpy()
    # saved tensors as the forward kernel writes them: segment-local w / T plus the transmittance at each segment start
    kps = S // segments
    t_start = T.detach()[:, ::kps]                                              # [R, segments]
    local = t_start.repeat_interleave(kps, dim=1)
    toff = torch.ones(16, R)
    toff[:segments] = t_start.t()
    safe = local.clamp(min=1e-30)
    got = hh.composite_backward(1, S, segments, near, far, jitter.numpy() if jittered else None, noise4.numpy(),
                                _rows(sigma.detach()).numpy(), _rows(w.detach() / safe).numpy(), _rows(T.detach() / safe).numpy(), None,
                                dist_acc.detach().numpy(), toff.numpy() if segments > 1 else None, None, g_dist.numpy(), g_op.numpy(),
                                g_dl.numpy(), dist_out.detach().numpy(), op.detach().numpy())
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2e-4 * scale, (np.abs(got - want).max(), scale)


@pytest.mark.parametrize("S,segments", [(16, 1), (32, 8)])
def test_colour_phase_matches_autograd(S, segments):
    g = torch.Generator().manual_seed(3 * S)
    R, near, far = 24, 1e-2, 1.0
    step = (far - near) / S
    k = torch.arange(S, dtype=torch.float32)[None, :].expand(R, S)
    ts, te = near + k * step, near + (k + 1) * step
    sigma = torch.exp(torch.randn(R, S, generator=g))
    w, T, _ = oracle.render_weight_from_density(ts, te, sigma)
    z = torch.randn(R, S, 3, generator=g).requires_grad_(True)
    y = torch.sigmoid(z)
    y16 = y.detach().half()
    g_rgb = torch.randn(R, 3, generator=g)
    # colours = sum_k w.detach() * rgb (nerf_renderer.py:183); the kernel uses the SAVED fp16 colour for y (1 - y)
    want = _rows(g_rgb[:, None, :] * w[..., None] * (y16.float() * (1 - y16.float()))).numpy()
    kps = S // segments
    t_start = T[:, ::kps]
    local = t_start.repeat_interleave(kps, dim=1)
    toff = torch.ones(16, R)
    toff[:segments] = t_start.t()
    rgb4 = torch.zeros(S * R, 4, dtype=torch.float16)
    rgb4[:, :3] = _rows(y16)
    got = hh.composite_backward(2, S, segments, near, far, None, None, None, _rows(w / local).numpy(), None, rgb4.numpy(), None,
                                toff.numpy() if segments > 1 else None, g_rgb.numpy(), None, None, None, None, None)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
