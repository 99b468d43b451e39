"""Input gradients of the hash-grid encode and their double backward:
the oracle itself (CPU) and the autograd wiring of ``tinycudann.Encoding`` with the kernels replaced
by oracle stand-ins (CPU); the kernels are compared with the same oracle in tests/test_gpu_encoding_grad.py."""
import pytest
import torch

from oracle.hashgrid import (GridConfig as OGrid, encode, encode_autograd, encode_input_grad,
                             encode_input_grad_backward, n_table_entries)

SMALL = [OGrid(n_levels=4, log2_hashmap_size=6, base_resolution=2, per_level_scale=1.7, interpolation="Linear"),
         OGrid(n_levels=4, log2_hashmap_size=6, base_resolution=2, per_level_scale=1.7, interpolation="Smoothstep")]


def _setup(cfg, n=12, seed=0):
    g = torch.Generator().manual_seed(seed)
    table = (torch.rand(n_table_entries(cfg), 2, generator=g) * 2 - 1).half().double()
    x = torch.rand(n, 3, generator=g)
    return g, table, x


@pytest.mark.parametrize("cfg", SMALL)
def test_encode_autograd_is_the_encode_and_its_derivatives_are_right(cfg):
    g, table, x = _setup(cfg)
    assert (encode(x, table.float(), cfg, blend="fp32").double() - encode_autograd(x, table, cfg)).abs().max() < 2e-6
    xd, td = x.double().requires_grad_(True), table.clone().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, b: encode_autograd(a, b, cfg), (xd, td), eps=1e-7, atol=1e-5)
    if cfg.interpolation == "Smoothstep":      # (Linear: the gradient is piecewise constant in each dim, the check is vacuous there)
        assert torch.autograd.gradgradcheck(lambda a, b: encode_autograd(a, b, cfg), (xd, td), eps=1e-7, atol=1e-4)


@pytest.mark.parametrize("cfg", SMALL)
def test_input_grad_closed_forms(cfg):
    """One level, table = linear function of the corner coordinates: trilinear interpolation reproduces
    it, so d feat / d x = scale * slope for Linear and scale * s'(p) * slope for Smoothstep."""
    one = OGrid(n_levels=1, log2_hashmap_size=12, base_resolution=8, per_level_scale=1.5, interpolation=cfg.interpolation)
    res, scale = 8, 7.0                                         # level 0: scale = base - 1, res = 8, dense
    gx, gy, gz = torch.meshgrid(torch.arange(res), torch.arange(res), torch.arange(res), indexing="ij")
    idx = (gx + gy * res + gz * res * res).reshape(-1)
    slope = torch.tensor([0.25, -0.5, 0.125])
    table = torch.zeros(n_table_entries(one), 2, dtype=torch.float64)
    table[idx, 0] = (gx * slope[0] + gy * slope[1] + gz * slope[2]).reshape(-1).double()
    x = torch.rand(64, 3, generator=torch.Generator().manual_seed(1)) * 0.8 + 0.05
    dfeat = torch.zeros(64, 2, dtype=torch.float64)
    dfeat[:, 0] = 1.0
    dx = encode_input_grad(x, table, dfeat, one)
    p = x.double() * scale + 0.5
    p = p - p.floor()
    want = scale * slope.double()[None, :] * (6 * p * (1 - p) if cfg.interpolation == "Smoothstep" else torch.ones_like(p))
    assert (dx - want).abs().max() < 1e-9


class _Stub:
    """CPU stand-ins with the signatures of the perf_b200.ops kernels wrappers, computed by the oracle."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.calls = []

    def params_to_half(self, params, out=None):
        return params.detach().half()

    def hashgrid_fwd(self, half, x, grid):
        self.calls.append("fwd")
        return encode_autograd(x, half.double(), self.cfg).float()       # unrounded: lets the wiring be checked tightly

    @torch.enable_grad()
    def hashgrid_bwd(self, x, dfeat, grid, out=None):
        self.calls.append("bwd_table")
        t = torch.zeros(n_table_entries(self.cfg), 2, dtype=torch.float64, requires_grad=True)
        y = encode_autograd(x, t, self.cfg)
        return torch.autograd.grad(y, t, grad_outputs=dfeat.double())[0].float()

    @torch.enable_grad()
    def hashgrid_bwd_input(self, half, x, dfeat, grid):
        self.calls.append("bwd_input")
        return encode_input_grad(x, half.double(), dfeat, self.cfg).float()

    @torch.enable_grad()
    def hashgrid_bwd_bwd_input(self, half, x, dfeat, ddx, grid, want=(True, True, True)):
        self.calls.append("bwd_bwd_input")
        g, t, dx2 = encode_input_grad_backward(x, half.double(), dfeat, ddx, self.cfg)
        f = lambda v, w: v.float() if (w and v is not None) else (torch.zeros(()) if w else None)
        return f(g, want[0]), f(t, want[1]), (f(dx2, want[2]) if dx2 is not None else (torch.zeros_like(x) if want[2] else None))


class _Sphere(torch.nn.Module):
    """The reference's SphereDistanceField wiring (`pano_joint_predictor.py:22-68`) around a given encode."""

    def __init__(self, enc):
        super().__init__()
        self.enc = enc
        g = torch.Generator().manual_seed(5)
        self.w1 = torch.nn.Parameter(torch.randn(8 + 3, 16, generator=g, dtype=torch.float64) * 0.5)
        self.w2 = torch.nn.Parameter(torch.randn(16, 1, generator=g, dtype=torch.float64) * 0.5)

    def forward(self, directions):
        directions.requires_grad_(True)
        scaled = directions * 0.49 + 0.49
        feat = self.enc(scaled)
        h = torch.nn.functional.softplus(torch.cat([directions.double(), feat.double()], -1) @ self.w1)
        distance = torch.nn.functional.softplus((h @ self.w2)[..., 0] + 1.0)
        grad = torch.autograd.grad(distance, directions, grad_outputs=torch.ones_like(distance), create_graph=True)[0]
        return distance, grad


@pytest.mark.parametrize("cfg", SMALL)
def test_encoding_double_backward_wiring(cfg, monkeypatch):
    """tinycudann.Encoding -> autograd.grad(create_graph=True) -> loss on (distance, grad) -> backward:
    with the kernels replaced by oracle stand-ins the shim must reproduce plain autograd through the
    oracle encode, for the table gradient, the MLP gradients and the returned input gradient."""
    from perf_b200 import ops, shims
    shims.install()
    import tinycudann as tcnn
    stub = _Stub(cfg)
    for name in ("params_to_half", "hashgrid_fwd", "hashgrid_bwd", "hashgrid_bwd_input", "hashgrid_bwd_bwd_input"):
        monkeypatch.setattr(ops, name, getattr(stub, name))
    enc = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": cfg.n_levels, "n_features_per_level": 2,
                            "log2_hashmap_size": cfg.log2_hashmap_size, "base_resolution": cfg.base_resolution,
                            "per_level_scale": cfg.per_level_scale, "interpolation": cfg.interpolation}, dtype=torch.float32)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        enc.params.copy_(((torch.rand(enc.params.shape, generator=g) * 2 - 1)).half().float())
    dirs = torch.nn.functional.normalize(torch.randn(40, 3, generator=g), dim=-1)

    def run(module, params):
        d, grad = module(dirs.clone())
        loss = (d * d).sum() + (grad * grad).sum() * 0.3 + (grad[:, 0] * d).sum()
        loss.backward()
        return d.detach(), grad.detach(), params.grad.detach().clone(), module.w1.grad.clone(), module.w2.grad.clone()

    got = run(_Sphere(enc), enc.params)
    table = enc.params.detach().double().reshape(-1, 2).requires_grad_(True)
    want = run(_Sphere(lambda x: encode_autograd(x, table, cfg)), table)
    assert {"fwd", "bwd_input", "bwd_bwd_input", "bwd_table"} <= set(stub.calls)
    for a, b, name in zip(got, want, ("distance", "grad", "d table", "d w1", "d w2")):
        b = b.reshape(a.shape)
        assert (a.double() - b.double()).abs().max() <= 2e-5 * (1 + b.abs().max()), name


# ------------------------------------------------------------------ the CUDA source's arithmetic, run on the host
import numpy as np                                                              # noqa: E402

SPHERE = OGrid(n_levels=16, log2_hashmap_size=19, base_resolution=16,
               per_level_scale=float(np.exp(np.log(2048 / 16) / 15)), interpolation="Smoothstep")
KERNEL_CFGS = [SPHERE, OGrid(), OGrid(interpolation="Smoothstep"),
               OGrid(n_levels=6, log2_hashmap_size=9, base_resolution=3, per_level_scale=1.9, interpolation="Smoothstep")]


def kernel_inputs(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    table = ((torch.rand(n_table_entries(cfg), 2, generator=g) * 2 - 1) * 0.5).half()
    x = torch.rand(n, 3, generator=g)
    x[:4] = torch.tensor([[0.5, 0.5, 0.5], [0.25, 0.75, 0.125], [1e-3, 0.999, 0.5], [0.98, 0.01, 0.49]])
    dfeat = torch.randn(n, cfg.n_levels * 2, generator=g)
    u = torch.randn(n, 3, generator=g)
    return table, x, dfeat, u


def assert_close(got, want, rel, name):
    want = want.double()
    err = (torch.as_tensor(got).double().cpu() - want).abs().max().item()
    assert err <= rel * (want.abs().max().item() + 1e-30), f"{name}: max|d|={err:.3e} vs scale {want.abs().max().item():.3e}"


def perf_grid(cfg):
    from perf_b200.config import GridConfig
    return GridConfig(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale, cfg.interpolation)


@pytest.mark.parametrize("cfg", KERNEL_CFGS)
def test_cuda_source_bodies_on_host_match_oracle(cfg):
    """perf_b200/csrc/encoding_grad.cu compiled with -DPERF_HOST_HARNESS (tests/host_harness.py): the same
    __host__ __device__ bodies the kernels call, over host arrays, against autograd through the oracle."""
    import host_harness as hh
    table, x, dfeat, u = kernel_inputs(cfg, 512, 21)
    t16 = table.numpy()
    got = hh.bwd_input(perf_grid(cfg), t16, x.numpy(), dfeat.numpy())
    assert_close(got, encode_input_grad(x, table.double(), dfeat, cfg, fp32_positions=True), 1e-5, "dx")
    g_g, g_t, g_x = hh.bwd_bwd_input(perf_grid(cfg), t16, x.numpy(), dfeat.numpy(), u.numpy())
    w_g, w_t, w_x = encode_input_grad_backward(x, table.double(), dfeat, u, cfg, fp32_positions=True)
    assert_close(g_g, w_g, 1e-5, "d dfeat")
    assert_close(g_t, w_t, 1e-5, "d table")
    assert_close(g_x, w_x, 1e-5, "d x")
