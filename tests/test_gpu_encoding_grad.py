"""GPU parity (through the C-ABI) of the encode's input gradient and its double backward
(perf_hashgrid_bwd_input / perf_hashgrid_bwd_bwd_input) against autograd
through the fp64 oracle encode, and of `tinycudann.Encoding` inside the reference's
SphereDistanceField wiring (`pano_joint_predictor.py:22-68`)."""
import numpy as np
import pytest
import torch

from oracle.hashgrid import (GridConfig as OGrid, encode_autograd, encode_input_grad,
                             encode_input_grad_backward, n_table_entries)

pytestmark = pytest.mark.gpu

# SphereDistanceField's grid (`pano_joint_predictor.py:24-41`: L16, T=2^19, 16 -> 2048), the PeRF field
# grid with both interpolations, and a small hash-heavy grid
SPHERE = OGrid(n_levels=16, log2_hashmap_size=19, base_resolution=16,
               per_level_scale=float(np.exp(np.log(2048 / 16) / 15)), interpolation="Smoothstep")
CFGS = [SPHERE, OGrid(), OGrid(interpolation="Smoothstep"),
        OGrid(n_levels=6, log2_hashmap_size=9, base_resolution=3, per_level_scale=1.9, interpolation="Smoothstep")]


def _pg(cfg):
    from perf_b200.config import GridConfig
    return GridConfig(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale, cfg.interpolation)


def _inputs(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    table = ((torch.rand(n_table_entries(cfg), 2, generator=g) * 2 - 1) * 0.5).half()
    x = torch.rand(n, 3, generator=g)
    x[:4] = torch.tensor([[0.5, 0.5, 0.5], [0.25, 0.75, 0.125], [1e-3, 0.999, 0.5], [0.98, 0.01, 0.49]])
    dfeat = torch.randn(n, cfg.n_levels * 2, generator=g)
    u = torch.randn(n, 3, generator=g)
    return table, x, dfeat, u


def _close(got, want, rel, name):
    want = want.double()
    err = (got.double().cpu() - want).abs().max().item()
    assert err <= rel * (want.abs().max().item() + 1e-30), f"{name}: max|d|={err:.3e} vs scale {want.abs().max().item():.3e}"


@pytest.mark.parametrize("cfg", CFGS)
def test_input_grad_matches_oracle(cfg):
    from perf_b200 import ops
    table, x, dfeat, _ = _inputs(cfg, 2048, 21)
    want = encode_input_grad(x, table.double(), dfeat, cfg, fp32_positions=True)
    got = ops.hashgrid_bwd_input(table.cuda(), x.cuda(), dfeat.cuda(), _pg(cfg))
    _close(got, want, 5e-5, "dx")                     # fp32 sums of 16 levels x 2 features, fine levels dominate
    assert ops.hashgrid_bwd_input(table.cuda(), x[:0].cuda(), dfeat[:0].cuda(), _pg(cfg)).shape == (0, 3)


@pytest.mark.parametrize("cfg", CFGS)
def test_input_grad_double_backward_matches_oracle(cfg):
    from perf_b200 import ops
    table, x, dfeat, u = _inputs(cfg, 1024, 22)
    w_g, w_t, w_x = encode_input_grad_backward(x, table.double(), dfeat, u, cfg, fp32_positions=True)
    g_g, g_t, g_x = ops.hashgrid_bwd_bwd_input(table.cuda(), x.cuda(), dfeat.cuda(), u.cuda(), _pg(cfg))
    _close(g_g, w_g, 5e-5, "d dfeat")
    _close(g_t, w_t, 5e-5, "d table")
    _close(g_x, w_x, 5e-5, "d x")
    # outputs are independent: asking for one only gives the same numbers
    only_t = ops.hashgrid_bwd_bwd_input(table.cuda(), x.cuda(), dfeat.cuda(), u.cuda(), _pg(cfg), want=(False, True, False))
    assert only_t[0] is None and only_t[2] is None
    _close(only_t[1], w_t, 5e-5, "d table (alone)")


class _Sphere(torch.nn.Module):
    """The reference's SphereDistanceField wiring around a given encode (softplus MLP in plain torch)."""

    def __init__(self, enc, n_feat, device):
        super().__init__()
        self.enc = enc
        g = torch.Generator().manual_seed(5)
        self.w1 = torch.nn.Parameter((torch.randn(n_feat + 3, 32, generator=g) * 0.3).to(device))
        self.w2 = torch.nn.Parameter((torch.randn(32, 1, generator=g) * 0.3).to(device))

    def forward(self, directions):
        directions.requires_grad_(True)
        feat = self.enc(directions * 0.49 + 0.49)
        h = torch.nn.functional.softplus(torch.cat([directions, feat.to(directions.dtype)], -1) @ self.w1.to(directions.dtype))
        distance = torch.nn.functional.softplus((h @ self.w2.to(directions.dtype))[..., 0] + 1.0)
        grad = torch.autograd.grad(distance, directions, grad_outputs=torch.ones_like(distance), create_graph=True)[0]
        return distance, grad


def test_sphere_distance_field_wiring_matches_oracle():
    """tcnn.Encoding(Smoothstep) -> MLP -> autograd.grad(create_graph=True) -> loss on (distance, grad)
    -> backward, on the GPU through the shim, against the same wiring on the fp64 oracle encode."""
    from perf_b200 import shims
    shims.install()
    import tinycudann as tcnn
    cfg = SPHERE
    enc = tcnn.Encoding(3, {"otype": "HashGrid", "n_levels": 16, "n_features_per_level": 2, "log2_hashmap_size": 19,
                            "base_resolution": 16, "per_level_scale": cfg.per_level_scale, "interpolation": "Smoothstep"}).cuda()
    g = torch.Generator().manual_seed(3)
    init = ((torch.rand(enc.params.shape, generator=g) * 2 - 1) * 0.05).half().float()
    with torch.no_grad():
        enc.params.copy_(init.cuda())
    dirs = torch.nn.functional.normalize(torch.randn(1500, 3, generator=g), dim=-1)

    def run(module, params, d_in):
        d, grad = module(d_in)
        loss = (d * d).sum() * 0.1 + (grad * grad).sum() * 0.03
        loss.backward()
        return d.detach().cpu(), grad.detach().cpu(), params.grad.detach().cpu().clone(), module.w1.grad.cpu().clone()

    got = run(_Sphere(enc, 32, "cuda"), enc.params, dirs.clone().cuda())
    table = init.double().reshape(-1, 2).requires_grad_(True)
    want = run(_Sphere(lambda x: encode_autograd(x, table, cfg, fp32_positions=True), 32, "cpu").double(), table, dirs.clone().double())
    # the shim's features are fp16 (tcnn's output precision): 2^-11 relative on the features
    _close(got[0], want[0], 2e-3, "distance")
    _close(got[1], want[1], 1e-2, "d distance / d directions")
    cos = torch.nn.functional.cosine_similarity(got[2].double().reshape(-1), want[2].reshape(-1), dim=0).item()
    assert cos > 0.999, f"table gradient cosine {cos}"
    _close(got[2], want[2].reshape(-1), 2e-2, "d table")
    _close(got[3], want[3], 2e-2, "d w1")
