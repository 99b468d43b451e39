"""GPU parity tests of the tensor-core paths: MLP forward (tcgen05) and the fused
encode+MLP network forward, against the mixed-precision oracle; the SIMT debug variant
triangulates layout bugs; autograd of the network against the oracle's autograd."""
import numpy as np
import pytest
import torch

import oracle
from oracle.field import APP_MLP as O_APP, GEO_MLP as O_GEO, PERF_GRID as O_GRID, network_forward
from oracle.hashgrid import encode
from oracle.mlp import flat_param_count, mlp_forward, mlp_hidden, split_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from perf_b200 import ops as _ops
    return _ops


def cfgs():
    from perf_b200.config import APP_MLP, GEO_MLP
    return {"geo": (O_GEO, GEO_MLP), "app": (O_APP, APP_MLP)}


def rand_weights(g, ocfg, scale=1.0):
    ws = []
    for o, i in oracle.mlp.layer_shapes(ocfg):
        lim = scale * (6.0 / (o + i)) ** 0.5
        ws.append(((torch.rand(o * i, generator=g) * 2 - 1) * lim))
    return torch.cat(ws).half()


# tolerance: operands are identical fp16 values; only the fp32 accumulation order differs, which
# can flip a hidden activation by one fp16 ulp (2^-11 relative) now and then.
def assert_close_half(got, want, what, atol=4e-3, frac_exact=0.9):
    got, want = got.float().cpu(), want.float().cpu()
    diff = (got - want).abs()
    assert diff.max() <= atol * max(1.0, float(want.abs().max())), (what, float(diff.max()))
    assert (diff == 0).float().mean() >= frac_exact, (what, float((diff == 0).float().mean()))


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("net", ["geo", "app"])
@pytest.mark.parametrize("N", [1, 127, 128, 129, 5000])
def test_mlp_fwd_matches_oracle(ops, net, simt, N):
    ocfg, pcfg = cfgs()[net]
    g = torch.Generator().manual_seed(21)
    w = rand_weights(g, ocfg)
    feat = (torch.randn(N, 32, generator=g)).half()
    want = mlp_forward(feat.float(), split_params(w.float(), ocfg), ocfg, mixed=True)
    hs = mlp_hidden(feat.float(), split_params(w.float(), ocfg), ocfg)
    out, h1, h2 = ops.mlp_fwd(w.cuda(), feat.cuda(), pcfg, save=True, simt=simt)
    assert_close_half(h1, hs[0], f"{net} h1")
    if ocfg.n_hidden_layers == 2:
        assert_close_half(h2, hs[1], f"{net} h2", frac_exact=0.85)
    assert_close_half(out, want, f"{net} out", frac_exact=0.7)


def test_mlp_tc_equals_simt(ops):
    ocfg, pcfg = cfgs()["app"]
    g = torch.Generator().manual_seed(22)
    w = rand_weights(g, ocfg).cuda()
    feat = torch.randn(100000, 32, generator=g).half().cuda()
    a = ops.mlp_fwd(w, feat, pcfg, simt=False).float()
    b = ops.mlp_fwd(w, feat, pcfg, simt=True).float()
    assert (a - b).abs().max() <= 4e-3
    assert ((a - b) == 0).float().mean() > 0.9


@pytest.mark.parametrize("simt", [True, False], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("net", ["geo", "app"])
def test_network_fwd_matches_oracle(ops, golden_field, net, simt):
    from perf_b200.config import PERF_GRID
    ocfg, pcfg = cfgs()[net]
    params = golden_field.geo_params if net == "geo" else golden_field.app_params
    g = torch.Generator().manual_seed(23)
    x = torch.rand(3001, 3, generator=g)
    x[:4] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [1., 0., 0.25]])
    n_mlp = flat_param_count(ocfg)
    ph = params.half()
    want = network_forward(x, params, O_GRID, ocfg, mixed=True)
    feat_want = encode(x, ph.float()[n_mlp:], O_GRID, blend="half")
    out, feat, h1, h2 = ops.network_fwd(ph.cuda(), x.cuda(), PERF_GRID, pcfg, save=True, simt=simt)
    assert_close_half(feat, feat_want, "feat", atol=1e-3, frac_exact=0.995)
    hs = mlp_hidden(feat_want, split_params(ph.float()[:n_mlp], ocfg), ocfg)
    assert_close_half(h1, hs[0], "h1")
    assert_close_half(out, want, f"{net} out", frac_exact=0.7)
    out2 = ops.network_fwd(ph.cuda(), x.cuda(), PERF_GRID, pcfg, simt=simt)
    assert torch.equal(out2, out)                          # deterministic, save flag changes nothing


@pytest.mark.parametrize("net", ["geo", "app"])
def test_network_autograd_matches_oracle_autograd(ops, golden_field, net):
    """dL/dparams through encode+MLP.  Oracle: torch autograd through the mixed-precision
    restatement (fp16 rounding points are straight-through casts, ReLU masks come from the
    rounded activations) evaluated at the fp16-rounded parameters -- the same operating point
    and the same saved activations the kernels use."""
    from perf_b200.config import PERF_GRID
    ocfg, pcfg = cfgs()[net]
    params = (golden_field.geo_params if net == "geo" else golden_field.app_params).half().float()
    g = torch.Generator().manual_seed(24)
    N = 2000
    x = torch.rand(N, 3, generator=g)
    dout = torch.randn(N, ocfg.n_out, generator=g)
    n_mlp = flat_param_count(ocfg)
    p_ref = params.clone().requires_grad_(True)
    feat = encode(x, p_ref[n_mlp:], O_GRID, blend="half")
    y = mlp_forward(feat, split_params(p_ref[:n_mlp], ocfg), ocfg, mixed=True)
    (y * dout).sum().backward()
    p = params.cuda().requires_grad_(True)
    out = ops.network_apply(p, x.cuda(), PERF_GRID, pcfg)
    (out.float() * dout.cuda()).sum().backward()
    got, want = p.grad.cpu(), p_ref.grad
    assert got.shape == want.shape
    gm, wm = got[:n_mlp], want[:n_mlp]                    # MLP matrices: dense sums over N samples
    assert (gm - wm).abs().max() <= 1e-2 * wm.abs().max(), float((gm - wm).abs().max() / wm.abs().max())
    gg, wg = got[n_mlp:], want[n_mlp:]                    # grid: sparse scatter
    assert ((wg != 0) == (gg != 0)).float().mean() > 0.999
    assert (gg - wg).abs().max() <= 1e-2 * wg.abs().max(), float((gg - wg).abs().max() / wg.abs().max())
    cos = torch.nn.functional.cosine_similarity(gg, wg, dim=0)
    assert cos > 0.9999, float(cos)
