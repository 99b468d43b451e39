"""Parity gaps named in VERDICT r1 ("What's weak" 1c, 1d), closed on the GPU:
  * perf_b200.field.NGPNeRF.query_density / query_rgb -- the product's mirror of the reference's glue
    (`ngp_nerf.py:136-162`) -- against tests/golden/field.npz, which was produced by running the reference's own
    unmodified NGPNeRF methods (tests/golden/make_golden.py) incl. points outside the box and on its faces;
  * the benchmark's full-size panorama (1024 x 2048, 128 samples/ray) against the oracle on rays strided over ALL
    rows: both poles, the +-pi seam columns, and a translated camera whose far rays leave the box."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def _nerf(golden_field):
    from perf_b200.field import NGPNeRF
    nerf = NGPNeRF(aabb=[-1., -1., -1., 1., 1., 1.]).cuda()
    with torch.no_grad():
        nerf.geo_mlp.params.copy_(golden_field.geo_params)
        nerf.app_mlp.params.copy_(golden_field.app_params)
    return nerf.eval()


def test_field_glue_on_gpu_matches_reference_golden(golden_dir, golden_field):
    g = np.load(os.path.join(golden_dir, "field.npz"))
    x = torch.from_numpy(g["x"])
    nerf = _nerf(golden_field)
    with torch.no_grad():
        sigma = nerf.query_density(x.cuda()).cpu()
        rgb = nerf.query_rgb(x.cuda()).cpu()
    want_s, want_c = torch.from_numpy(g["sigma_mixed"]), torch.from_numpy(g["rgb_mixed"])
    assert sigma.shape == want_s.shape and sigma.dtype == torch.float32
    assert rgb.shape == want_c.shape and rgb.dtype == torch.float16          # ngp_nerf.py:152-162 returns half
    outside = ((x <= -1) | (x >= 1)).any(-1)
    assert outside.sum() > 0
    assert (sigma[outside] == 0).all() and (rgb[outside] == 0).all()          # selector, exactly
    assert (sigma[~outside] > 0).all()
    # density: fp16 logit within one fp16 ulp-ish of the oracle's (4e-3 * max(1,|raw|)), i.e. in log space
    dlog = (sigma[~outside].log() - want_s[~outside].log()).abs()
    raw = want_s[~outside].log().abs().clamp(min=1.0)
    assert (dlog <= 4e-3 * raw + 1e-6).all(), float((dlog / raw).max())
    assert (dlog == 0).float().mean() > 0.7                                    # most logits bit-identical
    assert (rgb.float() - want_c.float()).abs().max() <= 4e-3
    assert ((rgb.float() - want_c.float()) == 0).float().mean() > 0.7


def test_full_size_panorama_strided_rays_match_oracle(golden_field):
    """1024 x 2048 x 128 (BASELINE configs[1] size) against the oracle on 16 rows x 256 columns that cover every
    regime of the panorama: pole rows 0 / 1023 (degenerate longitudes), rows next to them, the equator, the seam
    columns 0 / 2047 (alpha = +-pi), the centre column, and -- camera 0.45 off-centre -- rays whose far samples
    leave the box (selector = 0 tails)."""
    from perf_b200.renderer import FusedPanoRenderer
    H, W, S = 1024, 2048, 128
    r = FusedPanoRenderer.from_params(golden_field.geo_params.cuda(), golden_field.app_params.cuda())
    gen = torch.Generator().manual_seed(77)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=gen))
    pose = torch.eye(4); pose[:3, :3] = q; pose[:3, 3] = torch.tensor([0.45, -0.3, 0.2])
    got = r.render_pano(pose, H, W, S)
    rows = torch.tensor([0, 1, 2, 3, 100, 255, 400, 511, 512, 640, 768, 900, 1020, 1021, 1022, 1023])
    cols = torch.cat([torch.tensor([0, 1, 1023, 1024, 2046, 2047]), torch.randint(0, W, (250,), generator=gen)])
    o, d = oracle.gen_pano_rays(pose, H, W)
    oo, dd = o[rows][:, cols].reshape(-1, 3), d[rows][:, cols].reshape(-1, 3)
    want = oracle.render_rays(golden_field, oo, dd, S, mixed=True)
    pick = lambda t: t.cpu()[rows][:, cols].reshape(oo.shape[0], -1)
    # some of these rays really do leave the box before t = far
    assert float(((oo + dd).abs() > 1).any(-1).float().mean()) > 0.1
    for key, atol in (("rgb", 4e-3), ("distance", 4e-3), ("opacities", 4e-3)):
        err = (pick(got[key]) - want[key]).abs().max()
        assert err <= atol, (key, float(err))
    mse = float(((pick(got["rgb"]) - want["rgb"]) ** 2).mean())
    assert mse == 0 or -10 * np.log10(mse) >= 45.0


@pytest.mark.parametrize("aabb", [(-0.7, -1.3, -0.9, 1.1, 0.8, 1.4),                 # extents 1.8 / 2.1 / 2.3: the 3-FMA division by the extent
                                  (-1.0, -1.0, -1.0, 0.99999988, 1.0, 2.0)],          # extent 1.9999999 (all-ones significand): IEEE division path
                         ids=["generic-box", "all-ones-extent"])
def test_render_in_a_non_unit_box_matches_oracle(golden_field, aabb):
    """Position normalisation (p - lo) / (hi - lo) (`ngp_nerf.py:137-140`) for boxes other than PeRF's [-1,1]^3: the kernels
    divide by the launch-uniform extent with a correctly rounded 3-instruction sequence (render.cu::div_uniform) and fall
    back to the IEEE division for an extent whose significand is all ones; both must reproduce the oracle's division."""
    import dataclasses
    from perf_b200.renderer import FusedPanoRenderer
    field = dataclasses.replace(golden_field, aabb=torch.tensor(aabb))
    r = FusedPanoRenderer.from_params(golden_field.geo_params.cuda(), golden_field.app_params.cuda(), aabb=aabb)
    g = torch.Generator().manual_seed(77)
    R, S = 257, 96
    o = (torch.rand(R, 3, generator=g) - .5) * .4
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    want = oracle.render_rays(field, o, d, S, near=1e-2, far=2.5, mixed=True)
    got = r.render_rays(o.cuda(), d.cuda(), S, near=1e-2, far=2.5)
    np.testing.assert_allclose(got["rgb"].cpu().numpy(), want["rgb"].numpy(), atol=4e-3, rtol=0)
    np.testing.assert_allclose(got["distance"].cpu().numpy(), want["distance"].numpy(), atol=4e-3 * 2.5, rtol=0)
    np.testing.assert_allclose(got["opacities"].cpu().numpy(), want["opacities"].numpy(), atol=4e-3, rtol=0)
    assert float((want["opacities"] < 0.999).float().mean()) > 0.0          # some rays leave the box (selector exercised)
