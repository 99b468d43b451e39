"""CPU tests: the oracle against the golden fixtures minted from the reference's own code
(tests/golden/make_golden.py) and against closed-form / property checks."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.hashgrid import GridConfig, level_table, encode, encode_backward_table, n_table_entries

# SURVEY.md Appendix A (tcnn level table for L16 / base 16 / s=1.4472692012786865 / T=18)
APPENDIX_A = [
    (16, 4096, False, 0), (24, 13824, False, 4096), (34, 39304, False, 17920),
    (49, 117656, False, 57224), (71, 262144, True, 174880), (102, 262144, True, 437024),
    (148, 262144, True, 699168), (213, 262144, True, 961312), (308, 262144, True, 1223456),
    (446, 262144, True, 1485600), (646, 262144, True, 1747744), (934, 262144, True, 2009888),
    (1352, 262144, True, 2272032), (1956, 262144, True, 2534176), (2831, 262144, True, 2796320),
    (4096, 262144, True, 3058464)]


def test_level_table_matches_appendix_a():
    lv = level_table(GridConfig())
    assert [(l.resolution, l.size, l.hashed, l.offset) for l in lv] == APPENDIX_A
    assert n_table_entries(GridConfig()) == 3320608
    assert oracle.field.network_param_count(oracle.field.PERF_GRID, oracle.field.GEO_MLP) == 6644288
    assert oracle.field.network_param_count(oracle.field.PERF_GRID, oracle.field.APP_MLP) == 6648384
    assert abs(float(lv[15].scale) - 4094.9985) < 1e-3 and float(lv[0].scale) == 15.0


def test_proposal_grid_sizes():
    # NGPDensityField (ngp_nerf.py:226-248): L5, T=2^17, max res 128 / 256
    for max_res, n in ((128, 383264), (256, 430080)):
        s = float(np.exp((np.log(max_res) - np.log(16)) / 4))
        assert n_table_entries(GridConfig(n_levels=5, log2_hashmap_size=17, per_level_scale=s)) == n


def test_raygen_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "raygen.npz"))
    for name in ("eye_8x16", "rot_6x10", "rot_128x256", "rot_1024x2048"):
        h, w = g[name + "_hw"]
        o, d = oracle.gen_pano_rays(torch.from_numpy(g[name + "_pose"]), int(h), int(w))
        rows = g[name + "_rows"]
        assert np.array_equal(o[rows].numpy(), g[name + "_o"])
        assert np.array_equal(d[rows].numpy(), g[name + "_d"])
    for name in ("pers75_64", "pers90_33"):
        o, d = oracle.gen_pers_rays(torch.from_numpy(g[name + "_pose"]), float(g[name + "_fov"]), int(g[name + "_res"]))
        assert np.array_equal(o.numpy(), g[name + "_o"]) and np.array_equal(d.numpy(), g[name + "_d"])
    d = oracle.pano_dirs(16, 32)
    assert torch.allclose(d.norm(dim=-1), torch.ones(16, 32), atol=1e-6)
    assert d[0, 0, 2] > 0.99 and d[8, 16, 0] > 0.99        # row 0 looks +z, centre looks +x


@pytest.mark.parametrize("tag", ["mixed", "fp32"])
def test_field_matches_reference_glue(golden_dir, golden_field, tag):
    g = np.load(os.path.join(golden_dir, "field.npz"))
    x = torch.from_numpy(g["x"])
    mixed = tag == "mixed"
    sigma = oracle.query_density(golden_field, x, mixed=mixed)
    rgb = oracle.query_rgb(golden_field, x, mixed=mixed)
    assert np.array_equal(sigma.numpy(), g[f"sigma_{tag}"])
    assert np.array_equal(rgb.numpy(), g[f"rgb_{tag}"])
    outside = ((x <= -1) | (x >= 1)).any(-1)
    assert outside.sum() > 0 and (sigma[outside] == 0).all() and (rgb[outside] == 0).all()
    assert (sigma[~outside] > 0).all()


def test_mixed_vs_fp32_field_gap(golden_dir):
    g = np.load(os.path.join(golden_dir, "field.npz"))
    rel = np.abs(np.log(g["sigma_mixed"][g["sigma_fp32"] > 0]) - np.log(g["sigma_fp32"][g["sigma_fp32"] > 0]))
    assert rel.max() < 5e-2                                    # fp16 rounding points only
    assert np.abs(g["rgb_mixed"] - g["rgb_fp32"]).max() < 5e-3


@pytest.mark.parametrize("tag", ["mixed", "fp32"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_render_matches_reference_glue(golden_dir, golden_field, tag, mode):
    g = np.load(os.path.join(golden_dir, "render.npz"))
    r = oracle.render_rays(golden_field, torch.from_numpy(g["rays_o"]), torch.from_numpy(g["rays_d"]),
                           int(g["n_samples"]), float(g["near"]), float(g["far"]),
                           training=(mode == "train"), jitter=torch.from_numpy(g["jitter"]),
                           bg_noise=torch.from_numpy(g["bg_noise"]), mixed=(tag == "mixed"))
    for k in ("rgb", "distance", "opacities"):
        np.testing.assert_allclose(r[k].numpy(), g[f"{mode}_{tag}_{k}"], rtol=2e-5, atol=2e-6, err_msg=k)
    for k in ("weights", "trans"):
        np.testing.assert_allclose(r[k].reshape(-1).numpy(), g[f"{mode}_{tag}_{k}"], rtol=2e-5, atol=2e-6)


def test_trilinear_partition_of_unity_and_index_range():
    cfg = GridConfig()
    x = torch.rand(512, 3, generator=torch.Generator().manual_seed(3))
    x[:4] = torch.tensor([[0., 0, 0], [1., 1, 1], [0.5, 0.5, 0.5], [1., 0., 0.5]])
    ones = torch.ones(n_table_entries(cfg), 2)
    feat = encode(x, ones, cfg)
    assert torch.allclose(feat, torch.ones_like(feat), atol=1e-6)
    for lvl in level_table(cfg):
        for wt, idx in oracle.hashgrid._corner_weights_indices(x, lvl, False):
            assert (wt >= 0).all() and (wt <= 1).all()
            assert (idx >= lvl.offset).all() and (idx < lvl.offset + lvl.size).all()


def test_dense_level_is_exact_on_vertices():
    """A dense level reproduces its stored value at a grid vertex (x fastest axis)."""
    cfg = GridConfig(n_levels=1)
    lvl = level_table(cfg)[0]
    table = torch.arange(lvl.size * 2, dtype=torch.float32).reshape(-1, 2)
    gx, gy, gz = 3, 5, 7
    # vertex v sits at pos = v  ->  x = (v - 0.5) / scale
    x = torch.tensor([[(gx - .5) / 15., (gy - .5) / 15., (gz - .5) / 15.]])
    feat = encode(x, table, cfg)
    idx = gx + gy * 16 + gz * 256
    assert torch.allclose(feat[0], table[idx], rtol=1e-5)


def test_encode_backward_matches_autograd():
    cfg = GridConfig(n_levels=6, log2_hashmap_size=10)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(200, 3, generator=g)
    table = torch.randn(n_table_entries(cfg), 2, generator=g, requires_grad=True)
    dfeat = torch.randn(200, 12, generator=g)
    (encode(x, table, cfg) * dfeat).sum().backward()
    assert torch.allclose(encode_backward_table(x, dfeat, cfg), table.grad, atol=1e-5)


def test_composite_closed_forms():
    R, S, near, far = 3, 64, 0.01, 1.0
    ts, te = oracle.fixed_samples(R, S, near, far)
    sigma = torch.full((R, S), 2.5)
    out = oracle.composite_fixed(ts, te, sigma, torch.full((R, S, 3), 0.25))
    op = 1 - np.exp(-2.5 * (far - near))
    assert torch.allclose(out['opacities'], torch.full((R, 1), op), atol=1e-5)
    assert torch.allclose(out['rgb'], torch.full((R, 3), 0.25 * op), atol=1e-5)
    assert (out['weights'].sum(-1) <= 1 + 1e-6).all()
    # a single opaque sample -> one-hot weights
    sigma = torch.zeros(1, S); sigma[0, 10] = 1e6
    w, T, a = oracle.render_weight_from_density(ts[:1], te[:1], sigma)
    assert w[0, 10] == 1 and w.sum() == 1 and (T[0, 11:] == 0).all()


def test_packed_equals_dense_and_sampler_edges():
    g = torch.Generator().manual_seed(2)
    R, S = 5, 16
    jit = torch.rand(R, generator=g)
    ts, te = oracle.fixed_samples(R, S, 0.01, 1.0, jit)
    assert torch.equal(ts[:, 1:], te[:, :-1])
    assert torch.allclose(ts[:, 0], 0.01 + jit * (0.99 / S), atol=1e-7)
    sig = torch.rand(R, S, generator=g) * 20
    wd, Td, _ = oracle.render_weight_from_density(ts, te, sig)
    ri = torch.arange(R).repeat_interleave(S)
    wp, Tp, _ = oracle.render_weight_from_density(ts.reshape(-1), te.reshape(-1), sig.reshape(-1), ri)
    assert torch.allclose(wd.reshape(-1), wp, atol=1e-6) and torch.allclose(Td.reshape(-1), Tp, atol=1e-6)
    vals = torch.rand(R * S, 3, generator=g)
    a = oracle.accumulate_along_rays(wp, vals, ri, R)
    b = oracle.accumulate_along_rays(wd, vals.reshape(R, S, 3))
    assert torch.allclose(a, b, atol=1e-6)


def test_distloss_matches_quadratic_definition():
    g = torch.Generator().manual_seed(4)
    R, S = 4, 12
    w = torch.rand(R, S, generator=g) / S
    ts, te = oracle.fixed_samples(R, S, 0.0, 1.0)
    m, iv = (ts + te) / 2, te - ts
    ri = torch.arange(R).repeat_interleave(S)
    fast = oracle.flatten_eff_distloss(w.reshape(-1), m.reshape(-1), iv.reshape(-1), ri)
    brute = ((w[:, :, None] * w[:, None, :] * (m[:, :, None] - m[:, None, :]).abs()).sum()
             + (w * w * iv).sum() / 3) / R
    assert torch.allclose(fast, brute, atol=1e-6)


def test_host_fixed_sampler_equals_oracle():
    """perf_b200.scene.FixedSampleEstimator (host-side torch code feeding the modular path) produces
    the oracle's intervals bit for bit, packed ray-major."""
    from perf_b200.scene import FixedSampleEstimator
    g = torch.Generator().manual_seed(9)
    R, S = 13, 24
    jit = torch.rand(R, generator=g)
    est = FixedSampleEstimator(S, 1e-2, 1.0)
    o = torch.zeros(R, 3)
    for stratified, j in ((False, None), (True, jit)):
        ri, ts, te = est.sampling(o, o, stratified=stratified, jitter=j)
        ts0, te0 = oracle.fixed_samples(R, S, 1e-2, 1.0, j)
        assert torch.equal(ts, ts0.reshape(-1)) and torch.equal(te, te0.reshape(-1))
        assert torch.equal(ri, torch.arange(R).repeat_interleave(S))


def test_config_loader_overrides(tmp_path):
    from perf_b200.config import load_config
    (tmp_path / "device").mkdir()
    (tmp_path / "device" / "local.yaml").write_text("base_exp_dir: /tmp/exp\n")
    (tmp_path / "nerf.yaml").write_text("defaults:\n  - device: local\n  - _self_\nmode: train\nscene:\n  estimator_type: occ\n"
                                        "  train_conf:\n    geo_optimizer:\n      peak_lr: 1e-2\n    pixel_loss_batch_size: 8192\n")
    c = load_config(str(tmp_path), "nerf", ["mode=render_dense", "scene.train_conf.pixel_loss_batch_size=1024", "scene.new.key=0.5"])
    assert c.mode == "render_dense" and c.device.base_exp_dir == "/tmp/exp"
    assert c.scene.train_conf.geo_optimizer.peak_lr == 1e-2 and isinstance(c.scene.train_conf.geo_optimizer.peak_lr, float)
    assert c.scene.train_conf.pixel_loss_batch_size == 1024 and c.scene.new.key == 0.5


def test_c_port_matches_pytorch_oracle(golden_field):
    """oracle/cpath.c (the OpenMP CPU baseline of bench.py) against the PyTorch oracle: same mixed
    precision contract, so the composited pixels agree to fp16-rounding noise."""
    from oracle import cpath
    g = torch.Generator().manual_seed(8)
    R, S = 96, 48
    o = (torch.rand(R, 3, generator=g) - .5) * .3
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    o[:3] = 3.0                                                  # rays outside the box: pure background
    want = oracle.render_rays(golden_field, o, d, S, mixed=True)
    got = cpath.render_rays(golden_field, o, d, S, n_threads=4)
    for k, tol in (("rgb", 2e-3), ("distance", 2e-3), ("opacities", 2e-3)):
        assert (got[k] - want[k]).abs().max() <= tol, (k, float((got[k] - want[k]).abs().max()))
    assert torch.equal(got["rgb"][:3], torch.full((3, 3), 0.5)) and cpath.max_threads() >= 1
