"""GPU parity tests (through the C-ABI) of the non-tensor-core kernels: ray-gen, hash-grid
encode forward/backward, packed composite kernels, Adam.  Oracle = oracle/ on CPU."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.hashgrid import GridConfig as OGrid, encode, encode_backward_table, n_table_entries

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from perf_b200 import ops as _ops
    return _ops


def edge_points(g, n):
    x = torch.rand(n, 3, generator=g)
    x[:6] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [1., 0., 0.5], [1e-7, 0.999999, 0.25],
                          [0.9999999, 0.9999999, 0.9999999]])
    x[6:12] = torch.rand(6, 3, generator=g) * 1.1 - 0.05        # slightly outside the box
    # points exactly on cell boundaries of level 0 (scale 15): x = (v - 0.5) / 15
    x[12:20, 0] = (torch.arange(8, dtype=torch.float32) + 1 - 0.5) / 15.0
    return x


def test_raygen_matches_reference_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "raygen.npz"))
    for name in ("eye_8x16", "rot_6x10", "rot_128x256", "rot_1024x2048"):
        h, w = (int(v) for v in g[name + "_hw"])
        o, d = ops.raygen_pano(g[name + "_pose"], h, w)
        rows = g[name + "_rows"]
        np.testing.assert_array_equal(o.cpu().numpy()[rows], g[name + "_o"])
        np.testing.assert_allclose(d.cpu().numpy()[rows], g[name + "_d"], atol=2e-6, rtol=0, err_msg=name)
    for name in ("pers75_64", "pers90_33"):                      # perspective cameras (render_dense cam_type != pano)
        o, d = ops.raygen_pers(g[name + "_pose"], float(g[name + "_fov"]), int(g[name + "_res"]))
        np.testing.assert_array_equal(o.cpu().numpy(), g[name + "_o"])
        np.testing.assert_allclose(d.cpu().numpy(), g[name + "_d"], atol=2e-6, rtol=0, err_msg=name)
    # row window == slice of the full image, bit for bit
    pose = g["rot_128x256_pose"]
    o_full, d_full = ops.raygen_pano(pose, 128, 256)
    o_win, d_win = ops.raygen_pano(pose, 128, 256, row0=40, rows=17)
    assert torch.equal(d_win, d_full[40:57]) and torch.equal(o_win, o_full[40:57])


@pytest.mark.parametrize("cfg", [OGrid(), OGrid(n_levels=5, log2_hashmap_size=17, per_level_scale=1.6817928305074292),
                                 OGrid(n_levels=8, log2_hashmap_size=12, interpolation="Smoothstep")])
def test_hashgrid_fwd_matches_oracle(ops, cfg):
    from perf_b200.config import GridConfig
    g = torch.Generator().manual_seed(11)
    n = n_table_entries(cfg)
    table = ((torch.rand(n, 2, generator=g) * 2 - 1) * 0.5).half()
    x = edge_points(g, 4096)
    want = encode(x, table.float(), cfg, blend="half")
    pg = GridConfig(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale, cfg.interpolation)
    got = ops.hashgrid_fwd(table.cuda(), x.cuda(), pg).float().cpu()
    diff = (got - want).abs()
    # tcnn's fp16 fma chain, same corner order on both sides: identical to the bit
    assert (diff == 0).float().mean() > 0.995, float((diff == 0).float().mean())
    assert diff.max() <= 1e-3 * max(1.0, float(want.abs().max()))


def test_hashgrid_bwd_matches_oracle(ops):
    from perf_b200.config import GridConfig
    cfg = OGrid()
    g = torch.Generator().manual_seed(12)
    x = edge_points(g, 3000)
    dfeat = torch.randn(3000, 32, generator=g)
    dfeat[:50] = 0                                          # zero rows are skipped by the kernel
    want = encode_backward_table(x, dfeat, cfg)
    got = ops.hashgrid_bwd(x.cuda(), dfeat.cuda(), GridConfig()).cpu()
    assert got.shape == want.shape
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    # accumulate into an existing buffer
    acc = ops.hashgrid_bwd(x.cuda(), dfeat.cuda(), GridConfig(), out=torch.from_numpy(want.numpy()).cuda().clone())
    np.testing.assert_allclose(acc.cpu().numpy(), 2 * want.numpy(), rtol=1e-4, atol=2e-5)


def ragged_rays(g, n_rays, max_len):
    lens = torch.randint(0, max_len, (n_rays,), generator=g)
    lens[3] = 0; lens[n_rays - 1] = 0; lens[5] = 1; lens[7] = 33; lens[8] = 32; lens[9] = 65
    ri = torch.repeat_interleave(torch.arange(n_rays), lens)
    N = int(lens.sum())
    ts = torch.rand(N, generator=g)
    te = ts + torch.rand(N, generator=g) * 0.05
    sig = torch.rand(N, generator=g) * 30
    sig[::17] = 0
    return ri, ts, te, sig


def test_packed_composite_matches_oracle(ops):
    g = torch.Generator().manual_seed(13)
    n_rays = 200
    ri, ts, te, sig = ragged_rays(g, n_rays, 150)
    w0, T0, a0 = oracle.render_weight_from_density(ts, te, sig, ri)
    w, T, a = ops.weights_from_density(ts.cuda(), te.cuda(), sig.cuda(), ri.cuda(), n_rays)
    np.testing.assert_allclose(w.cpu().numpy(), w0.numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(T.cpu().numpy(), T0.numpy(), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(a.cpu().numpy(), a0.numpy(), rtol=2e-5, atol=1e-7)
    vals = torch.rand(ri.numel(), 3, generator=g)
    acc0 = oracle.accumulate_along_rays(w0, vals, ri, n_rays)
    acc = ops.accumulate_along_rays(w, vals.cuda(), ri.cuda(), n_rays)
    np.testing.assert_allclose(acc.cpu().numpy(), acc0.numpy(), rtol=2e-5, atol=1e-6)
    op0 = oracle.accumulate_along_rays(w0, None, ri, n_rays)
    op = ops.accumulate_along_rays(w, None, ri.cuda(), n_rays)
    np.testing.assert_allclose(op.cpu().numpy(), op0.numpy(), rtol=2e-5, atol=1e-6)
    assert float(op.max()) <= 1 + 1e-5 and float(op[3]) == 0.0


def test_packed_composite_backward_matches_autograd(ops):
    g = torch.Generator().manual_seed(14)
    n_rays = 64
    ri, ts, te, sig = ragged_rays(g, n_rays, 90)
    sig = sig.clone().requires_grad_(True)
    w0, T0, _ = oracle.render_weight_from_density(ts, te, sig, ri)
    gw, gT = torch.randn(w0.shape, generator=g), torch.randn(w0.shape, generator=g)
    ((w0 * gw).sum() + (T0 * gT).sum()).backward()
    w, T, _ = ops.weights_from_density(ts.cuda(), te.cuda(), sig.detach().cuda(), ri.cuda(), n_rays)
    gs = ops.weights_from_density_bwd(ts.cuda(), te.cuda(), sig.detach().cuda(), ri.cuda(), n_rays, w, T, gw.cuda(), gT.cuda())
    np.testing.assert_allclose(gs.cpu().numpy(), sig.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_adam_matches_torch(ops):
    g = torch.Generator().manual_seed(15)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.0)
    p = p0.clone().cuda()
    m, v, ph = torch.zeros_like(p), torch.zeros_like(p), torch.empty(n, dtype=torch.float16, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * 128
        lr = 1e-2 * step / 5
        for grp in opt.param_groups:
            grp["lr"] = lr
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad.cuda(), m, v, step, lr, params_half=ph)
    np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert torch.equal(ph.cpu(), p.cpu().half())


def test_adam_unaligned_views_and_batch_gather(ops):
    """The vectorised Adam kernel falls back to one element per thread on views that are not 16-byte aligned (a shard of
    an odd-length parameter); perf_gather_rows == indexing each array."""
    g = torch.Generator().manual_seed(151)
    n = 4099
    base = [torch.randn(n + 1, generator=g).cuda() for _ in range(4)]
    base[3].abs_()
    p, gr, m, v = (t[1:] for t in base)                            # +4 bytes: misaligned for float4
    ref = [t.clone() for t in (p, gr, m, v)]
    ph = torch.empty(n + 1, dtype=torch.float16, device="cuda")[1:]
    ops.adam_step(p, gr, m, v, 3, 1e-2, params_half=ph)
    pa, ga, ma, va = (t.clone().contiguous() for t in ref)         # aligned copies through the vector path
    pha = torch.empty(n, dtype=torch.float16, device="cuda")
    ops.adam_step(pa, ga, ma, va, 3, 1e-2, params_half=pha)
    assert torch.equal(p, pa) and torch.equal(m, ma) and torch.equal(v, va) and torch.equal(ph, pha)
    M, B = 5000, 777
    arrays = [torch.randn(M, w, generator=g).cuda() for w in (3, 3, 3, 1, 3)]
    idx = torch.randint(0, M, (B,), generator=g).cuda()
    got = ops.gather_rows(idx, *arrays)
    for a, o in zip(arrays, got):
        assert torch.equal(o, a[idx])


def test_draw_gather_rows_matches_the_host_formula(ops):
    """perf_draw_gather_rows: index b = floor(csum[b] / csum[B] * M) (sorted uniform draw, tests/test_batch_draw.py) + gather."""
    g = torch.Generator().manual_seed(152)
    M, B = 524288, 8192
    csum = torch.cumsum(torch.empty(B + 1, dtype=torch.float64).exponential_(generator=g), 0).cuda()
    arrays = [torch.randn(M, w, generator=g).cuda() for w in (3, 3, 1)]
    *got, idx = ops.draw_gather_rows(csum, M, *arrays, want_idx=True)
    want = (csum[:-1] / csum[-1] * M).to(torch.int64).clamp_(0, M - 1)
    assert torch.equal(idx, want) and bool((idx[1:] >= idx[:-1]).all())
    for a, o in zip(arrays, got):
        assert torch.equal(o, a[idx])


def test_params_to_half_and_pack(ops):
    from perf_b200.config import APP_MLP, GEO_MLP, PERF_GRID
    g = torch.Generator().manual_seed(16)
    n_e = PERF_GRID.n_entries
    geo = torch.randn(GEO_MLP.n_params + 2 * n_e, generator=g)
    app = torch.randn(APP_MLP.n_params + 2 * n_e, generator=g)
    gh, ah = ops.params_to_half(geo.cuda()), ops.params_to_half(app.cuda())
    assert torch.equal(gh.cpu(), geo.half()) and torch.equal(ah.cpu(), app.half())
    packed = ops.pack_tables(gh, ah).cpu()
    gt, at = geo.half()[GEO_MLP.n_params:].view(-1, 2), app.half()[APP_MLP.n_params:].view(-1, 2)
    assert torch.equal(packed[:n_e, :2], gt) and torch.equal(packed[:n_e, 2:], at)
    # cell-major copy of the four dense levels behind the entries: cell (gx,gy,gz), corner k -> entry
    # ((gx+kx) + res (gy+ky) + res^2 (gz+kz)) % size of that level (include/perfb200.h::perf_pack_tables)
    from oracle.hashgrid import GridConfig as OGrid, level_table
    lv = level_table(OGrid(n_levels=16, log2_hashmap_size=18, base_resolution=16, per_level_scale=PERF_GRID.per_level_scale))
    start = n_e
    for l in range(4):
        assert not lv[l].hashed
        res, size, off = lv[l].resolution, lv[l].size, lv[l].offset
        c = torch.arange(res ** 3)
        gx, gy, gz = c % res, (c // res) % res, c // (res * res)
        k = torch.arange(8)
        e = (gx[:, None] + (k & 1)) + res * ((gy[:, None] + ((k >> 1) & 1)) + res * (gz[:, None] + (k >> 2)))
        e = (e % size + off).reshape(-1)
        blk = packed[start:start + 8 * res ** 3]
        assert torch.equal(blk[:, :2], gt[e]) and torch.equal(blk[:, 2:], at[e]), l
        start += 8 * res ** 3
    assert start == packed.shape[0] == ops.packed_table_entries()


def test_hashgrid_bwd_rays_matches_oracle(ops):
    """Ray-ordered grid scatter (per-ray cell accumulation for coarse levels, direct atomics for the
    fine ones) == the oracle's scatter on the same positions."""
    g = torch.Generator().manual_seed(17)
    R, S, near, far = 333, 40, 1e-2, 1.0
    o = (torch.rand(R, 3, generator=g) - .5) * .2
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    jit = torch.rand(R, generator=g)
    dfeat = torch.randn(S * R, 32, generator=g)
    dfeat[::7] = 0
    ts, te = oracle.fixed_samples(R, S, near, far, jit)                       # [R,S]
    pos = o[:, None, :] + d[:, None, :] * (ts + te)[..., None] / 2.0
    x01 = ((pos + 1.0) / 2.0).permute(1, 0, 2).reshape(-1, 3)                 # sample-major rows
    want = encode_backward_table(x01, dfeat, OGrid())
    got = ops.hashgrid_bwd_rays(o.cuda(), d.cuda(), jit.cuda(), S, near, far, dfeat.cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-4, atol=2e-5)


def test_occ_sampler_matches_oracle(ops):
    """Occupancy-grid marcher == oracle/occ_sampler.py, bit for bit, incl. rays that
    miss the box, start inside/outside, and a stratified offset per ray."""
    from oracle.occ_sampler import occ_sample
    g = torch.Generator().manual_seed(18)
    res = 32
    binaries = torch.rand(res, res, res, generator=g) < 0.15
    aabb = torch.tensor([-1., -1., -1., 1., 1., 1.])
    R = 257
    o = (torch.rand(R, 3, generator=g) - .5) * 1.0
    o[:20] *= 5.0                                               # origins outside the box
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d[5] = torch.tensor([0., 0., 1.])                           # axis-aligned (zero components)
    for jit in (None, torch.rand(R, generator=g)):
        ri0, ts0, te0 = occ_sample(binaries, aabb, o, d, 0.0, 1.5, 5e-3, jit)
        ri, ts, te = ops.occ_sample(binaries.cuda(), aabb.tolist(), o.cuda(), d.cuda(), 0.0, 1.5, 5e-3,
                                    None if jit is None else jit.cuda())
        assert ri.numel() == ri0.numel() and ri0.numel() > 1000
        assert torch.equal(ri.cpu(), ri0)
        assert torch.equal(ts.cpu(), ts0) and torch.allclose(te.cpu(), te0, atol=1e-7)


def test_occ_estimator_shim_end_to_end(ops, golden_field):
    """The reference's renderer call pattern on the occupancy estimator shim: sampling with a sigma_fn
    (visibility culling), packed weights, accumulate -- against the oracle on the same samples."""
    from perf_b200.shims.nerfacc.estimators.occ_grid import OccGridEstimator
    from perf_b200.field import NGPNeRF
    nerf = NGPNeRF(aabb=[-1., -1., -1., 1., 1., 1.]).cuda()
    with torch.no_grad():
        nerf.geo_mlp.params.copy_(golden_field.geo_params); nerf.app_mlp.params.copy_(golden_field.app_params)
    nerf.eval()
    est = OccGridEstimator(roi_aabb=torch.tensor([-1., -1., -1., 1., 1., 1.]), resolution=32, levels=1).cuda()
    est.train()
    g = torch.Generator().manual_seed(19)
    occ_cpu = (torch.rand(32 ** 3, generator=g) < 0.2).float()
    est.update_every_n_steps(step=0, occ_eval_fn=lambda x: occ_cpu.cuda()[est._cell_index(x)], occ_thre=1e-2, ema_decay=0.1, warmup_steps=4, n=1)
    assert 0.15 < float(est.binaries.float().mean()) < 0.25
    est.eval()
    R = 64
    o = ((torch.rand(R, 3, generator=g) - .5) * .3).cuda()
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).cuda()

    def sigma_fn(ts, te, ri):
        return nerf.query_density(o[ri] + d[ri] * (ts + te)[:, None] / 2.0).squeeze(-1)
    ri, ts, te = est.sampling(o, d, sigma_fn=sigma_fn, near_plane=0., far_plane=1.5, render_step_size=5e-3, stratified=False,
                              cone_angle=0., alpha_thre=0.)
    assert ri.numel() > 0 and bool((ri[1:] >= ri[:-1]).all())
    with torch.no_grad():
        sig = sigma_fn(ts, te, ri)
    w0, T0, _ = oracle.render_weight_from_density(ts.cpu(), te.cpu(), sig.cpu(), ri.cpu())
    assert float(T0.min()) >= 1e-4 * 0.999                     # culled below early_stop_eps
    from perf_b200.shims import nerfacc
    w, T, _ = nerfacc.render_weight_from_density(ts, te, sig, ray_indices=ri, n_rays=R)
    np.testing.assert_allclose(w.cpu().numpy(), w0.numpy(), rtol=1e-4, atol=1e-6)


def test_shims_under_cuda_default_tensor_type(golden_field):
    """The reference's main() makes CUDA the default tensor type (core_exp_runner.py:266); the plugin
    modules must construct and run in that regime (parameters land on the GPU like tcnn's)."""
    import perf_b200.shims.tinycudann as tcnn
    from perf_b200.field import ENCODING_CONFIG, GEO_NETWORK_CONFIG
    from perf_b200.shims.nerfacc.estimators.occ_grid import OccGridEstimator
    torch.set_default_device("cuda")
    try:
        net = tcnn.NetworkWithInputEncoding(3, 1, ENCODING_CONFIG, GEO_NETWORK_CONFIG)
        assert net.params.is_cuda and net.params.dtype == torch.float32
        y = net(torch.rand(100, 3))
        assert y.is_cuda and y.dtype == torch.float16 and y.shape == (100, 1)
        est = OccGridEstimator(roi_aabb=torch.tensor([-1., -1., -1., 1., 1., 1.]), resolution=16, levels=1)
        assert est.binaries.is_cuda
        est.binaries |= True
        ri, ts, te = est.sampling(torch.zeros(4, 3), torch.nn.functional.normalize(torch.rand(4, 3), dim=-1), near_plane=0., far_plane=1.5,
                                  render_step_size=0.05, stratified=True)
        assert ri.numel() > 0
    finally:
        torch.set_default_device("cpu")
