"""The sort-free batch draw of `perf_b200.scene.RaySupervision` (pool in Morton order + sorted uniform indices from exponential
spacings): same distribution as `torch.randint(0, M, (B,))` followed by a sort (`sup_info.py:253-257` draws B i.i.d. rows with
replacement; the order inside a batch never matters to the step).  CPU test of the host logic; the CUDA kernel that evaluates the
same formula is compared with it in test_gpu_basic.py."""
import numpy as np
import torch

from perf_b200.scene import RaySupervision, Rays


def _pool(h=32, w=64):
    from perf_b200.scene import gen_pano_rays  # noqa: F401  (CUDA-only helper; build the pool by hand on the CPU)
    n = h * w
    g = torch.Generator().manual_seed(0)
    o, d = torch.zeros(n, 3), torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    pool = RaySupervision(Rays(o, d), torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3), torch.rand(n, 1, generator=g))
    pool.morton_sorted = True
    return pool, n


def test_sorted_uniform_indices_have_the_distribution_of_sorted_randint():
    pool, M = _pool()
    B, trials = 512, 400
    hist = np.zeros(16)
    firsts, lasts = [], []
    for _ in range(trials):
        rays, colors, dist, nrm = pool.rand_ray_color_data(B)
        idx = colors[:, 0].long()                                   # the pool's colour channel holds the row index
        assert idx.shape == (B,) and bool((idx[1:] >= idx[:-1]).all()) and int(idx.min()) >= 0 and int(idx.max()) < M
        assert torch.equal(rays.d, pool.all_sup_rays.d[idx]) and torch.equal(dist, pool.all_sup_distances[idx])
        hist += np.bincount((idx.numpy() * 16) // M, minlength=16)
        firsts.append(int(idx[0])); lasts.append(int(idx[-1]))
    # uniform marginal: chi-square over 16 bins with 204 800 draws (99.9 % quantile of chi2_15 = 37.7)
    expect = B * trials / 16
    chi2 = float(((hist - expect) ** 2 / expect).sum())
    assert chi2 < 37.7, chi2
    # order statistics: E[min] = M / (B + 1), E[max] = M B / (B + 1)
    assert abs(np.mean(firsts) - M / (B + 1)) < 0.35 * M / (B + 1) and abs(np.mean(lasts) - M * B / (B + 1)) < 0.35 * M / (B + 1)
    # duplicates occur at the rate of sampling WITH replacement: E[#distinct] = M (1 - (1 - 1/M)^B)
    rays, colors, _, _ = pool.rand_ray_color_data(B)
    distinct = len(set(colors[:, 0].long().tolist()))
    assert abs(distinct - M * (1 - (1 - 1 / M) ** B)) < 25
