"""The two grid-gradient scatter kernels of the fused training step (perf_b200/csrc/train.cu: per-ray cell
accumulation on the coarse levels, per-row vector atomics on the fine levels) compiled for the host
(tests/host_harness.py) against the oracle's table gradient, for ragged sample counts, jitter, every split of the
rays into pieces and both atomic widths."""
import numpy as np
import pytest
import torch

import host_harness as hh
from oracle.hashgrid import GridConfig as OGrid, encode_backward_table, n_table_entries
from perf_b200.config import GridConfig

AABB = (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0)


def _positions(o, d, jitter, S, near, far):
    """x01 [S*R, 3], rows sample-major, with the renderer's fp32 recipe (nerf_renderer.py:127 on fixed-S intervals)."""
    R = o.shape[0]
    step = (torch.tensor(far) - torch.tensor(near)) / torch.tensor(float(S))
    k = torch.arange(S, dtype=torch.float32)[:, None]
    jit = torch.zeros(R) if jitter is None else jitter
    ts = near + (k + jit[None, :]) * step
    te = near + (k + 1 + jit[None, :]) * step
    tsum = ts + te
    p = o[None, :, :] + (d[None, :, :] * tsum[..., None]) * 0.5
    return ((p - torch.tensor(AABB[:3])) / (torch.tensor(AABB[3:]) - torch.tensor(AABB[:3]))).reshape(-1, 3)


@pytest.mark.parametrize("cfg", [OGrid(), OGrid(n_levels=10, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.7),
                                 OGrid(n_levels=6, log2_hashmap_size=14, interpolation="Smoothstep")],
                         ids=["perf", "ten-levels", "six-smoothstep"])
@pytest.mark.parametrize("S,pieces", [(16, 1), (48, 2), (37, 4)])
def test_host_compiled_scatter_matches_oracle(cfg, S, pieces):
    g = torch.Generator().manual_seed(S)
    R, near, far = 96, 1e-2, 1.0
    o = (torch.rand(R, 3, generator=g) - 0.5) * 0.2
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    d[0] = torch.tensor([1.0, 0.0, 0.0])
    jitter = torch.rand(R, generator=g) if S != 16 else None
    x01 = _positions(o, d, jitter, S, near, far)
    dfeat = torch.randn(S * R, cfg.n_levels * 2, generator=g)
    inside = ((x01 > 0) & (x01 < 1)).all(-1)
    dfeat = dfeat * inside[:, None]                                  # the field's selector zeroes samples outside the box
    dfeat[::7] = 0.0                                                 # rows without gradient are skipped by the kernels
    want = encode_backward_table(x01, dfeat, cfg).numpy()
    pcfg = GridConfig(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale, cfg.interpolation)
    assert pcfg.n_entries == n_table_entries(cfg)
    scale = np.abs(want).max()
    for v4 in (0, 1, 2, 3):                              # bit 0: pair atomics on the fine levels, bit 1: on the coarse flush
        got = hh.hashgrid_bwd_rays(pcfg, AABB, o.numpy(), d.numpy(), None if jitter is None else jitter.numpy(), S, near, far,
                                   dfeat.numpy(), v4=v4, pieces=pieces)
        assert np.abs(got - want).max() <= 2e-5 * scale, (v4, np.abs(got - want).max(), scale)
