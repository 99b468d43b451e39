"""Hash-grid addressing of perf_b200/csrc/common.cuh, compiled for the host (tests/host_harness.py):
the generic `level_corners` against the oracle for arbitrary grid configurations, and the specialised
branch-free `level_corners_fast` against the generic one wherever `fast_addressing_ok` admits it."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import host_harness as hh
from oracle.hashgrid import GridConfig as OGrid, _corner_weights_indices, level_table
from perf_b200.config import GridConfig


def _points(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g)
    x[:8] = torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.5, 0.5], [1., 0., 0.5], [1e-7, 0.999999, 0.25],
                          [0.9999999, 0.9999999, 0.9999999], [0., 1., 0.], [0.33333334, 0.6666667, 1.0]])
    return x


def _oracle(cfg, level, x):
    lvl = level_table(cfg)[level]
    corners = _corner_weights_indices(x, lvl, cfg.interpolation == "Smoothstep")
    idx = torch.stack([c[1] for c in corners], 1).numpy().astype(np.uint32)
    w = torch.stack([c[0] for c in corners], 1).numpy()
    return idx, w


@settings(max_examples=60, deadline=None)
@given(n_levels=st.integers(1, 16), log2_t=st.integers(4, 22), base=st.integers(1, 64),
       scale=st.floats(1.0, 2.5, width=32), smooth=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_generic_addressing_matches_oracle(n_levels, log2_t, base, scale, smooth, seed):
    interp = "Smoothstep" if smooth else "Linear"
    ocfg = OGrid(n_levels=n_levels, log2_hashmap_size=log2_t, base_resolution=base, per_level_scale=float(scale), interpolation=interp)
    lv = level_table(ocfg)
    if lv[-1].offset + lv[-1].size >= 2 ** 31:                 # the library refuses grids this large
        return
    pcfg = GridConfig(n_levels, 2, log2_t, base, float(scale), interp)
    x = _points(64, seed)
    for level in sorted({0, n_levels // 2, n_levels - 1}):
        idx, w, _ = hh.level_corners(pcfg, level, x.numpy(), fast=False)
        o_idx, o_w = _oracle(ocfg, level, x)
        assert np.array_equal(idx, o_idx), (level, "indices")
        assert np.array_equal(w, o_w), (level, "weights")


@settings(max_examples=60, deadline=None)
@given(log2_t=st.integers(12, 22), base=st.integers(4, 32), scale=st.floats(1.25, 2.0, width=32), seed=st.integers(0, 2 ** 16))
def test_fast_addressing_equals_generic_where_admitted(log2_t, base, scale, seed):
    """For every 16-level Linear grid: whenever the host-side gate admits the specialised path for some number
    of leading dense levels, it produces exactly the generic indices and weights on [0,1]^3."""
    ocfg = OGrid(n_levels=16, log2_hashmap_size=log2_t, base_resolution=base, per_level_scale=float(scale))
    lv = level_table(ocfg)
    if lv[-1].offset + lv[-1].size >= 2 ** 31:
        return
    pcfg = GridConfig(16, 2, log2_t, base, float(scale), "Linear")
    x = _points(64, seed).numpy()
    n_dense = sum(1 for l in lv if not l.hashed)
    admitted = hh.level_corners(pcfg, 0, x[:1], fast=True, n_dense=n_dense)[2]
    pow2_hashed = all((l.size & (l.size - 1)) == 0 for l in lv if l.hashed)
    assert admitted == pow2_hashed                              # the gate is exactly "dense prefix + pow2 hashed tail"
    if not admitted:
        return
    for level in range(16):
        a = hh.level_corners(pcfg, level, x, fast=False)
        b = hh.level_corners(pcfg, level, x, fast=True, n_dense=n_dense)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), level
        r = hh.level_corners(pcfg, level, x, fast="rel", n_dense=n_dense)           # level-local indices + offset
        assert np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1]), level


def test_perf_grid_is_admitted_with_four_dense_levels():
    pcfg = GridConfig(16, 2, 18, 16, 1.4472692012786865, "Linear")
    x = _points(256, 1).numpy()
    assert hh.level_corners(pcfg, 0, x[:1], fast=True, n_dense=4)[2]
    assert not hh.level_corners(pcfg, 0, x[:1], fast=True, n_dense=3)[2]
    assert not hh.level_corners(GridConfig(16, 2, 18, 16, 1.4472692012786865, "Smoothstep"), 0, x[:1], fast=True, n_dense=4)[2]
    # the far faces / corner of the box are where a dense level's `% size` wrap triggers: level_corners_rel guards its
    # eight per-corner wraps with one comparison per level
    faces = np.array([[1, 1, 1], [1, .5, 1], [.3, 1, 1], [1, 1, .2], [0, 0, 0], [1, 0, 0], [0, 1, 1], [.999999, .999999, 1]], np.float32)
    x = np.concatenate([x, faces])
    for level in range(16):
        a, b = hh.level_corners(pcfg, level, x, fast=False), hh.level_corners(pcfg, level, x, fast=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        r = hh.level_corners(pcfg, level, x, fast="rel")
        assert np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1]), level
        if level < 4:
            # dense level read through its cell-major copy: cell (gx,gy,gz) -> the pack kernel's corner rule gives the same
            # eight entries as the generic addressing, with the same weights
            lv = level_table(OGrid(n_levels=16, log2_hashmap_size=18, base_resolution=16, per_level_scale=1.4472692012786865))[level]
            cidx, cw, _ = hh.level_corners(pcfg, level, x, fast="cell")
            cell = cidx[:, 0].astype(np.int64)
            res = lv.resolution
            assert cell.max() < res ** 3
            gx, gy, gz = cell % res, (cell // res) % res, cell // (res * res)
            k = np.arange(8)
            e = (gx[:, None] + (k & 1)) + res * ((gy[:, None] + ((k >> 1) & 1)) + res * (gz[:, None] + (k >> 2)))
            assert np.array_equal(e % lv.size + lv.offset, a[0].astype(np.int64)) and np.array_equal(cw, a[1]), level


def test_vector_atomic_scatter_pairs_equal_plain_scatter():
    """scatter8<V4> (x-neighbour pairs in one 16-byte slot -> one vector add) against scatter8<false> and numpy, on
    index patterns covering: aligned pair in both orders, unaligned neighbours (odd, odd+1), equal indices, far apart,
    and real level addressing of the PeRF grid."""
    rng = np.random.default_rng(0)
    n_entries = 4096
    rows = []
    for _ in range(500):
        base = rng.integers(0, n_entries // 2 - 2, size=4) * 2
        kind = rng.integers(0, 5, size=4)
        pairs = []
        for b, kd in zip(base, kind):
            pairs += {0: [b, b + 1], 1: [b + 1, b], 2: [b + 1, b + 2], 3: [b, b], 4: [b, (b + 1000) % n_entries]}[int(kd)]
        rows.append(pairs)
    idx = np.array(rows, np.uint32)
    v = rng.standard_normal((len(rows), 8, 2)).astype(np.float32)
    want = np.zeros((n_entries, 2), np.float64)
    np.add.at(want, idx.reshape(-1), v.reshape(-1, 2).astype(np.float64))
    plain, vec = hh.scatter8(idx, v, n_entries, v4=False), hh.scatter8(idx, v, n_entries, v4=True)
    np.testing.assert_allclose(plain, want, atol=1e-5)
    np.testing.assert_allclose(vec, want, atol=1e-5)
    # real addressing: every level of the PeRF grid
    pcfg = GridConfig(16, 2, 18, 16, 1.4472692012786865, "Linear")
    x = _points(512, 3).numpy()
    for level in (0, 3, 4, 9, 15):
        idx, w, _ = hh.level_corners(pcfg, level, x, fast=False)
        val = np.stack([w, -w], -1).astype(np.float32)
        a, b = hh.scatter8(idx, val, pcfg.n_entries, v4=False), hh.scatter8(idx, val, pcfg.n_entries, v4=True)
        np.testing.assert_allclose(a, b, atol=1e-6)
        assert abs(float(b[:, 0].sum()) - len(x)) < 1e-2          # the 8 weights of a sample sum to 1


def test_uniform_extent_division_is_the_ieee_division():
    """render.cu normalises positions with (p - lo) / (hi - lo) (`ngp_nerf.py:137-140`).  The kernels divide by the launch-
    uniform extent with q = n r, q' = fma(fma(-q, ext, n), r, q), r = RN(1 / ext): bit-equal to the IEEE quotient for every
    admitted extent; an extent whose significand is all ones is not admitted and takes the IEEE division."""
    rng = np.random.default_rng(5)
    for ext in (2.0, 1.0, 1.8, 2.1, 2.3, 0.1, 3.0, 5.123, 1.0000001, 123.456, 7e-3):
        e = np.float32(ext)
        pos = ((rng.random(400_000) * 2 - 1) * 4 * float(e)).astype(np.float32)                     # positions around the box
        bits = (rng.integers(0, 1 << 23, 400_000, dtype=np.uint32) | (rng.integers(100, 156, 400_000, dtype=np.uint32) << 23)
                | (rng.integers(0, 2, 400_000, dtype=np.uint32) << 31)).view(np.float32)              # any normal numerator
        n = np.concatenate([pos, bits, np.float32([0.0, e, -e, np.nextafter(e, np.float32(0)), 1e-30])])
        got, ok = hh.div_uniform(n, float(e))
        assert ok
        assert np.array_equal(got.view(np.uint32), (n / e).astype(np.float32).view(np.uint32)), ext
    allones = np.uint32(0x3FFFFFFF).view(np.float32)                                                 # 1.9999999
    got, ok = hh.div_uniform(np.float32([0.3, -1.7, 1.9999999]), float(allones))
    assert not ok and np.array_equal(got, np.float32([0.3, -1.7, 1.9999999]) / allones)
