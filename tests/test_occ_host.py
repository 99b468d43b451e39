"""The occupancy-grid interval sampler of perf_b200/csrc/occ.cu compiled for the host (tests/host_harness.py) against
oracle/occ_sampler.py, bit for bit, over random grids, rays (inside / outside the box, axis-parallel, pointing away),
step sizes, near / far planes and jitters -- many more cases than the GPU test can afford."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import host_harness as hh
from oracle.occ_sampler import occ_sample


def _case(seed, res, n_rays, occupancy):
    g = torch.Generator().manual_seed(seed)
    binaries = torch.rand(*res, generator=g) < occupancy
    o = (torch.rand(n_rays, 3, generator=g) * 2 - 1) * 1.3                 # some origins outside [-1,1]^3
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, generator=g), dim=-1)
    d[0] = torch.tensor([1.0, 0.0, 0.0])                                    # axis-parallel: two zero components
    d[1] = torch.tensor([0.0, 0.0, -1.0])
    o[2] = torch.tensor([3.0, 3.0, 3.0]); d[2] = torch.tensor([1.0, 0.0, 0.0])   # never enters the box
    o[3] = torch.zeros(3)
    return g, binaries, o, d


@settings(max_examples=80, deadline=None)
@given(seed=st.integers(0, 2 ** 20), rx=st.integers(1, 24), ry=st.integers(1, 24), rz=st.integers(1, 24),
       step=st.sampled_from([5e-4, 3e-3, 1e-2, 0.0625, 0.3]), near=st.sampled_from([0.0, 1e-2, 0.2]),
       far=st.sampled_from([0.5, 1.5, 4.0]), occupancy=st.sampled_from([0.0, 0.1, 0.6, 1.0]), jittered=st.booleans())
def test_host_compiled_sampler_matches_oracle(seed, rx, ry, rz, step, near, far, occupancy, jittered):
    n_rays = 24 if step < 1e-3 else 64
    g, binaries, o, d = _case(seed, (rx, ry, rz), n_rays, occupancy)
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    jitter = torch.rand(n_rays, generator=g) if jittered else None
    want = occ_sample(binaries, aabb, o, d, near, far, step, jitter)
    got = hh.occ_sample(binaries.numpy(), aabb.tolist(), o.numpy(), d.numpy(), near, far, step, None if jitter is None else jitter.numpy())
    assert np.array_equal(got[0], want[0].numpy()), "ray_indices"
    assert np.array_equal(got[1], want[1].numpy()), "t_starts"
    assert np.array_equal(got[2], want[2].numpy()), "t_ends"
    if occupancy == 0.0:
        assert got[0].size == 0


def test_non_cubic_box_and_full_grid_counts():
    """All cells occupied: every lattice midpoint inside the ray/box overlap is emitted, so the count per ray is the
    overlap length over the step (to within the two end intervals)."""
    g = torch.Generator().manual_seed(5)
    aabb = torch.tensor([-0.5, -1.0, -0.25, 1.5, 0.5, 0.75])
    binaries = torch.ones(7, 5, 3, dtype=torch.bool)
    o = torch.zeros(32, 3) + torch.tensor([0.2, -0.1, 0.1])
    d = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1)
    step, near, far = 0.01, 0.0, 10.0
    ri, ts, te = hh.occ_sample(binaries.numpy(), aabb.tolist(), o.numpy(), d.numpy(), near, far, step)
    want = occ_sample(binaries, aabb, o, d, near, far, step)
    assert np.array_equal(ri, want[0].numpy()) and np.array_equal(ts, want[1].numpy())
    inv = 1.0 / d
    t_exit = torch.minimum(torch.maximum((aabb[:3] - o) * inv, (aabb[3:] - o) * inv).amin(-1), torch.tensor(far))
    counts = np.bincount(ri, minlength=32)
    assert np.all(np.abs(counts - (t_exit / step).numpy()) <= 2)


def test_one_cell_grid_reproduces_the_global_lattice():
    """The lattice decision written in oracle/occ_sampler.py: samples sit on t_k = near + (k + u) * step of ONE
    per-ray lattice.  A single occupied cell in the middle of the box: the first emitted interval starts at a
    lattice point, NOT at the cell entry (a per-run re-phased sampler would start there), and exactly the
    lattice intervals whose midpoint is inside the cell are emitted."""
    res, step, near = 5, 0.03, 0.0
    binaries = torch.zeros(res, res, res, dtype=torch.bool)
    binaries[2, 2, 2] = True                                   # cell [-0.2, 0.2]^3
    aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
    o = torch.tensor([[-0.96, 0.01, -0.02], [-0.96, 0.01, -0.02]])
    d = torch.tensor([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]])
    jitter = torch.tensor([0.0, 0.37])
    t_in, t_out = 0.76, 1.16                                    # the ray crosses the cell for t in [0.76, 1.16]
    for impl in ("oracle", "host"):
        if impl == "oracle":
            ri, ts, te = occ_sample(binaries, aabb, o, d, near, 1.5, step, jitter)
            ri, ts, te = ri.numpy(), ts.numpy(), te.numpy()
        else:
            ri, ts, te = hh.occ_sample(binaries.numpy(), aabb.tolist(), o.numpy(), d.numpy(), near, 1.5, step, jitter.numpy())
        for r in range(2):
            t0 = ts[ri == r]
            k = (t0 - near) / step - float(jitter[r])
            assert np.allclose(k, np.round(k), atol=1e-3), "samples must sit on the global lattice"
            mids = t0 + 0.5 * step
            assert mids.min() >= t_in and mids.max() <= t_out
            assert mids.min() - step < t_in and mids.max() + step > t_out        # nothing inside the cell was skipped
            assert abs(t0[0] - t_in) > 1e-3                                      # not re-phased to the cell entry
            assert np.allclose(np.diff(t0), step, atol=1e-6)


def test_piece_parallel_marching_is_bit_identical():
    """Cutting every ray's lattice walk into pieces marched by different threads (what fills the GPU for an 8192-ray batch)
    and skipping empty cells emit exactly the samples of the ray-serial, point-by-point oracle, in the same order."""
    for seed in range(12):
        g = torch.Generator().manual_seed(100 + seed)
        res = tuple(int(v) for v in torch.randint(1, 40, (3,), generator=g))
        binaries = torch.rand(*res, generator=g) < [0.02, 0.1, 0.6][seed % 3]
        n = 24
        o = (torch.rand(n, 3, generator=g) * 2 - 1) * 1.2
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        d[0] = torch.tensor([0.0, 1.0, 0.0])
        step, near, far = [5e-4, 3e-3, 0.0625][seed % 3], [0.0, 0.2][seed % 2], [1.5, 4.0][seed % 2]
        jit = torch.rand(n, generator=g) if seed % 2 else None
        aabb = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0])
        want = occ_sample(binaries, aabb, o, d, near, far, step, jit)
        for pieces in (1, 2, 7, 32):
            got = hh.occ_sample(binaries.numpy(), aabb.tolist(), o.numpy(), d.numpy(), near, far, step, None if jit is None else jit.numpy(), pieces=pieces)
            assert np.array_equal(got[0], want[0].numpy()) and np.array_equal(got[1], want[1].numpy()) and np.array_equal(got[2], want[2].numpy()), (seed, pieces)
        # single march: the count pass records sample bits, the write pass expands them (needs <= 128 lattice points per piece)
        pieces = int(np.ceil(((far - near) / step + 17) / 127))
        if pieces <= 1024:
            got = hh.occ_sample(binaries.numpy(), aabb.tolist(), o.numpy(), d.numpy(), near, far, step, None if jit is None else jit.numpy(), pieces=pieces, masks=True)
            assert np.array_equal(got[0], want[0].numpy()) and np.array_equal(got[1], want[1].numpy()) and np.array_equal(got[2], want[2].numpy()), (seed, "masks")
