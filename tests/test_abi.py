"""CPU tests of the C-ABI boundary: the library builds/loads, exports exactly the symbols
include/perfb200.h declares, and its host-side logic (level table, parameter layout, argument
validation) agrees with the oracle.  No kernels are launched (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from oracle.hashgrid import GridConfig as OGrid, level_table

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from perf_b200 import _lib
    return _lib.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "perfb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(perf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from perf_b200 import _lib
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in perfb200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.perf_abi_version() == 1


@pytest.mark.parametrize("cfg", [OGrid(), OGrid(n_levels=5, log2_hashmap_size=17, per_level_scale=float(np.exp((np.log(128) - np.log(16)) / 4))),
                                 OGrid(n_levels=16, log2_hashmap_size=19, per_level_scale=1.3819)])
def test_level_table_matches_oracle(cfg):
    from perf_b200.config import GridConfig
    lv, n = GridConfig(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale).levels()
    want = level_table(cfg)
    assert n == oracle.hashgrid.n_table_entries(cfg)
    for a, b in zip(lv, want):
        assert np.float32(a.scale) == b.scale          # bit-identical fp32 scale
        assert (a.resolution, a.size, a.offset, bool(a.hashed)) == (b.resolution, b.size, b.offset, b.hashed)


def test_param_counts_match_checkpoint_layout():
    # nerf.geo_mlp.params / nerf.app_mlp.params lengths of a PeRF checkpoint
    from perf_b200.config import APP_MLP, GEO_MLP, PERF_GRID, network_param_count
    assert network_param_count(PERF_GRID, GEO_MLP) == 6644288 == 3072 + 6641216
    assert network_param_count(PERF_GRID, APP_MLP) == 6648384 == 7168 + 6641216
    assert GEO_MLP.n_params == 3072 and APP_MLP.n_params == 7168


def test_bad_arguments_fail_loudly(lib):
    from perf_b200 import _lib
    from perf_b200.config import GridConfig, MLPConfig
    n = _lib.u64(0)
    bad = GridConfig(n_levels=17).c()
    assert lib.perf_grid_describe(bad, None, n) == -2 and b"n_levels" in lib.perf_last_error()
    bad = GridConfig(n_features_per_level=4).c()
    assert lib.perf_grid_describe(bad, None, n) == -2
    assert lib.perf_network_param_count(GridConfig().c(), MLPConfig(n_neurons=128).c(), n) == -2
    assert lib.perf_network_param_count(GridConfig().c(), MLPConfig(n_hidden_layers=3).c(), n) == -2
    with pytest.raises(_lib.PerfError):
        _lib.check(lib.perf_params_to_half(None, None, 10, None))
    with pytest.raises(ValueError):
        GridConfig.from_dict({"otype": "Frequency"})
    with pytest.raises(ValueError):
        MLPConfig.from_dict({"otype": "FullyFusedMLP", "activation": "Tanh"}, 32, 1)


def test_no_cpu_path():
    """The product refuses CPU tensors instead of silently computing something else."""
    import torch
    from perf_b200 import ops
    from perf_b200.renderer import FusedPanoRenderer
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.params_to_half(torch.zeros(16))
    with pytest.raises(RuntimeError):
        FusedPanoRenderer.from_params(torch.zeros(6644288), torch.zeros(6648384))


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import perf_b200, perf_b200.ops, perf_b200.renderer; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "perf_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "import oracle" not in open(os.path.join(dirpath, f)).read(), f


def test_level_table_matches_oracle_property():
    """Randomised configurations (hypothesis): the library's host-side level table (fp32 scale bits,
    resolution, size, offset, hashed flag) equals the oracle's restatement of tcnn's constructor."""
    from hypothesis import given, settings, strategies as st
    from perf_b200.config import GridConfig

    @settings(max_examples=300, deadline=None)
    @given(n_levels=st.integers(1, 16), log2t=st.integers(8, 22), base=st.sampled_from([4, 8, 16, 32]),
           scale=st.floats(1.0625, 2.0, allow_nan=False, width=32))
    def check(n_levels, log2t, base, scale):
        cfg = OGrid(n_levels=n_levels, log2_hashmap_size=log2t, base_resolution=base, per_level_scale=float(np.float32(scale)))
        want = level_table(cfg)
        if want[-1].offset + want[-1].size >= 2 ** 31:
            return
        lv, n = GridConfig(n_levels, 2, log2t, base, float(np.float32(scale))).levels()
        assert n == want[-1].offset + want[-1].size
        for a, b in zip(lv, want):
            assert np.float32(a.scale) == b.scale
            assert (a.resolution, a.size, a.offset, bool(a.hashed)) == (b.resolution, b.size, b.offset, b.hashed)
    check()


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """The boundary is a C ABI: include/perfb200.h compiles as strict C99 and examples/render_pano_host.c (a host
    that uses nothing but the header and the CUDA runtime) compiles and links against libperfb200.so."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "perfb200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    cuda = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("no CUDA toolkit headers")
    from perf_b200.build import build
    lib_dir = os.path.dirname(build())
    exe = str(tmp_path / "render_pano_host")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                    os.path.join(root, "examples", "render_pano_host.c"), "-L", lib_dir, "-lperfb200",
                    "-L", os.path.join(cuda, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    usage = subprocess.run([exe], capture_output=True, text=True)
    assert usage.returncode == 2 and "usage" in usage.stderr
