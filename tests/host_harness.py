"""TEST HARNESS: compiles perf_b200/csrc/encoding_grad.cu with -DPERF_HOST_HARNESS (plus api_basic.cu for the
level-table builder) into tests/_build/libperf_host_harness.so, a SEPARATE shared object whose two extra
entry points run the kernels' __host__ __device__ bodies over host arrays.  It lets the CPU test-suite check
the arithmetic of the CUDA source against the oracle; the product library (perf_b200/libperfb200.so) is built
without the macro and has no host path."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "perf_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libperf_host_harness.so")
SOURCES = [os.path.join(CSRC, "api_basic.cu"), os.path.join(CSRC, "encoding_grad.cu"), os.path.join(CSRC, "occ.cu"), os.path.join(CSRC, "train.cu"),
           os.path.join(CSRC, "mlp_bwd.cu"),
           os.path.join(HERE, "host_harness.cu")]
_LIB = None


def build() -> str:
    from perf_b200.build import _nvcc
    deps = SOURCES + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "mlp_tc.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        tmp = f"{OUT}.{os.getpid()}.tmp"
        cmd = [_nvcc(), "-DPERF_HOST_HARNESS", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--shared",
               "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-ffp-contract=off"] + SOURCES + ["-o", tmp]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        os.replace(tmp, OUT)
    return OUT


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bwd_input(grid_cfg, table_half: np.ndarray, x01: np.ndarray, dfeat: np.ndarray) -> np.ndarray:
    """grid_cfg: perf_b200.config.GridConfig; table_half [n,2] float16; x01 [N,3] f32; dfeat [N,2L] f32."""
    n = x01.shape[0]
    dx = np.zeros((n, 3), np.float32)
    rc = lib().perf_host_hashgrid_bwd_input(C.byref(grid_cfg.c()), _p(table_half), _p(x01), _p(dfeat), C.c_uint64(n), _p(dx))
    assert rc == 0, rc
    return dx


def bwd_bwd_input(grid_cfg, table_half, x01, dfeat, ddx):
    n = x01.shape[0]
    ddfeat = np.zeros((n, grid_cfg.n_features), np.float32)
    dtable = np.zeros((grid_cfg.n_entries, 2), np.float32)
    dx2 = np.zeros((n, 3), np.float32)
    rc = lib().perf_host_hashgrid_bwd_bwd_input(C.byref(grid_cfg.c()), _p(table_half), _p(x01), _p(dfeat), _p(ddx), C.c_uint64(n),
                                                _p(ddfeat), _p(dtable), _p(dx2))
    assert rc == 0, rc
    return ddfeat, dtable, dx2


def level_corners(grid_cfg, level: int, x01: np.ndarray, fast, n_dense: int = 4):
    """(idx [N,8] uint32, w [N,8] f32, fast_ok) from level_corners (fast=False), level_corners_fast (True) or the
    level-local level_corners_rel of the fused field kernels with the level offset added back (fast="rel"); fast="cell":
    level_cell_dense -- the cell index of a dense level in every column of idx, and the weights."""
    n = x01.shape[0]
    idx, w, ok = np.zeros((n, 8), np.uint32), np.zeros((n, 8), np.float32), C.c_int(0)
    rc = lib().perf_host_level_corners(C.byref(grid_cfg.c()), int(level), {"rel": 2, "cell": 3}.get(fast, int(bool(fast))) if isinstance(fast, str) else int(bool(fast)), C.c_uint32(n_dense), _p(x01), C.c_uint64(n),
                                       _p(idx), _p(w), C.byref(ok))
    assert rc == 0, rc
    return idx, w, bool(ok.value)


def div_uniform(n: np.ndarray, ext: float):
    """(n / ext by common.cuh::div_uniform [N] f32, admitted): the launch-uniform division of the field kernels."""
    n = np.ascontiguousarray(n, np.float32)
    out, ok = np.zeros_like(n), C.c_int(0)
    rc = lib().perf_host_div_uniform(_p(n), C.c_uint64(n.size), C.c_float(ext), _p(out), C.byref(ok))
    assert rc == 0, rc
    return out, bool(ok.value)


def scatter8(idx: np.ndarray, v: np.ndarray, n_entries: int, v4: bool) -> np.ndarray:
    """scatter8<V4> of common.cuh: idx [N,8] uint32, v [N,8,2] f32 -> dtable [n_entries,2] f32."""
    raw = np.zeros(2 * n_entries + 4, np.float32)
    shift = (-raw.ctypes.data % 16) // 4                       # 16-byte aligned view
    dtable = raw[shift:shift + 2 * n_entries].reshape(n_entries, 2)
    rc = lib().perf_host_scatter8(int(v4), _p(np.ascontiguousarray(idx, np.uint32)), _p(np.ascontiguousarray(v, np.float32)),
                                  C.c_uint64(idx.shape[0]), _p(dtable))
    assert rc == 0, rc
    return dtable.copy()


def occ_sample(binaries: np.ndarray, aabb, rays_o: np.ndarray, rays_d: np.ndarray, near: float, far: float, step: float, jitter=None, pieces: int = 1,
               masks: bool = False):
    """occ_march_ray of csrc/occ.cu (count pass, exclusive scan, write pass) -> (ray_indices, t_starts, t_ends).
    ``pieces``: every ray's lattice range cut into that many parts (counts / offsets per (ray, piece))."""
    R = rays_o.shape[0]
    bin8 = np.ascontiguousarray(binaries, np.uint8)
    res3 = (C.c_int * 3)(*binaries.shape)
    a6 = (C.c_float * 6)(*[float(v) for v in aabb])
    o, d = np.ascontiguousarray(rays_o, np.float32), np.ascontiguousarray(rays_d, np.float32)
    j = None if jitter is None else np.ascontiguousarray(jitter, np.float32)
    counts = np.zeros(R * pieces, np.int32)
    f = lib().perf_host_occ_march
    args = (_p(bin8), res3, a6, _p(o), _p(d), _p(j), C.c_uint64(R), C.c_float(near), C.c_float(far), C.c_float(step), C.c_uint32(pieces))
    mk = np.zeros(R * pieces * 4, np.uint32) if masks else None      # count pass records the sample bits, write pass expands them
    assert f(0, *args, _p(counts), None, None, None, None, _p(mk)) == 0
    offsets = np.concatenate([[0], np.cumsum(counts, dtype=np.int64)]).astype(np.int64)
    n = int(offsets[-1])
    ri, ts, te = np.zeros(max(n, 1), np.int64), np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
    assert f(1, *args, None, _p(offsets), _p(ri), _p(ts), _p(te), _p(mk)) == 0
    return ri[:n], ts[:n], te[:n]


def hashgrid_bwd_rays(grid_cfg, aabb, rays_o, rays_d, jitter, n_samples: int, near: float, far: float, dfeat: np.ndarray,
                      v4: int = 1, pieces: int = 1) -> np.ndarray:
    """Both scatter bodies of csrc/train.cu (coarse: per-ray cell accumulation in `pieces` pieces; fine: per row)
    -> d table [n_entries, 2] f32.  `v4` bit 0 / bit 1: 16-byte pair atomics on the fine / coarse levels.
    dfeat [S*R, 2L] f32, rows sample-major (row = k * R + ray)."""
    R = rays_o.shape[0]
    raw = np.zeros(2 * grid_cfg.n_entries + 4, np.float32)
    shift = (-raw.ctypes.data % 16) // 4
    dtable = raw[shift:shift + 2 * grid_cfg.n_entries].reshape(-1, 2)
    a6 = (C.c_float * 6)(*[float(v) for v in aabb])
    o, d = np.ascontiguousarray(rays_o, np.float32), np.ascontiguousarray(rays_d, np.float32)
    j = None if jitter is None else np.ascontiguousarray(jitter, np.float32)
    rc = lib().perf_host_hashgrid_bwd_rays(C.byref(grid_cfg.c()), a6, _p(o), _p(d), _p(j), C.c_uint64(R), C.c_uint32(n_samples),
                                           C.c_float(near), C.c_float(far), _p(np.ascontiguousarray(dfeat, np.float32)), _p(dtable),
                                           int(v4), C.c_uint32(pieces))
    assert rc == 0, rc
    return dtable.copy()


def composite_backward(phase: int, S: int, segments: int, near: float, far: float, jitter, bg_noise, sigma, w, T, rgb_half, dist_acc,
                       seg_trans, g_rgb, g_dist, g_op, g_dl, dist_out, op_out) -> np.ndarray:
    """composite_bwd_ray of csrc/train.cu.  Rows are sample-major (row = k * R + ray); returns d loss / d (raw output):
    [S*R] for the density phase (phase 1), [S*R, 3] for the colour phase (phase 2).  None = NULL pointer."""
    from perf_b200._lib import TrainBuffers
    R = (w.shape[0]) // S
    keep = [None if x is None else np.ascontiguousarray(x) for x in (jitter, bg_noise, sigma, w, T, rgb_half, dist_acc, seg_trans,
                                                                      g_rgb, g_dist, g_op, g_dl, dist_out, op_out)]
    jitter, bg_noise, sigma, w, T, rgb_half, dist_acc, seg_trans, g_rgb, g_dist, g_op, g_dl, dist_out, op_out = keep
    ptr = lambda x: None if x is None else x.ctypes.data
    buf = TrainBuffers(ptr(sigma), ptr(w), ptr(T), ptr(rgb_half), None, None, None, ptr(dist_acc), None, ptr(seg_trans), None)
    out = np.zeros(S * R * (3 if phase == 2 else 1), np.float32)
    rc = lib().perf_host_train_backward_composite(int(phase), C.c_uint32(S), C.c_uint32(segments), C.c_float(near), C.c_float(far), C.c_uint64(R),
                                                  _p(jitter), _p(bg_noise), C.byref(buf), _p(g_rgb), _p(g_dist), _p(g_op), _p(g_dl),
                                                  _p(dist_out), _p(op_out), _p(out))
    assert rc == 0, rc
    return out.reshape(S * R, 3) if phase == 2 else out


def mlp_backward(mlp_cfg, weights_half: np.ndarray, feat: np.ndarray, h1: np.ndarray, h2, dz: np.ndarray):
    """One CTA of csrc/mlp_bwd.cu's CUDA-core twin emulated on the host (perf_host_mlp_bwd).
    weights_half [n_params] f16, feat [N,32] f16, h1 / h2 [N,64] f16 (h2 None for one hidden layer), dz [N,n_out] f32
    -> (d weights [n_params] f32, dfeat [N,32] f32)."""
    N = feat.shape[0]
    c16 = lambda x: None if x is None else np.ascontiguousarray(x, np.float16)
    w, f, a1, a2 = c16(weights_half), c16(feat), c16(h1), c16(h2)
    dz = np.ascontiguousarray(dz, np.float32).reshape(N, -1)
    dW, dfeat = np.zeros(w.shape[0], np.float32), np.zeros((N, 32), np.float32)
    rc = lib().perf_host_mlp_bwd(C.byref(mlp_cfg.c()), _p(w), _p(f), _p(a1), _p(a2), _p(dz), C.c_uint64(N), _p(dW), _p(dfeat))
    assert rc == 0, rc
    return dW, dfeat
