"""perf_b200 -- B200-native (sm_100a) implementation of PeRF's per-ray hot path.

Layout:
  csrc/           CUDA kernels + the C-ABI (include/perfb200.h) -> libperfb200.so
  _lib.py         ctypes binding (no fallback: raises if the library is missing)
  config.py       grid / MLP configs mirroring the tcnn config dicts the reference passes
  ops.py          torch-facing wrappers + autograd Functions over the C-ABI
  renderer.py     FusedPanoRenderer: NeRFScene.render / render_dense on the fused megakernel
  shims/          import-compatible ``tinycudann`` / ``nerfacc`` / ``torch_efficient_distloss``
                  modules so the reference's own files run unmodified on this library
"""
from .config import GridConfig, MLPConfig, GEO_MLP, APP_MLP, PERF_GRID  # noqa: F401

__version__ = "0.1.0"
