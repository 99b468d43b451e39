"""CPU oracle for the PeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, no custom code paths) restatement of the
arithmetic PeRF's renderer performs per ray (plus ``cpath.c``/``cpath.py``: the same eval
path in plain C + OpenMP, cross-checked against the PyTorch restatement in
``tests/test_oracle.py`` and used as the multi-threaded CPU baseline in ``bench.py``):

    equirect ray-gen  ->  fixed-S point sampling  ->  multi-resolution hash-grid
    encode + 64-wide MLP (density field, colour field)  ->  alpha composite

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU baseline.
Nothing under ``perf_b200/`` imports it: the product path fails loudly when the
CUDA library is missing.

PARITY STATUS: **parity unpinned for the third-party arithmetic.**  The hash-grid /
MLP maths (tiny-cuda-nn 1.7), the transmittance scan / accumulate (nerfacc 0.5.3)
and the distortion loss (torch_efficient_distloss 0.1.3) are pip dependencies of the
reference (`/root/reference/requirements.txt:16,34,36`) that are neither vendored nor
installed here; the reference has no tests or golden vectors for them.
Those functions restate the published upstream algorithms
and each one is isolated and named so it can be corrected if a real tcnn output ever
disagrees.  What IS pinned by the reference's own code, imported unmodified from
`/root/reference` when the fixtures in ``tests/golden/`` were generated
(``tests/golden/make_golden.py``): ray generation (`utils/camera_utils.py`), the field
glue (`modules/fields/ngp_nerf.py`: aabb normalise, selector, trunc_exp), the renderer
glue (`modules/scene/nerf_renderer.py`: sample position rule, weights.detach, the
background rules) and the chunking in `modules/scene/nerf.py`.
"""
from .hashgrid import GridConfig, level_table, encode, encode_backward_table  # noqa: F401
from .mlp import MLPConfig, mlp_forward, split_params, flat_param_count          # noqa: F401
from .field import Field, query_density, query_rgb                               # noqa: F401
from .raygen import pano_dirs, gen_pano_rays, gen_pers_rays                                     # noqa: F401
from .sampler import fixed_samples                                               # noqa: F401
from .composite import (render_weight_from_density, accumulate_along_rays,      # noqa: F401
                        composite_fixed, flatten_eff_distloss)
from .render import render_rays, render_pano                                     # noqa: F401
