"""Generate the golden fixtures in this directory.  Run HERE (the build container), never
on the GPU box: it imports the reference, which only exists at /root/reference.

    python tests/golden/make_golden.py

What is the reference's own code and what is restated:

* ``raygen.npz``   -- `utils/camera_utils.py` (``gen_pano_rays``, imported unmodified; only
  ``trimesh.creation.icosphere`` is stubbed because trimesh is not installed and the
  symbol is unused on this path).
* ``field.npz``    -- `modules/fields/ngp_nerf.py` ``NGPNeRF.query_density/query_rgb``
  imported unmodified, running on a CPU stand-in for the un-vendored ``tinycudann``
  module whose arithmetic is ``oracle.network_forward``.
* ``render.npz``   -- `modules/scene/nerf_renderer.py` ``NeRFOCCRenderer.render`` imported
  unmodified, on CPU stand-ins for ``nerfacc`` (``oracle.composite``) and an estimator
  whose ``sampling`` returns the fixed-S intervals of ``oracle.sampler``.

So the glue (aabb normalise, selector, trunc_exp, sample-position rule, weights.detach,
background rules, dtype promotions) is pinned by the reference itself; the third-party
arithmetic is pinned only to our restatement ("parity unpinned", see oracle/__init__.py).
The field itself is NOT stored (2 x 6.6 M params): it is regenerated from
``oracle.Field.random(seed, grid_scale)`` -- torch's CPU generator is stable.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle.field import APP_MLP, GEO_MLP, PERF_GRID  # noqa: E402
from oracle.hashgrid import GridConfig  # noqa: E402
from oracle.mlp import MLPConfig  # noqa: E402

FIELD_SEED, FIELD_GRID_SCALE = 1337, 0.5


def install_stand_ins(mixed: bool):
    """CPU stand-ins for the reference's absent third-party imports."""
    trimesh = types.ModuleType("trimesh"); creation = types.ModuleType("trimesh.creation")
    creation.icosphere = lambda *a, **k: None
    trimesh.creation = creation
    sys.modules.update({"trimesh": trimesh, "trimesh.creation": creation})

    tcnn = types.ModuleType("tinycudann")

    class NetworkWithInputEncoding(torch.nn.Module):
        def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
            super().__init__()
            self.grid = GridConfig.from_dict(encoding_config)
            self.mlp = MLPConfig.from_dict(network_config, self.grid.n_levels * self.grid.n_features_per_level,
                                           n_output_dims)
            n = oracle.field.network_param_count(self.grid, self.mlp)
            self.params = torch.nn.Parameter(torch.zeros(n))
            self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims

        def forward(self, x):
            y = oracle.field.network_forward(x, self.params, self.grid, self.mlp, mixed=mixed)
            return y.half() if mixed else y

    tcnn.NetworkWithInputEncoding = NetworkWithInputEncoding
    sys.modules["tinycudann"] = tcnn

    nerfacc = types.ModuleType("nerfacc")
    nerfacc.accumulate_along_rays = oracle.composite.accumulate_along_rays
    nerfacc.render_weight_from_density = oracle.composite.render_weight_from_density
    nerfacc.render_transmittance_from_alpha = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    est = types.ModuleType("nerfacc.estimators")
    prop = types.ModuleType("nerfacc.estimators.prop_net"); prop.PropNetEstimator = type("PropNetEstimator", (), {})
    occ = types.ModuleType("nerfacc.estimators.occ_grid"); occ.OccGridEstimator = type("OccGridEstimator", (), {})
    sys.modules.update({"nerfacc": nerfacc, "nerfacc.estimators": est,
                        "nerfacc.estimators.prop_net": prop, "nerfacc.estimators.occ_grid": occ})
    for m in [k for k in sys.modules if k.startswith("modules.") or k == "modules"]:
        del sys.modules[m]
    if REF not in sys.path:
        sys.path.insert(1, REF)


class FixedEstimator:
    """Stands where ``OccGridEstimator`` stands in ``NeRFOCCRenderer.render``; returns the
    packed fixed-S intervals."""
    def __init__(self, n_samples, near, far, jitter):
        self.S, self.near, self.far, self.jitter = n_samples, near, far, jitter

    def sampling(self, rays_o, rays_d, sigma_fn=None, stratified=False, **kw):
        R = rays_o.shape[0]
        ts, te = oracle.fixed_samples(R, self.S, self.near, self.far, self.jitter if stratified else None)
        ray_indices = torch.arange(R).repeat_interleave(self.S)
        return ray_indices, ts.reshape(-1), te.reshape(-1)


def rand_pose(g):
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    pose = torch.eye(4); pose[:3, :3] = q; pose[:3, 3] = (torch.rand(3, generator=g) - .5) * .4
    return pose


def make_raygen():
    install_stand_ins(True)
    from utils.camera_utils import gen_pano_rays
    g = torch.Generator().manual_seed(0)
    out = {}
    for name, (pose, h, w, rows) in {
        "eye_8x16": (torch.eye(4), 8, 16, None),
        "rot_6x10": (rand_pose(g), 6, 10, None),
        "rot_128x256": (rand_pose(g), 128, 256, [0, 1, 63, 64, 127]),
        "rot_1024x2048": (rand_pose(g), 1024, 2048, [0, 511, 1023]),
    }.items():
        rays = gen_pano_rays(pose, h, w)
        rows = list(range(h)) if rows is None else rows
        out[name + "_pose"] = pose.numpy(); out[name + "_hw"] = np.array([h, w])
        out[name + "_rows"] = np.array(rows)
        out[name + "_o"] = rays.o[rows].numpy(); out[name + "_d"] = rays.d[rows].numpy()
    from utils.camera_utils import gen_pers_rays
    for name, (pose, fov, res) in {"pers75_64": (rand_pose(g), float(np.deg2rad(75.)), 64), "pers90_33": (rand_pose(g), float(np.deg2rad(90.)), 33)}.items():
        rays = gen_pers_rays(pose, fov=fov, res=res)
        out[name + "_pose"], out[name + "_fov"], out[name + "_res"] = pose.numpy(), np.array(fov), np.array(res)
        out[name + "_o"], out[name + "_d"] = rays.o.numpy(), rays.d.numpy()
    np.savez_compressed(os.path.join(HERE, "raygen.npz"), **out)


def field_points(g, n):
    x = (torch.rand(n, 3, generator=g) * 2 - 1) * 1.1          # some outside the aabb
    x[:8] = torch.tensor([[0., 0., 0.], [1., 0., 0.], [-1., 0.5, 0.5], [0.999999, 0.2, -0.3],
                          [0.5, 0.5, 0.5], [-0.5, 0.25, 0.125], [1.0, 1.0, 1.0], [-1., -1., -1.]])
    return x


def make_field_and_render():
    out_f, out_r = {}, {}
    g = torch.Generator().manual_seed(1)
    x = field_points(g, 2048)
    R, S = 48, 32
    rays_o = (torch.rand(R, 3, generator=g) - .5) * .3
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    jitter = torch.rand(R, generator=g)
    out_f["x"] = x.numpy()
    out_r.update(rays_o=rays_o.numpy(), rays_d=rays_d.numpy(), jitter=jitter.numpy(),
                 n_samples=np.array(S), near=np.array(1e-2), far=np.array(1.0))
    fld = oracle.Field.random(FIELD_SEED, FIELD_GRID_SCALE)
    for mixed in (True, False):
        tag = "mixed" if mixed else "fp32"
        install_stand_ins(mixed)
        from modules.fields.ngp_nerf import NGPNeRF
        from modules.scene.nerf_renderer import NeRFOCCRenderer
        nerf = NGPNeRF(aabb=torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0]))
        with torch.no_grad():
            nerf.geo_mlp.params.copy_(fld.geo_params); nerf.app_mlp.params.copy_(fld.app_params)
            out_f[f"sigma_{tag}"] = nerf.query_density(x).numpy()
            out_f[f"rgb_{tag}"] = nerf.query_rgb(x).float().numpy()
            renderer = NeRFOCCRenderer(max_radius=2, bg_color="rand_noise")
            est = FixedEstimator(S, 1e-2, 1.0, jitter)
            near = 1e-2 * torch.ones(R, 1); far = torch.ones(R, 1)
            for mode in ("eval", "train"):
                nerf.train(mode == "train")
                torch.manual_seed(7)
                res = renderer.render(nerf, est, rays_o, rays_d, near, far)
                for k in ("rgb", "distance", "opacities", "weights", "trans"):
                    out_r[f"{mode}_{tag}_{k}"] = res[k].float().numpy()
            # the random numbers the training branch drew (nerf_renderer.py:185-192)
            torch.manual_seed(7)
            bg = torch.rand(R, 3); dn = torch.rand(R, 1)
            out_r["bg_noise"] = torch.cat([bg, dn], 1).numpy()
    out_f["seed"] = np.array(FIELD_SEED); out_f["grid_scale"] = np.array(FIELD_GRID_SCALE)
    np.savez_compressed(os.path.join(HERE, "field.npz"), **out_f)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out_r)


if __name__ == "__main__":
    make_raygen()
    make_field_and_render()
    for f in ("raygen.npz", "field.npz", "render.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
