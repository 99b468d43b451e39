"""Tiny end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck` (and racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from perf_b200 import ops, synthetic
from perf_b200.config import APP_MLP, GEO_MLP, PERF_GRID
from perf_b200.renderer import FusedPanoRenderer
from perf_b200.scene import FusedAdam, GraphedTrainStep, NeRFScene, RaySupervision

torch.manual_seed(0)
n_e = PERF_GRID.n_entries
mk = lambda mlp: torch.cat([(torch.rand(mlp.n_params, device="cuda") - .5) * .6, (torch.rand(2 * n_e, device="cuda") - .5)])
r = FusedPanoRenderer.from_params(mk(GEO_MLP), mk(APP_MLP))
pose = torch.eye(4)
for kern in ("march", "march_generic", "scan"):
    r.kernel = kern
    r.render_pano(pose, 20, 40, 24)                     # ragged patches (20 % 8 != 0, 40 % 16 != 0)
    o, d = ops.raygen_pano(pose, 6, 10)
    r.render_rays(o.reshape(-1, 3), d.reshape(-1, 3), 7)
    r.render_rays(o, d, 32)
r.kernel = "march"
ri = torch.tensor([0, 0, 0, 2, 2, 5], device="cuda"); ts = torch.rand(6, device="cuda")
r.render_packed(o.reshape(-1, 3)[:8], d.reshape(-1, 3)[:8], ri, ts, ts + 0.01)
ops.raygen_pers(pose, 1.2, 9)
x = torch.rand(300, 3, device="cuda")
gh = ops.params_to_half(mk(GEO_MLP))
feat = ops.hashgrid_fwd(gh[GEO_MLP.n_params:].view(-1, 2), x)
ops.hashgrid_bwd(x, torch.randn(300, 32, device="cuda"))
ops.network_fwd(gh, x, PERF_GRID, GEO_MLP, save=True)
ops.network_fwd(ops.params_to_half(mk(APP_MLP)), x, PERF_GRID, APP_MLP, save=True)
ops.mlp_fwd(gh[:GEO_MLP.n_params].clone(), feat, GEO_MLP, save=True)
rid = torch.repeat_interleave(torch.arange(10, device="cuda"), torch.tensor([0, 3, 40, 1, 0, 33, 32, 7, 0, 2], device="cuda"))
n = rid.numel(); t0 = torch.rand(n, device="cuda"); sg = torch.rand(n, device="cuda") * 9
w, T, a = ops.weights_from_density(t0, t0 + .01, sg, rid, 10)
ops.weights_from_density_bwd(t0, t0 + .01, sg, rid, 10, w, T, torch.randn(n, device="cuda"), torch.randn(n, device="cuda"))
ops.accumulate_along_rays(w, torch.rand(n, 3, device="cuda"), rid, 10)
binaries = torch.rand(16, 16, 16, device="cuda") < 0.3
ops.occ_sample(binaries, [-1, -1, -1, 1, 1, 1], o.reshape(-1, 3), d.reshape(-1, 3), 0.0, 1.5, 0.02, torch.rand(60, device="cuda"))
# training: eager fused (segments), graphed, modular, occ
h, wd = 16, 32
rgb, dist = synthetic.smooth_rgb(h, wd, device="cuda"), synthetic.box_room_distance(h, wd, device="cuda")
for fused in (True, False):
    sc = NeRFScene(n_samples=24, fused_train=fused); sc.train_conf["pixel_loss_batch_size"] = 100
    sc.set_train(); pool = RaySupervision.from_panorama(torch.eye(4), rgb, dist)
    for phase in ("geo", "app"):
        net = sc.nerf.geo_mlp if phase == "geo" else sc.nerf.app_mlp
        opt = FusedAdam(net.params, lr=1e-3, module=net)
        (sc.train_one_step_geo if phase == "geo" else sc.train_one_step_app)(opt, pool, progress=0.3)
sc.render_pano(torch.eye(4), h, wd)
# round 2: occupancy-sampler scene (native grid update, fused packed train step both phases, fused eval render), the
# L0-in-shared-memory render variant, one-kernel loss, batch gather, vector / scalar Adam, mlp_bwd CUDA-core twin
sc = NeRFScene(estimator_type="occ", occ_resolution=32); sc.train_conf["pixel_loss_batch_size"] = 100
pool = RaySupervision.from_panorama(torch.eye(4), rgb, dist)
sc.build_occupancy(pool, n_updates=3)
sc.OCC_STEP = 5e-3
sc.set_train()
for phase in ("geo", "app"):
    net = sc.nerf.geo_mlp if phase == "geo" else sc.nerf.app_mlp
    opt = FusedAdam(net.params, lr=1e-3, module=net)
    (sc.train_one_step_geo if phase == "geo" else sc.train_one_step_app)(opt, pool, progress=0.3)
sc.set_eval()
sc.render_pano(torch.eye(4), h, wd)
r.kernel = "march_l0smem"; r.render_pano(pose, 20, 40, 24); r.kernel = "march"
W = ((torch.rand(APP_MLP.n_params, device="cuda") - .5) * .6).half()
f16 = (torch.rand(300, 32, device="cuda") - .5).half(); h1 = torch.rand(300, 64, device="cuda").half(); h2 = torch.rand(300, 64, device="cuda").half()
for simt in (True, False):
    ops.mlp_backward_fused(APP_MLP, W, f16, h1, h2, torch.randn(300, 3, device="cuda"), simt=simt)
    ops.mlp_backward_fused(GEO_MLP, W[:GEO_MLP.n_params].clone(), f16, h1, None, torch.randn(300, 1, device="cuda"), simt=simt)
pp = torch.randn(1003, device="cuda"); ops.adam_step(pp, torch.randn(1003, device="cuda"), torch.zeros(1003, device="cuda"), torch.zeros(1003, device="cuda"), 1, 1e-3,
                                                   params_half=torch.empty(1003, dtype=torch.float16, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
