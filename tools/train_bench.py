"""Timing probe of the training step (configs[1]: 8192 rays x 128 samples per step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from perf_b200 import synthetic, parallel
from perf_b200.scene import NeRFScene, RaySupervision, FusedAdam, GraphedTrainStep

rank, world, local = parallel.init()
torch.cuda.set_device(local)
h, w, S, B = 512, 1024, 128, 8192
rgb = synthetic.smooth_rgb(h, w, device="cuda"); dist = synthetic.box_room_distance(h, w, device="cuda")
occ = os.environ.get('OCC', '0') == '1'
sc = NeRFScene(n_samples=S, fused_train=os.environ.get('FUSED', '1') == '1', **({"estimator_type": "occ", "occ_resolution": 256} if occ else {}))
pool = RaySupervision.from_panorama(torch.eye(4), rgb, dist)
if occ:
    sc.build_occupancy(pool)
sc.set_train()
for phase in os.environ.get("PHASES", "geo,app").split(","):
    net = sc.nerf.geo_mlp if phase == "geo" else sc.nerf.app_mlp
    opt = FusedAdam(net.params, lr=1e-3, module=net)
    if os.environ.get("GRAPH", "0") == "1":
        g = GraphedTrainStep(sc, phase, pool, opt)
        step = lambda opt_, pool_, progress=0.5: g(progress)
    else:
        step = sc.train_one_step_geo if phase == "geo" else sc.train_one_step_app
    for _ in range(5):
        step(opt, pool, progress=0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get('NSTEPS', 20))
    e0.record()
    for _ in range(n):
        step(opt, pool, progress=0.5)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    if rank == 0:
        print(f"[{phase}] fused={sc.fused_train} graph={os.environ.get('GRAPH', '0')} train step B={B} S={S} world={world}: {ms:.3f} ms/step  {B*S/ms/1e3:.1f} Msamples/s", flush=True)
