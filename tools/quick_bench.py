"""Throw-away timing probe (not the bench contract): fused render at several sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from perf_b200.renderer import FusedPanoRenderer
from perf_b200.config import GEO_MLP, APP_MLP, PERF_GRID

torch.manual_seed(0)
n_e = PERF_GRID.n_entries
def rand_params(mlp):
    w = (torch.rand(mlp.n_params, device="cuda") * 2 - 1) * 0.3
    g = (torch.rand(2 * n_e, device="cuda") * 2 - 1) * 0.5
    return torch.cat([w, g])
r = FusedPanoRenderer.from_params(rand_params(GEO_MLP), rand_params(APP_MLP))
pose = torch.eye(4)
import itertools
for kern, (H, W, S, rows) in itertools.product(["march", "march_generic"], [(128, 256, 32, 128), (1024, 2048, 128, 128), (1024, 2048, 128, 1024), (2048, 4096, 256, 256)]):
    r.kernel = kern
    out = r.render_pano(pose, H, W, S, rows=rows); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        r.render_pano(pose, H, W, S, rows=rows, out=(out["rgb"], out["distance"], out["opacities"]))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ns = rows * W * S
    print(f"[{kern}] render {H}x{W} rows={rows} S={S}: {ms:.3f} ms  {ns/ms/1e3:.1f} Msamples/s  algGB/s={ns*1024/ms/1e6:.0f}", flush=True)

# explicit rays: image-shaped (pixel-patch tiling) vs flattened (128 consecutive rays per tile)
from perf_b200 import ops
r.kernel = "march"
H, W, S, rows = 1024, 2048, 128, 256
o, d = ops.raygen_pano(pose, H, W, row0=384, rows=rows)
for name, (oo, dd) in {"image [H,W,3]": (o, d), "flat [R,3]": (o.reshape(-1, 3), d.reshape(-1, 3))}.items():
    r.render_rays(oo, dd, S); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r.render_rays(oo, dd, S)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"render_rays {name}: {ms:.3f} ms  {rows*W*S/ms/1e3:.1f} Msamples/s", flush=True)
