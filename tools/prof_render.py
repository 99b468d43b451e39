"""One warm-up + one measured fused render (256 rows of the 1024x2048x128 panorama) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from perf_b200.renderer import FusedPanoRenderer
geo, app = bench.make_field("cuda")
r = FusedPanoRenderer.from_params(geo, app, kernel=os.environ.get('KERNEL', 'march'))
pose = bench.bench_pose()
rows = int(os.environ.get("ROWS", 256))
for _ in range(2):
    out = r.render_pano(pose, bench.H, bench.W, bench.S, row0=(bench.H - rows) // 2, rows=rows)
torch.cuda.synchronize()
print("ok", float(out["rgb"].mean()))
