"""Render-kernel variants on the benchmark workload (1024 x 2048 x 128, L2 flushed between runs): Msamples/s of each
selectable kernel and bit-equality of the experimental ones with the default.  VERDICT r1 next #4.
    python tools/render_variants.py [steps=5]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from perf_b200.renderer import FusedPanoRenderer  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
geo, app = bench.make_field("cuda")
pose = bench.bench_pose()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for kernel in ("march", "march_l0smem", "march_generic", "scan"):
    r = FusedPanoRenderer.from_params(geo, app, kernel=kernel)
    for _ in range(2):
        out = r.render_pano(pose, bench.H, bench.W, bench.S)
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = r.render_pano(pose, bench.H, bench.W, bench.S); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= steps
    if ref is None:
        ref = out
    same = all(torch.equal(out[k], ref[k]) for k in ("rgb", "distance", "opacities"))
    err = max(float((out[k] - ref[k]).abs().max()) for k in ("rgb", "distance", "opacities"))
    print(f"{kernel:14s}: {ms:8.3f} ms  {bench.H * bench.W * bench.S / ms / 1e3:8.1f} Msamples/s   bit-identical to default: {same} (max |d| {err:.1e})", flush=True)
