"""End-to-end example of the hot path: fit the two-field NeRF to one RGB-D panorama with the
reference's schedule (`configs/nerf.yaml`: 3000 density + 1500 colour iterations of 8192 rays,
`modules/scene/nerf.py:137-184`), then render a panorama like `render_dense`, and report PSNR.

    python examples/fit_and_render.py [--height 512 --width 1024 --n-samples 128 --geo-iters 3000 --app-iters 1500]

The panorama is synthetic (closed-form box room + smooth colours, perf_b200/synthetic.py): the
reference's kitchen example needs Omnidata depth, whose checkpoints are not shipped.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from perf_b200 import synthetic
from perf_b200.scene import NeRFScene, RaySupervision


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--n-samples", type=int, default=128)
    ap.add_argument("--geo-iters", type=int, default=3000)
    ap.add_argument("--app-iters", type=int, default=1500)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--sampler", default="fixed", choices=["fixed", "occ"],
                    help="fixed: the benchmark's S samples per ray; occ: PeRF's occupancy-grid marcher (configs/nerf.yaml:25)")
    ap.add_argument("--out", default=None, help="directory for PNGs (optional)")
    args = ap.parse_args()
    dev = "cuda"
    h, w = args.height, args.width
    rgb, dist = synthetic.smooth_rgb(h, w, seed=0, device=dev), synthetic.box_room_distance(h, w, device=dev)
    conf = dict(NeRFScene(n_samples=8).train_conf)
    conf.update(raw_phase_iter_geo=args.geo_iters, raw_phase_iter_app=args.app_iters)
    torch.manual_seed(0)
    sc = NeRFScene(train_conf=conf, n_samples=args.n_samples, graph_train=not args.no_graph, estimator_type=args.sampler)
    pool = RaySupervision.from_panorama(torch.eye(4), rgb, dist, seed=0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sc.fit(pool)
    torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
    n_steps = args.geo_iters + args.app_iters
    if args.sampler == "occ":
        print(f"fit (occupancy sampler, step 5e-4): {n_steps} steps x 8192 rays in {t_fit:.2f} s ({1e3 * t_fit / n_steps:.2f} ms/step incl. "
              f"the 256 grid updates and graph capture)")
    else:
        print(f"fit: {n_steps} steps x 8192 rays x {args.n_samples} samples in {t_fit:.2f} s "
              f"({1e3 * t_fit / n_steps:.2f} ms/step, {n_steps * 8192 * args.n_samples / t_fit / 1e6:.0f} Msamples/s incl. graph capture)")
    t0 = time.perf_counter()
    out = sc.render_pano(torch.eye(4), h, w)
    torch.cuda.synchronize(); t_r = time.perf_counter() - t0
    psnr = -10 * np.log10(float(((out["rgb"].reshape(h, w, 3) - rgb) ** 2).mean()))
    derr = float((out["distance"].reshape(h, w, -1) - dist.reshape(h, w, -1)).abs().mean())
    print(f"render {h}x{w}x{args.n_samples}: {1e3 * t_r:.1f} ms; PSNR vs supervision {psnr:.2f} dB; mean |distance error| {derr:.4f}")
    pose = torch.eye(4); pose[:3, 3] = torch.tensor([0.15, -0.1, 0.05])
    novel = sc.render_pano(pose, h, w)
    print(f"novel view: rgb mean {float(novel['rgb'].mean()):.3f}, opacity mean {float(novel['opacities'].mean()):.3f}")
    if args.out:
        import cv2
        os.makedirs(args.out, exist_ok=True)
        for name, img in (("fit.png", out["rgb"]), ("novel.png", novel["rgb"]), ("gt.png", rgb)):
            cv2.imwrite(os.path.join(args.out, name), (img.clamp(0, 1) * 255).byte().cpu().numpy()[..., ::-1])


if __name__ == "__main__":
    main()
