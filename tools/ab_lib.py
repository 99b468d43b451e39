"""A/B of two BUILDS of libperfb200.so on the benchmark workload (1024 x 2048 x 128, L2 flushed between runs) and on one
training step: each build runs in its own process (PERF_B200_LIB), alternating, so clocks / box are shared.
    python tools/ab_lib.py perf_b200/_variants/libperfb200_r02base.so [more.so ...]     (the in-tree build is always included)
    python tools/ab_lib.py --child   (internal)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    import bench
    from perf_b200.renderer import FusedPanoRenderer
    geo, app = bench.make_field("cuda")
    pose = bench.bench_pose()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    r = FusedPanoRenderer.from_params(geo, app)
    res = {}
    for name, (H, W, S) in {"c1_1024x2048x128": (bench.H, bench.W, bench.S), "c4_2048x4096x256_rows512": (2048, 4096, 256)}.items():
        rows = H if H == bench.H else 512
        for _ in range(2):
            out = r.render_pano(pose, H, W, S, rows=rows)
        torch.cuda.synchronize()
        best, tot, n = 1e9, 0.0, 5
        for _ in range(n):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = r.render_pano(pose, H, W, S, rows=rows); e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1); tot += ms; best = min(best, ms)
        res[name] = {"ms": tot / n, "best_ms": best, "msamples_s": rows * W * S / (tot / n) / 1e3,
                     "checksum": float(out["rgb"].double().sum()), "dist_checksum": float(out["distance"].double().sum())}
    print("ABRESULT " + json.dumps(res), flush=True)


def main():
    libs = [("in-tree", None)] + [(os.path.basename(p), os.path.abspath(p)) for p in sys.argv[1:]]
    for rnd in range(2):
        for name, path in libs:
            env = dict(os.environ)
            if path:
                env["PERF_B200_LIB"] = path
            else:
                env.pop("PERF_B200_LIB", None)
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=600)
            line = [l for l in out.stdout.splitlines() if l.startswith("ABRESULT ")]
            if not line:
                print(name, "FAILED", out.stderr[-2000:]); continue
            d = json.loads(line[0][9:])
            print(f"round {rnd} {name:34s} " + "  ".join(f"{k}: {v['ms']:.3f} ms (best {v['best_ms']:.3f}) {v['msamples_s']:.0f} Ms/s sum {v['checksum']:.4f}" for k, v in d.items()), flush=True)


if __name__ == "__main__":
    child() if "--child" in sys.argv else main()
