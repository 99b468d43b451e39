"""Bring-up diagnosis of perf_mlp_bwd: per-block errors (dfeat, dW1, dW2, dWout) against a torch fp32 reference.
    python tools/diag_mlp_bwd.py {density|colour} {simt|tc} [dbg]
One process per variant (a trapping kernel poisons its CUDA context)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from perf_b200 import ops  # noqa: E402
from perf_b200.config import APP_MLP, GEO_MLP  # noqa: E402

net, mode = sys.argv[1], sys.argv[2]
dbg = int(sys.argv[3]) if len(sys.argv) > 3 else 0
two = net == "colour"
mlp = APP_MLP if two else GEO_MLP
g = torch.Generator().manual_seed(7)
N = 128 * 5 + 53
W = ((torch.rand(mlp.n_params, generator=g) * 2 - 1) * 0.3).half().cuda()
feat = ((torch.rand(N, 32, generator=g) * 2 - 1) * 0.5).half().cuda()
w1 = W[:2048].view(64, 32).float()
h1 = torch.relu(feat.float() @ w1.t()).half()
p = 2048
w2 = None
h2 = None
if two:
    w2 = W[p:p + 4096].view(64, 64).float(); p += 4096
    h2 = torch.relu(h1.float() @ w2.t()).half()
wout = W[p:p + 16 * 64].view(16, 64)[:mlp.n_out].float()
dz = (torch.randn(N, mlp.n_out, generator=g) * 0.1).cuda()
dzh = dz.half().float()
hl = (h2 if two else h1).float()
dh = ((dz @ wout) * (hl > 0)).half().float()
ref = {"dWout": dzh.t() @ hl}
if two:
    ref["dW2"] = dh.t() @ h1.float()
    dh = ((dh @ w2) * (h1.float() > 0)).half().float()
ref["dW1"] = dh.t() @ feat.float()
ref["dfeat"] = dh @ w1
got_w, got_f = ops.mlp_backward_fused(mlp, W, feat, h1, h2, dz, simt=(mode == "simt"), dbg=dbg)
torch.cuda.synchronize()
got = {"dfeat": got_f, "dW1": got_w[:2048].view(64, 32)}
if two:
    got["dW2"] = got_w[2048:6144].view(64, 64)
got["dWout"] = got_w[p:p + 16 * 64].view(16, 64)[:mlp.n_out]
for k in ref:
    e = (got[k] - ref[k]).abs().max().item(); m = ref[k].abs().max().item()
    print(f"{net} {mode} dbg={dbg} {k:6s}: max err {e:.3e} of {m:.3e}  rel {e / m:.2e}  {'OK' if e <= 3e-3 * m else 'BAD'}", flush=True)
