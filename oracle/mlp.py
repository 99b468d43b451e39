"""Oracle: bias-free 64-wide MLP (tiny-cuda-nn ``FullyFusedMLP`` semantics).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates tiny-cuda-nn 1.7
``src/fully_fused_mlp.cu`` (third-party, not vendored) as configured by the reference
at `/root/reference/modules/fields/ngp_nerf.py:107-113,127-133`:
density net 32 -> 64 ReLU -> 1 (no output activation), colour net
32 -> 64 ReLU -> 64 ReLU -> 3 Sigmoid; no biases; weight matrices row-major
``[out, in]``; the last matrix is padded to 16 output rows of which only the first
``n_output_dims`` are meaningful.

Precision contract (``mixed=True``, what the CUDA path implements; the encode feeding it
uses tcnn's fp16 blend, see ``oracle/hashgrid.py``): operands are
fp16 (input features, weights, hidden activations after ReLU), every dot product is
accumulated wider (tcnn: fp16 accumulators; ours: fp32 in TMEM -- stated deviation,
strictly more accurate), the output pre-activation is rounded to fp16, the output
activation is evaluated in fp32 and the result rounded to fp16 (tcnn returns
``__half``).  ``mixed=False`` is the plain fp32 network.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch


@dataclass(frozen=True)
class MLPConfig:
    n_in: int = 32
    n_out: int = 1
    n_neurons: int = 64
    n_hidden_layers: int = 1
    activation: str = "ReLU"
    output_activation: str = "None"      # "None" | "Sigmoid"

    @property
    def padded_out(self) -> int:
        return (self.n_out + 15) // 16 * 16

    @staticmethod
    def from_dict(d: dict, n_in: int, n_out: int) -> "MLPConfig":
        assert d.get("otype", "FullyFusedMLP") in ("FullyFusedMLP", "CutlassMLP"), d
        return MLPConfig(n_in=n_in, n_out=n_out,
                         n_neurons=int(d.get("n_neurons", 64)),
                         n_hidden_layers=int(d.get("n_hidden_layers", 1)),
                         activation=str(d.get("activation", "ReLU")),
                         output_activation=str(d.get("output_activation", "None")))


def layer_shapes(cfg: MLPConfig):
    """[(out, in), ...] in flat-``params`` order (tcnn layout)."""
    shapes = [(cfg.n_neurons, cfg.n_in)]
    shapes += [(cfg.n_neurons, cfg.n_neurons)] * (cfg.n_hidden_layers - 1)
    shapes += [(cfg.padded_out, cfg.n_neurons)]
    return shapes


def flat_param_count(cfg: MLPConfig) -> int:
    return sum(o * i for o, i in layer_shapes(cfg))


def split_params(flat: torch.Tensor, cfg: MLPConfig) -> List[torch.Tensor]:
    out, p = [], 0
    for o, i in layer_shapes(cfg):
        out.append(flat[p:p + o * i].reshape(o, i))
        p += o * i
    return out


def _act(x: torch.Tensor, name: str) -> torch.Tensor:
    if name == "ReLU":
        return torch.relu(x)
    if name == "Sigmoid":
        return torch.sigmoid(x)
    if name == "None":
        return x
    raise ValueError(f"unsupported activation {name!r}")


def mlp_forward(x: torch.Tensor, weights: List[torch.Tensor], cfg: MLPConfig,
                mixed: bool = True, accum: torch.dtype = torch.float64) -> torch.Tensor:
    """``x`` [N, n_in] -> [N, n_out] fp32.

    ``mixed``: fp16 operands / wide accumulate / fp16 rounding points described in the
    module docstring (``accum=float64`` gives the correctly-rounded accumulation the
    tensor-core result is compared to; ``float32`` is the fast variant used when the
    oracle is *timed* as the CPU baseline)."""
    if not mixed:
        h = x.float()
        for W in weights[:-1]:
            h = _act(h @ W.float().t(), cfg.activation)
        return _act(h @ weights[-1].float()[:cfg.n_out].t(), cfg.output_activation)

    q = lambda t: t.half().to(accum)          # round to fp16, compute wide
    h = q(x)
    for W in weights[:-1]:
        h = q(_act((h @ q(W).t()).float(), cfg.activation))
    o = (h @ q(weights[-1][:cfg.n_out]).t()).float().half().float()
    return _act(o, cfg.output_activation).half().float()


def mlp_hidden(x: torch.Tensor, weights: List[torch.Tensor], cfg: MLPConfig,
               accum: torch.dtype = torch.float64) -> List[torch.Tensor]:
    """Hidden activations (fp16-valued, fp32 dtype) of the mixed-precision network;
    used by the backward oracle and by layer-by-layer kernel tests."""
    q = lambda t: t.half().to(accum)
    hs, h = [], q(x)
    for W in weights[:-1]:
        h = q(_act((h @ q(W).t()).float(), cfg.activation))
        hs.append(h.float())
    return hs
