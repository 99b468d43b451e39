"""Oracle: multi-resolution hash-grid encode (tiny-cuda-nn ``HashGrid`` semantics).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Restates the algorithm of
tiny-cuda-nn 1.7 ``include/tiny-cuda-nn/encodings/grid.h`` (third-party, pinned by
`/root/reference/requirements.txt:34`, not vendored) as configured by the reference at
`/root/reference/modules/fields/ngp_nerf.py:94-134` (L=16, F=2, T=2^18, base 16,
per_level_scale 1.4472692012786865).  SURVEY.md Appendix A is the written contract.

Every upstream-derived rule lives in exactly one named function here:
``grid_scale`` / ``grid_resolution`` / ``level_table`` (offset table),
``pos_fract`` (fmaf(scale, x, 0.5), floor, fractional part),
``grid_index`` (dense stride walk, xor-prime hash, ``% hashmap_size``),
``encode`` (8-corner trilinear blend, corner order bit0=x, level-major output).

Blend arithmetic: ``blend="half"`` is tcnn's own for ``__half`` parameters --
``result = fma((half)weight_k, value_k, result)`` over corners k = 0..7 with ONE fp16
rounding per fma, starting from zero -- and is what the CUDA kernels implement (HFMA2),
so oracle and kernel are comparable to the bit.  ``blend="fp32"`` (fp32 `fmaf` chain,
optionally rounded to fp16 at the end) is the plain-precision reference used by the
``mixed=False`` field and by gradient checks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

PRIMES = (1, 2654435761, 805459861)
_U32 = 0xFFFFFFFF


@dataclass(frozen=True)
class GridConfig:
    """Mirror of the ``encoding_config`` dict the reference passes to tcnn
    (`ngp_nerf.py:99-106`)."""
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 18
    base_resolution: int = 16
    per_level_scale: float = 1.4472692012786865
    interpolation: str = "Linear"          # "Linear" | "Smoothstep"

    @staticmethod
    def from_dict(d: dict) -> "GridConfig":
        assert d.get("otype", "HashGrid") in ("HashGrid", "Grid"), d
        return GridConfig(
            n_levels=int(d.get("n_levels", 16)),
            n_features_per_level=int(d.get("n_features_per_level", 2)),
            log2_hashmap_size=int(d.get("log2_hashmap_size", 19)),
            base_resolution=int(d.get("base_resolution", 16)),
            per_level_scale=float(d.get("per_level_scale", 2.0)),
            interpolation=str(d.get("interpolation", "Linear")),
        )


@dataclass(frozen=True)
class Level:
    scale: np.float32      # grid_scale(level)
    resolution: int        # ceil(scale) + 1
    size: int              # entries in this level (hashmap_size)
    offset: int            # first entry of this level in the flat table
    hashed: bool           # dense stride walk overflowed the level -> xor-prime hash


def grid_scale(level: int, per_level_scale: float, base_resolution: int) -> np.float32:
    """tcnn ``grid_scale``: ``exp2f(level * log2f(s)) * base - 1``.  Upstream evaluates it in fp32 on
    the device, where ``exp2f`` is only accurate to ~2 ulp, so the last bit is implementation
    defined; we pin it: ``x = fp32(level) * fp32(log2(s))`` then ``fp32(exp2(fp64(x)) * base - 1)``
    with a single rounding.  The C library computes the same on the host (`perf_grid_describe`)
    and a hypothesis test asserts bit-equality over random configurations."""
    log2s = np.float32(np.log2(np.float64(np.float32(per_level_scale))))
    x = np.float32(np.float32(level) * log2s)
    return np.float32(np.exp2(np.float64(x)) * np.float64(base_resolution) - 1.0)


def grid_resolution(scale: np.float32) -> int:
    """tcnn ``grid_resolution``: ``(uint32)ceilf(scale) + 1``."""
    return int(np.ceil(np.float32(scale))) + 1


def level_table(cfg: GridConfig) -> List[Level]:
    """tcnn ``GridEncodingTemplated`` constructor: per-level entry counts, rounded up
    to a multiple of 8 and clamped to 2^T; SURVEY.md Appendix A table."""
    levels, offset = [], 0
    for l in range(cfg.n_levels):
        scale = grid_scale(l, cfg.per_level_scale, cfg.base_resolution)
        res = grid_resolution(scale)
        max_params = 0xFFFFFFFF // 2
        dense = res ** 3 if float(res) ** 3 <= float(max_params) else max_params
        dense = (dense + 7) // 8 * 8
        size = min(dense, 1 << cfg.log2_hashmap_size)
        # "hashed" == the stride walk in grid_index ends with stride > size
        stride, dims = 1, 0
        while dims < 3 and stride <= size:
            stride *= res
            dims += 1
        levels.append(Level(scale, res, size, offset, size < stride))
        offset += size
    return levels


def n_table_entries(cfg: GridConfig) -> int:
    lv = level_table(cfg)
    return lv[-1].offset + lv[-1].size


def pos_fract(x: torch.Tensor, scale: np.float32):
    """tcnn ``pos_fract``: pos = fmaf(scale, x, 0.5); g = floor(pos); w = pos - g.

    fmaf's single rounding is emulated through fp64 (the 24x24-bit product and the
    add of 0.5 are exact in fp64 for x in [0, 1])."""
    pos = (x.double() * float(scale) + 0.5).float()
    g = torch.floor(pos)
    w = pos - g
    return g.to(torch.int64) & _U32, w


def grid_index(gc: torch.Tensor, lvl: Level) -> torch.Tensor:
    """tcnn ``grid_index`` with ``GridType::Hash`` / ``HashType::CoherentPrime``.
    ``gc``: int64 [..., 3] holding uint32 values.  Returns int64 index < lvl.size."""
    if lvl.hashed:
        idx = (gc[..., 0] * PRIMES[0]) & _U32
        idx = idx ^ ((gc[..., 1] * PRIMES[1]) & _U32)
        idx = idx ^ ((gc[..., 2] * PRIMES[2]) & _U32)
    else:
        stride, idx = 1, torch.zeros_like(gc[..., 0])
        for dim in range(3):
            if stride > lvl.size:
                break
            idx = (idx + gc[..., dim] * stride) & _U32
            stride = (stride * lvl.resolution) & _U32
    return idx % lvl.size


def _corner_weights_indices(x01: torch.Tensor, lvl: Level, smoothstep: bool):
    g, w = pos_fract(x01, lvl.scale)
    if smoothstep:
        w = w * w * (3.0 - 2.0 * w)
    out = []
    for c in range(8):
        wt = torch.ones_like(w[..., 0])
        gc = g.clone()
        for dim in range(3):
            if c & (1 << dim):
                wt = wt * w[..., dim]
                gc[..., dim] = (g[..., dim] + 1) & _U32
            else:
                wt = wt * (1.0 - w[..., dim])
        out.append((wt, grid_index(gc, lvl) + lvl.offset))
    return out


def _round_half_exact(x64: torch.Tensor) -> torch.Tensor:
    """fp64 -> fp16 with ONE correct rounding (numpy converts directly; torch goes through fp32)."""
    return torch.from_numpy(x64.numpy().astype(np.float16))


def encode(x01: torch.Tensor, table: torch.Tensor, cfg: GridConfig = GridConfig(),
           out_half: bool = False, exact_fma: bool = True, blend: str = "fp32") -> torch.Tensor:
    """Hash-grid encode.  ``x01`` [N,3] fp32 in [0,1]; ``table`` [n_entries, F] (any float
    dtype; values are used as fp32).  Returns [N, L*F] fp32, level-major
    (``[l0f0, l0f1, l1f0, ...]``).

    ``blend="half"``: tcnn's fp16 fma chain (table values and weights rounded to fp16, one
    fp16 rounding per corner); the result is fp16-valued.  ``blend="fp32"``: fp32 fmaf chain;
    with ``out_half`` the finished feature is rounded to fp16.  ``exact_fma=False`` selects
    the fast variants used when the oracle is *timed* as the CPU baseline (fp32 intermediates
    instead of fp64: identical except on rounding ties)."""
    assert x01.dim() == 2 and x01.shape[1] == 3 and blend in ("fp32", "half")
    F = cfg.n_features_per_level
    table = table.reshape(-1, F).float()
    x01 = x01.float()
    feats = []
    for lvl in level_table(cfg):
        if blend == "half":
            corners = _corner_weights_indices(x01, lvl, cfg.interpolation == "Smoothstep")
            if exact_fma and not table.requires_grad:
                acc = torch.zeros(x01.shape[0], F, dtype=torch.float64)
                for wt, idx in corners:
                    acc = _round_half_exact(wt.half().double()[:, None] * table[idx].half().double() + acc).double()
                feats.append(acc.float())
            else:
                # differentiable / fast variant: fp32 intermediate, fp16 rounding as a straight-through cast
                acc = torch.zeros(x01.shape[0], F)
                for wt, idx in corners:
                    acc = (wt.half().float()[:, None] * table[idx].half().float() + acc).half().float()
                feats.append(acc)
        elif exact_fma:
            acc = torch.zeros(x01.shape[0], F, dtype=torch.float64)
            for wt, idx in _corner_weights_indices(x01, lvl, cfg.interpolation == "Smoothstep"):
                # fp32 fmaf(weight, value, acc): exact product + add in fp64, one fp32 rounding
                acc = (wt.double()[:, None] * table[idx].double() + acc).float().double()
            feats.append(acc.float())
        else:
            acc = torch.zeros(x01.shape[0], F)
            for wt, idx in _corner_weights_indices(x01, lvl, cfg.interpolation == "Smoothstep"):
                acc = acc + wt[:, None] * table[idx]
            feats.append(acc)
    out = torch.cat(feats, dim=1)
    if out_half:
        out = out.half().float()
    return out


def encode_backward_table(x01: torch.Tensor, dfeat: torch.Tensor,
                          cfg: GridConfig = GridConfig()) -> torch.Tensor:
    """d(loss)/d(table) for ``encode``: scatter-add of ``weight * dfeat`` into the 8
    corners of every level (tcnn ``kernel_grid_backward``).  fp64 accumulation; returns
    fp32 [n_entries, F]."""
    F = cfg.n_features_per_level
    grad = torch.zeros(n_table_entries(cfg), F, dtype=torch.float64)
    x01 = x01.float()
    for l, lvl in enumerate(level_table(cfg)):
        d = dfeat[:, l * F:(l + 1) * F].double()
        for wt, idx in _corner_weights_indices(x01, lvl, cfg.interpolation == "Smoothstep"):
            grad.index_add_(0, idx, wt.double()[:, None] * d)
    return grad.float()


def encode_autograd(x01: torch.Tensor, table: torch.Tensor, cfg: GridConfig = GridConfig(),
                    fp32_positions: bool = False) -> torch.Tensor:
    """Hash-grid encode as a plain differentiable fp64 torch expression: autograd through it gives the
    oracle for the gradients w.r.t. the INPUT positions (tcnn ``kernel_grid_backward_input``) and for
    their double backward (``kernel_grid_backward_input_backward_grid / _backward_dLdoutput /
    _backward_input``), which `SphereDistanceField.forward(requires_grad=True)` uses
    (`/root/reference/modules/geo_predictors/pano_joint_predictor.py:48-68`).

    ``x01`` [N,3] (grad flows), ``table`` [n_entries, F] (grad flows; pass fp16-rounded values to match
    the kernels, which read the fp16 shadow).  The cell of every sample is taken from the fp32
    ``pos_fract`` (bit-identical to the kernels); inside the cell the fractional position is
    ``scale*x + 0.5 - cell`` in fp64, whose derivative w.r.t. x is ``scale`` (tcnn: ``floor`` has no
    gradient).  Smoothstep: ``s = w^2 (3 - 2w)``, so ``ds/dw = 6w(1-w)`` and ``d2s/dw2 = 6 - 12w``.

    ``fp32_positions``: take the VALUE of the fractional position from the fp32 ``pos_fract`` as tcnn and
    the kernels do (at scale 2047 an fp32 position carries ~3e-5 of absolute error, which the smoothstep
    derivative amplifies to ~1e-4 relative) while keeping ``d w / d x = scale``.  This is the mode the
    kernels are compared in; the pure-fp64 mode is the one ``gradcheck`` can differentiate numerically."""
    F = cfg.n_features_per_level
    table = table.reshape(-1, F).double()
    xd = x01.double()
    feats = []
    for lvl in level_table(cfg):
        g, w32 = pos_fract(x01.detach().float(), lvl.scale)
        if fp32_positions:
            w = w32.double() + (xd - xd.detach()) * float(lvl.scale)
        else:
            g_signed = torch.floor((x01.detach().float().double() * float(lvl.scale) + 0.5).float()).double()
            w = xd * float(lvl.scale) + 0.5 - g_signed
        if cfg.interpolation == "Smoothstep":
            w = w * w * (3.0 - 2.0 * w)
        acc = torch.zeros(x01.shape[0], F, dtype=torch.float64)
        for c in range(8):
            wt = torch.ones_like(w[..., 0])
            gc = g.clone()
            for dim in range(3):
                if c & (1 << dim):
                    wt = wt * w[..., dim]
                    gc[..., dim] = (g[..., dim] + 1) & _U32
                else:
                    wt = wt * (1.0 - w[..., dim])
            acc = acc + wt[:, None] * table[grid_index(gc, lvl) + lvl.offset]
        feats.append(acc)
    return torch.cat(feats, dim=1)


def encode_input_grad(x01: torch.Tensor, table: torch.Tensor, dfeat: torch.Tensor,
                      cfg: GridConfig = GridConfig(), create_graph: bool = False, fp32_positions: bool = False):
    """d(loss)/d(x01) [N,3] fp64 = sum_f dfeat_f * d feat_f / d x01 (tcnn ``kernel_grid_backward_input``),
    by autograd through :func:`encode_autograd`."""
    x = x01.double() if x01.requires_grad else x01.detach().double().requires_grad_(True)
    y = encode_autograd(x, table, cfg, fp32_positions)
    return torch.autograd.grad(y, x, grad_outputs=dfeat.double(), create_graph=create_graph)[0]


def encode_input_grad_backward(x01: torch.Tensor, table: torch.Tensor, dfeat: torch.Tensor, ddx: torch.Tensor,
                               cfg: GridConfig = GridConfig(), fp32_positions: bool = False):
    """Double backward of :func:`encode_input_grad`: gradients of ``(dx * ddx).sum()`` w.r.t.
    ``(dfeat, table, x01)`` (tcnn ``kernel_grid_backward_input_backward_dLdoutput / _grid / _input``)."""
    x = x01.detach().double().requires_grad_(True)
    t = table.detach().double().requires_grad_(True)
    g = dfeat.detach().double().requires_grad_(True)
    y = encode_autograd(x, t, cfg, fp32_positions)
    dx = torch.autograd.grad(y, x, grad_outputs=g, create_graph=True)[0]
    return torch.autograd.grad((dx * ddx.double()).sum(), (g, t, x), allow_unused=True)
