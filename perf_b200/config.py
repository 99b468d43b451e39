"""Grid / MLP configurations (host mirror of perf_grid_cfg / perf_mlp_cfg).

Field shapes follow `/root/reference/modules/fields/ngp_nerf.py:94-134`: both networks use the
same hash grid (L=16, F=2, T=2^18, base 16, per_level_scale 1.4472692012786865); the density
net is 32->64->1 (no output activation), the colour net 32->64->64->3 (Sigmoid).
"""
from __future__ import annotations

from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class GridConfig:
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 18
    base_resolution: int = 16
    per_level_scale: float = 1.4472692012786865
    interpolation: str = "Linear"

    @staticmethod
    def from_dict(d: dict) -> "GridConfig":
        otype = d.get("otype", "HashGrid")
        if otype not in ("HashGrid", "Grid"):
            raise ValueError(f"unsupported encoding otype {otype!r} (only HashGrid)")
        if d.get("type", "Hash") != "Hash":
            raise ValueError(f"unsupported grid type {d.get('type')!r} (only Hash)")
        return GridConfig(int(d.get("n_levels", 16)), int(d.get("n_features_per_level", 2)),
                          int(d.get("log2_hashmap_size", 19)), int(d.get("base_resolution", 16)),
                          float(d.get("per_level_scale", 2.0)), str(d.get("interpolation", "Linear")))

    def c(self) -> "_lib.GridCfg":
        interp = {"Linear": 0, "Smoothstep": 1}
        if self.interpolation not in interp:
            raise ValueError(f"unsupported interpolation {self.interpolation!r}")
        return _lib.GridCfg(self.n_levels, self.n_features_per_level, self.log2_hashmap_size,
                            self.base_resolution, self.per_level_scale, interp[self.interpolation])

    @property
    def n_features(self) -> int:
        return self.n_levels * self.n_features_per_level

    def levels(self):
        lv = (_lib.Level * self.n_levels)()
        n = _lib.u64(0)
        cfg = self.c()
        _lib.check(_lib.load().perf_grid_describe(cfg, lv, n))
        return list(lv), int(n.value)

    @property
    def n_entries(self) -> int:
        return self.levels()[1]


@dataclass(frozen=True)
class MLPConfig:
    n_in: int = 32
    n_out: int = 1
    n_neurons: int = 64
    n_hidden_layers: int = 1
    output_activation: str = "None"

    @staticmethod
    def from_dict(d: dict, n_in: int, n_out: int) -> "MLPConfig":
        if d.get("otype", "FullyFusedMLP") not in ("FullyFusedMLP", "CutlassMLP"):
            raise ValueError(f"unsupported network otype {d.get('otype')!r}")
        if d.get("activation", "ReLU") != "ReLU":
            raise ValueError(f"unsupported activation {d.get('activation')!r} (only ReLU)")
        return MLPConfig(n_in, n_out, int(d.get("n_neurons", 64)), int(d.get("n_hidden_layers", 1)),
                         str(d.get("output_activation", "None")))

    def c(self) -> "_lib.MlpCfg":
        act = {"None": 0, "Sigmoid": 1}
        if self.output_activation not in act:
            raise ValueError(f"unsupported output_activation {self.output_activation!r}")
        return _lib.MlpCfg(self.n_in, self.n_out, self.n_neurons, self.n_hidden_layers, act[self.output_activation])

    @property
    def padded_out(self) -> int:
        return (self.n_out + 15) // 16 * 16

    @property
    def n_params(self) -> int:
        return (self.n_neurons * self.n_in + (self.n_hidden_layers - 1) * self.n_neurons ** 2
                + self.padded_out * self.n_neurons)


PERF_GRID = GridConfig()
GEO_MLP = MLPConfig(32, 1, 64, 1, "None")       # ngp_nerf.py:107-113
APP_MLP = MLPConfig(32, 3, 64, 2, "Sigmoid")    # ngp_nerf.py:127-133


def network_param_count(grid: GridConfig, mlp: MLPConfig) -> int:
    n = _lib.u64(0)
    _lib.check(_lib.load().perf_network_param_count(grid.c(), mlp.c(), n))
    return int(n.value)


# ------------------------------------------------------------------ experiment configuration
class Conf(dict):
    """Attribute-style dict (stands in for the OmegaConf node the reference receives from Hydra)."""
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None         # hasattr / copy.deepcopy / pickle rely on AttributeError

    @staticmethod
    def wrap(v):
        if isinstance(v, dict):
            return Conf({k: Conf.wrap(x) for k, x in v.items()})
        if isinstance(v, list):
            return [Conf.wrap(x) for x in v]
        return v


def _coerce(text: str):
    import yaml
    v = yaml.safe_load(text)
    if isinstance(v, str):
        try:
            return float(v)                      # YAML 1.1 reads "1e-2" as a string; Hydra reads a float
        except ValueError:
            return v
    return v


import re as _re

_YAML11_MISSED_FLOAT = _re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)[eE][+-]?\d+$")


def _fix_floats(node):
    if isinstance(node, dict):
        return {k: _fix_floats(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_fix_floats(v) for v in node]
    if isinstance(node, str) and _YAML11_MISSED_FLOAT.match(node):
        # PyYAML (YAML 1.1) reads "1e-2" as a string because it has no dot; OmegaConf / YAML 1.2 read a float.  Only that
        # shape is converted: "007", "1_000", "nan", dates or names stay what the file said
        return float(node)
    return node


def load_config(config_dir: str, config_name: str = "nerf", overrides=()) -> Conf:
    """Read PeRF's Hydra YAML (`/root/reference/configs/nerf.yaml` + its ``defaults`` list, e.g.
    ``device: local`` -> ``configs/device/local.yaml``) without hydra, then apply dotted
    ``key=value`` overrides exactly like the reference's command line
    (`/root/reference/core_exp_runner.py:259`)."""
    import os
    import yaml
    with open(os.path.join(config_dir, config_name + ".yaml")) as f:
        root = yaml.safe_load(f)
    merged = {}
    for item in root.pop("defaults", []):
        if item == "_self_":
            continue
        for group, choice in item.items():
            with open(os.path.join(config_dir, group, str(choice) + ".yaml")) as f:
                merged[group] = yaml.safe_load(f)
    merged.update(root)
    merged = _fix_floats(merged)
    for ov in overrides:
        key, _, val = ov.partition("=")
        node = merged
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = _coerce(val)
    return Conf.wrap(merged)
