"""Drop-in plugin modules (see README.md in this directory)."""
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))


def install() -> None:
    """Make ``import tinycudann`` / ``import nerfacc`` / ``import torch_efficient_distloss`` resolve to
    the perf_b200 implementations (prepends this directory to sys.path and drops cached imports)."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    for name in [m for m in sys.modules if m.split(".")[0] in ("tinycudann", "nerfacc", "torch_efficient_distloss")]:
        if not getattr(sys.modules[name], "__perf_b200_shim__", False):
            del sys.modules[name]
