"""Import-only stand-in: PeRF imports ``PropNetEstimator`` but runs ``estimator_type: occ``
(`/root/reference/configs/nerf.yaml:25`); its proposal renderer is broken upstream
(`nerf_renderer.py:73`)."""
import torch

__perf_b200_shim__ = True


class PropNetEstimator(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("perf_b200 nerfacc: PropNetEstimator is not implemented (PeRF uses the occupancy estimator)")
