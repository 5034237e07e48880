"""Mirror of mlx_nerf/sampling/linear_disparity.py."""
import torch

from .. import ops
from .uniform import _as_col


def sample_z(near, far, n_samples: int):
    """linear_disparity.sample_z (sampling/linear_disparity.py:8-19), AS WRITTEN in the reference:
    1/(1/(near(1-t)) + 1/(far t)) -- the end points evaluate to 0 through +-inf (reference quirk, replicated)."""
    scalar = not isinstance(near, torch.Tensor) and not isinstance(far, torch.Tensor)
    near_t = _as_col(near, far)
    far_t = _as_col(far, near_t)
    near_t, far_t = torch.broadcast_tensors(near_t, far_t)
    z = ops.sample_z(near_t, far_t, n_samples, lindisp=True)
    return z[0] if scalar else z.reshape(*near_t.shape[:-1], n_samples)
