"""Mirror of mlx_nerf/sampling/uniform.py."""
import torch

from .. import ops


def _as_col(v, like=None):
    if isinstance(v, torch.Tensor):
        return v
    dev = like.device if isinstance(like, torch.Tensor) else "cuda"
    return torch.tensor([[float(v)]], dtype=torch.float32, device=dev)


def sample_z(near, far, n_samples: int):
    """uniform.sample_z (sampling/uniform.py:7-18): z = near*(1-t) + far*t, t = linspace(0,1,n).
    near/far: [B,1] tensors (as render_rays passes them) -> [B, n]; python floats -> [n]."""
    scalar = not isinstance(near, torch.Tensor) and not isinstance(far, torch.Tensor)
    near_t = _as_col(near, far)
    far_t = _as_col(far, near_t)
    near_t, far_t = torch.broadcast_tensors(near_t, far_t)
    z = ops.sample_z(near_t, far_t, n_samples, lindisp=False)
    return z[0] if scalar else z.reshape(*near_t.shape[:-1], n_samples)
