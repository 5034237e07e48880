"""Mirror of mlx_nerf/rendering/ray.py (camera -> rays).  O(rays) elementwise prep that feeds the hot path; on the
device so no host round trip precedes the kernels."""
import numpy as np
import torch

from .. import ops


def get_rays(H: int, W: int, K, c2w, device=None):
    """get_rays (rendering/ray.py:7-35): pinhole rays; returns (rays_o, rays_d) [H, W, 3] (nmx_gen_rays; the direction
    is evaluated in float64 and cast to fp32, which is what the reference does with its float64 K)."""
    if device is None:
        device = c2w.device if isinstance(c2w, torch.Tensor) and c2w.is_cuda else "cuda"
    c2w = torch.as_tensor(np.asarray(c2w, dtype=np.float32) if not isinstance(c2w, torch.Tensor) else c2w)
    rays = ops.gen_rays(H, W, K, c2w.to(device=device, dtype=torch.float32), None, 0.0, 1.0, 6)
    return rays[:, 0:3].reshape(H, W, 3), rays[:, 3:6].reshape(H, W, 3)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """ndc_rays (rendering/ray.py:39-70)."""
    t_n = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t_n[..., None] * rays_d
    o_x, o_y, o_z = rays_o[..., 0], rays_o[..., 1], rays_o[..., 2]
    o0 = (-focal / (0.5 * W)) * (o_x / o_z)
    o1 = (-focal / (0.5 * H)) * (o_y / o_z)
    o2 = (1.0 + 2.0 * near / o_z)
    d_x, d_y, d_z = rays_d[..., 0], rays_d[..., 1], rays_d[..., 2]
    d0 = (-focal / (0.5 * W)) * (d_x / d_z - o_x / o_z)
    d1 = (-focal / (0.5 * H)) * (d_y / d_z - o_y / o_z)
    d2 = -2.0 * near * (1.0 / o_z)
    return torch.stack([o0, o1, o2], dim=-1), torch.stack([d0, d1, d2], dim=-1)
