"""Oracle (TEST INFRASTRUCTURE): volume rendering in torch-CPU fp32.
Follows mlx_nerf/rendering/render.py:20-345, mlx_nerf/rendering/ray.py:7-70 and
mlx_nerf/ops/pose.py:7-58 of the reference.
"""
import numpy as np
import torch

from . import sampling as osamp
from . import models as omodels


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, noise=None):
    """raw2outputs (render.py:20-96).  Quirks replicated: transmittance = exp(-exclusive_cumsum(tau))
    with the RAW tau while alpha uses relu(tau); rgb used raw (no sigmoid); last delta = 1e10;
    weights keeps a trailing unit dim.  `noise` (N(0,1), [B,n]) is explicit when raw_noise_std>0."""
    raw, z_vals, rays_d = _t(raw), _t(z_vals), _t(rays_d)
    raw_rgb = raw[..., :3]
    sigma = raw[..., 3]
    if raw_noise_std > 0.0:
        sigma = sigma + _t(noise) * raw_noise_std
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    limit = torch.full((z_vals.shape[0], 1), 1e10, dtype=torch.float32)
    dists = torch.cat([dists, limit], dim=-1)
    dists = dists * torch.sqrt(torch.sum(rays_d[..., None, :] * rays_d[..., None, :], dim=-1))
    tau = (dists * sigma)[..., None]  # [B, n, 1]
    alphas = 1.0 - torch.exp(-torch.relu(tau))
    trans = torch.cumsum(tau[..., :-1, :], dim=-2)
    trans = torch.cat([torch.zeros((trans.shape[0], 1, 1), dtype=torch.float32), trans], dim=-2)
    trans = torch.exp(-trans)
    weights = alphas * trans
    rgb_map = torch.sum(weights * raw_rgb, dim=-2)
    depth_map = torch.sum(weights[..., 0] * z_vals, dim=-1)[..., None]
    acc_map = torch.sum(weights, dim=-2)
    disp_map = 1.0 / torch.maximum(1e-10 * torch.ones_like(depth_map), depth_map / acc_map)
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map)
    return rgb_map, disp_map, acc_map, weights, depth_map


def decompose_ray_batch(rays):
    """render.py:98-110."""
    rays = _t(rays)
    rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
    near, far = rays[:, 6:7], rays[:, 7:8]
    viewdirs = rays[:, -3:]
    return rays_o, rays_d, near, far, viewdirs


def _coarse_z(near, far, n, lindisp, perturb, t_rand):
    f = osamp.sample_z_lindisp if lindisp else osamp.sample_z_uniform
    z = f(near.numpy(), far.numpy(), n)
    z = osamp.add_noise_z(z, t_rand, float(perturb))
    return torch.from_numpy(np.ascontiguousarray(z))


def render_rays(rays, network_coarse, network_query_fn, n_depth_samples, retraw=False, lindisp=False,
                perturb=0.0, N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0,
                t_rand=None, **kwargs):
    """render_rays (render.py:112-162): coarse pass only; rgb_map IS the coarse result."""
    rays_o, rays_d, near, far, viewdirs = decompose_ray_batch(rays)
    z_vals = _coarse_z(near, far, n_depth_samples, lindisp, perturb, t_rand)
    pos = rays_o[..., None, :] + z_vals[..., :, None] * rays_d[..., None, :]
    raw = network_query_fn(pos, viewdirs, network_coarse)
    ret = {}
    if retraw:
        ret["raw"] = raw
    rgb, disp, acc, weights, depth = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd)
    ret.update(rgb_map=rgb, disp_map=disp, acc_map=acc, rgb_coarse=rgb, disp_coarse=disp,
               acc_coarse=acc, z_vals=z_vals, weights=weights)
    return ret


def render_rays_eval(rays, network_coarse, network_query_fn, n_depth_samples, retraw=False,
                     lindisp=False, perturb=0.0, N_importance=0, network_fine=None, white_bkgd=False,
                     raw_noise_std=0.0, t_rand=None, u_vals=None, **kwargs):
    """render_rays_eval (render.py:164-241): coarse -> inverse-CDF resample -> sort-merge -> fine
    (or coarse if network_fine is None).  z_vals / weights returned are the COARSE ones."""
    rays_o, rays_d, near, far, viewdirs = decompose_ray_batch(rays)
    ret = render_rays(rays, network_coarse, network_query_fn, n_depth_samples, retraw, lindisp, perturb,
                      N_importance, network_fine, white_bkgd, raw_noise_std, t_rand)
    z_vals, weights = ret["z_vals"], ret["weights"]
    z_imp = osamp.sample_pdf(z_vals.detach().numpy(), weights.detach().numpy(), u_vals)
    z_all = torch.from_numpy(osamp.merge_sorted(z_vals.detach().numpy(), z_imp))
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
    run_fn = network_fine if network_fine is not None else network_coarse
    raw = network_query_fn(pts, viewdirs, run_fn)
    rgb, disp, acc, _, _ = raw2outputs(raw, z_all, rays_d, raw_noise_std, white_bkgd)
    ret.update(rgb_map=rgb, disp_map=disp, acc_map=acc)
    ret["z_vals_fine"] = z_all  # extra (not in the reference dict) for parity tests
    return ret


def batchify_rays(rays, chunk=1024 * 32, **kwargs):
    """render.py:243-266."""
    fn = kwargs["render_rays_func"]
    acc = {}
    u_all = kwargs.pop("u_vals", None)
    for i in range(0, rays.shape[0], chunk):
        kw = dict(kwargs)
        if u_all is not None:
            kw["u_vals"] = u_all[i:i + chunk]
        res = fn(rays[i:i + chunk], **kw)
        for k, v in res.items():
            acc.setdefault(k, []).append(v)
    return {k: torch.cat(v, dim=0) for k, v in acc.items()}


def get_rays(H, W, K, c2w):
    """ray.get_rays (ray.py:7-35), NumPy, as written (fp32 meshgrid; K may be float64)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    fx, fy, cx, cy = K[0][0], K[1][1], K[0][2], K[1][2]
    dirs = np.stack([(i - cx) / fx, -(j - cy) / fy, -np.ones_like(i)], axis=-1)
    c2w = np.asarray(c2w)
    rays_d = np.sum(dirs[..., None, :] * c2w[:3, :3], axis=-1)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    return rays_o, rays_d


def pose_spherical(theta, phi, radius):
    """ops/pose.py:7-58 -> 4x4 c2w (fp32)."""
    def trans(r):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, r], [0, 0, 0, 1]], dtype=np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], dtype=np.float32)

    def rot_theta(t):
        return np.array([[np.cos(t), 0, -np.sin(t), 0], [0, 1, 0, 0], [np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1]], dtype=np.float32)

    pose = trans(radius)
    pose = rot_phi(phi / 180.0 * np.pi) @ pose
    pose = rot_theta(theta / 180.0 * np.pi) @ pose
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return (flip @ pose).astype(np.float32)


def build_rays(H, W, K, c2w, near, far, use_viewdirs=True):
    """Ray assembly of render() (render.py:283-328) with ndc=False: [o, d, near, far, viewdirs]."""
    rays_o, rays_d = get_rays(H, W, K, c2w)
    rays_o = np.reshape(rays_o, [-1, 3]).astype(np.float32)
    rays_d = np.reshape(rays_d, [-1, 3]).astype(np.float32)
    cols = [rays_o, rays_d, near * np.ones_like(rays_d[..., :1]), far * np.ones_like(rays_d[..., :1])]
    if use_viewdirs:
        vd = rays_d / np.sqrt(np.sum(rays_d * rays_d, axis=-1, keepdims=True))
        cols.append(vd.astype(np.float32))
    return np.concatenate(cols, axis=-1).astype(np.float32)


def render(H, W, K, chunk=1024 * 32, c2w=None, near=0.0, far=1.0, use_viewdirs=False, **kwargs):
    """render (render.py:268-345) for the ndc=False, c2w-given case the callers use."""
    rays = torch.from_numpy(build_rays(H, W, K, c2w, near, far, use_viewdirs))
    kwargs.pop("ndc", None)
    res = batchify_rays(rays, chunk, **kwargs)
    for k, v in res.items():
        res[k] = v.reshape([H, W] + list(v.shape[1:]))
    keys = ["rgb_map", "disp_map", "acc_map"]
    return [res[k] for k in keys] + [{k: v for k, v in res.items() if k not in keys}]


def make_query_fn(n_freqs_pos, n_freqs_dir, netchunk=64 * 1024):
    """network_query_fn closure of create_NeRF (NeRF.py:75-80)."""
    return lambda inputs, viewdirs, model: omodels.run_model(inputs, n_freqs_pos, viewdirs, n_freqs_dir, model, netchunk)
