"""Mirror of mlx_nerf/rendering/render.py on CUDA tensors.

Execution flow (as the reference): render -> batchify_rays -> render_rays[_eval] -> raw2outputs."""
import torch

from .. import ops, sampling
from ..ops import library as _library  # noqa: F401  (registers torch.ops.nmx.*)
from ..sampling import uniform, linear_disparity
from . import ray


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, noise=None):
    """raw2outputs (rendering/render.py:20-96) -> (rgb_map [B,3], disp_map [B,1], acc_map [B,1], weights [B,n,1],
    depth_map [B,1]).  Reference quirks kept: T = exp(-exclusive_cumsum(tau)) with the raw tau, alpha with
    relu(tau); rgb used raw (no sigmoid); last delta 1e10."""
    raw = raw.float().contiguous()
    z_vals = z_vals.float().contiguous()
    rays_d = rays_d.float().contiguous()
    if raw_noise_std > 0.0 and noise is None:
        noise = torch.randn(raw.shape[:-1], device=raw.device)  # render.py:41-43
    if raw_noise_std <= 0.0:
        noise = None
    # one fused kernel each way (K4), as the custom operator torch.ops.nmx.composite_fwd (backward: nmx.composite_bwd)
    return torch.ops.nmx.composite_fwd(raw[..., :4].contiguous() if raw.shape[-1] != 4 else raw, z_vals, rays_d, noise,
                                       float(raw_noise_std), bool(white_bkgd))


def decompose_ray_batch(rays_batch_linear, is_time_included: bool = False):
    """decompose_ray_batch (rendering/render.py:98-110)."""
    rays_o, rays_d = rays_batch_linear[:, 0:3], rays_batch_linear[:, 3:6]
    k = int(is_time_included)
    bounds = rays_batch_linear[..., 6:8 + k].reshape(-1, 1, 2 + k)
    near, far = bounds[..., 0], bounds[..., 1]
    frame_time = bounds[..., 2] if is_time_included else None
    viewdirs = rays_batch_linear[:, -3:]
    return rays_o, rays_d, near, far, viewdirs, frame_time


def _query(network_query_fn, model, rays_batch_linear, rays_o, rays_d, viewdirs, z_vals):
    """pos = o + z*d then network_query_fn(pos, viewdirs, model) (render.py:142-144).  When the closure comes from
    create_NeRF (it carries `fused_embedders`) the positional encodings are generated inside the MLP's operand
    producer from (rays, z) and never written to memory; arithmetic is identical."""
    fused = getattr(network_query_fn, "fused_embedders", None)
    if fused is not None and hasattr(model, "forward_rays") and rays_batch_linear.shape[-1] >= 9 \
            and model._cfg.n_freqs_pos == fused[0] and fused[1] is not None and model.is_use_view_directions:
        return model.forward_rays(rays_batch_linear, z_vals)
    pos = ops.ray_points(rays_batch_linear, z_vals)
    return network_query_fn(pos, viewdirs, model)


def _coarse_z(near, far, n_depth_samples, lindisp, perturb, t_rand=None):
    if not lindisp:
        z_vals = uniform.sample_z(near, far, n_depth_samples)
    else:
        z_vals = linear_disparity.sample_z(near, far, n_depth_samples)
    return sampling.add_noise_z(z_vals, float(perturb), t_rand)


def render_rays(rays_batch_linear, network_coarse, network_query_fn, n_depth_samples, retraw=False, lindisp=False,
                perturb=0.0, N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0.0, verbose=False,
                pytest=False, **kwargs):
    """render_rays (rendering/render.py:112-162): COARSE pass only; rgb_map is the coarse result."""
    rays_batch_linear = rays_batch_linear.float().contiguous()
    rays_o, rays_d, near, far, viewdirs, _ = decompose_ray_batch(rays_batch_linear)
    z_vals = _coarse_z(near, far, n_depth_samples, lindisp, perturb, kwargs.get("t_rand"))
    raw = _query(network_query_fn, network_coarse, rays_batch_linear, rays_o, rays_d, viewdirs, z_vals)
    ret = {}
    if retraw:
        ret["raw"] = raw
    rgb_coarse, disp_coarse, acc_coarse, weights, depth_map = raw2outputs(raw, z_vals, rays_d.contiguous(),
                                                                           raw_noise_std, white_bkgd, pytest)
    ret["rgb_map"] = rgb_coarse
    ret["disp_map"] = disp_coarse
    ret["acc_map"] = acc_coarse
    ret["rgb_coarse"] = rgb_coarse
    ret["disp_coarse"] = disp_coarse
    ret["acc_coarse"] = acc_coarse
    ret["z_vals"] = z_vals
    ret["weights"] = weights
    return ret


def render_rays_eval(rays_batch_linear, network_coarse, network_query_fn, n_depth_samples, retraw=False,
                     lindisp=False, perturb=0.0, N_importance=0, network_fine=None, white_bkgd=False,
                     raw_noise_std=0.0, verbose=False, pytest=False, **kwargs):
    """render_rays_eval (rendering/render.py:164-241): coarse -> inverse-CDF resample (detached) -> sort-merge ->
    fine (or coarse when network_fine is None).  The reference's device->host->device round trip (:215-223) is
    replaced by one on-device kernel.  z_vals / weights in the result stay the COARSE ones."""
    rays_batch_linear = rays_batch_linear.float().contiguous()
    rays_o, rays_d, near, far, viewdirs, _ = decompose_ray_batch(rays_batch_linear)
    ret = render_rays(rays_batch_linear, network_coarse, network_query_fn, n_depth_samples, retraw, lindisp, perturb,
                      N_importance, network_fine, white_bkgd, raw_noise_std, verbose, pytest, **kwargs)
    z_vals, weights = ret["z_vals"], ret["weights"]
    with torch.no_grad():
        z_all, _ = sampling.sample_and_merge(z_vals.detach(), weights.detach(), N_importance, u_vals=kwargs.get("u_vals"))
    run_fn = network_fine if network_fine else network_coarse
    raw = _query(network_query_fn, run_fn, rays_batch_linear, rays_o, rays_d, viewdirs, z_all)
    rgb, disp, acc, weight, depth = raw2outputs(raw, z_all, rays_d.contiguous(), raw_noise_std, white_bkgd)
    ret["rgb_map"] = rgb
    ret["disp_map"] = disp
    ret["acc_map"] = acc
    return ret


def batchify_rays(rays_linear, chunk=1024 * 32, **kwargs):
    """batchify_rays (rendering/render.py:243-266)."""
    render_rays_func = kwargs["render_rays_func"]
    u_all = kwargs.pop("u_vals", None)
    results_batched = {}
    for i in range(0, rays_linear.shape[0], chunk):
        kw = kwargs if u_all is None else dict(kwargs, u_vals=u_all[i:i + chunk])
        results = render_rays_func(rays_linear[i:i + chunk], **kw)
        for key, val in results.items():
            results_batched.setdefault(key, []).append(val)
    return {key: torch.cat(val, dim=0) for key, val in results_batched.items()}


def build_rays(H, W, K, c2w, near, far, use_viewdirs=False, ndc=False, c2w_staticcam=None, rays=None, device="cuda"):
    """Ray assembly of render() (rendering/render.py:283-328): [o, d, near, far (, viewdirs)] fp32 [H*W, 8|11]."""
    if c2w is not None and not ndc and c2w_staticcam is None:
        # the case every caller uses (NeRF.py:146-148): one kernel from (K, c2w) to the assembled ray buffer
        rays_linear = ops.gen_rays(H, W, K, torch.as_tensor(c2w).to(device), None, near, far, 11 if use_viewdirs else 8)
        return rays_linear, (H, W, 3)
    if c2w is None and rays is not None:
        rays_o, rays_d = rays
    else:
        rays_o, rays_d = ray.get_rays(H, W, K, c2w, device=device)
    rays_shape = rays_d.shape
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = ray.get_rays(H, W, K, c2w_staticcam, device=device)
            rays_o = rays_o.reshape(-1, 3).float()
            rays_d = rays_d.reshape(-1, 3).float()
        viewdirs = viewdirs / torch.linalg.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = viewdirs.reshape(-1, 3).float()
    if ndc:
        rays_o, rays_d = ray.ndc_rays(H, W, K[0][0], 1.0, rays_o, rays_d)
    near_t = near * torch.ones_like(rays_d[..., :1])
    far_t = far * torch.ones_like(rays_d[..., :1])
    cols = [rays_o, rays_d, near_t, far_t]
    if use_viewdirs:
        cols.append(viewdirs)
    return torch.cat(cols, dim=-1).contiguous(), rays_shape


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0.0, far=1.0, use_viewdirs=False,
           c2w_staticcam=None, process_group=None, **kwargs):
    """render (rendering/render.py:268-345) -> [rgb_map, disp_map, acc_map, extras_dict], each reshaped to [H, W, .].

    `process_group` (not in the reference, which is single-device): shard the frame's [H*W, 11] ray buffer into one
    contiguous ray tile per rank, render the local tile with the same chunk loop, and all_gather the
    per-ray results, so every rank returns the same assembled frame as a single-GPU call -- rays are independent, so
    the assembled frame is bit-identical to it.  No collective runs inside the chunk loop."""
    device = kwargs.pop("device", "cuda")
    rays_linear, rays_shape = build_rays(H, W, K, c2w, near, far, use_viewdirs, ndc, c2w_staticcam, rays, device)
    if process_group is None:
        results_batched = batchify_rays(rays_linear, chunk, **kwargs)
    else:
        import torch.distributed as dist
        from ..parallel import gather_tiles, ray_tile
        n_rays = rays_linear.shape[0]
        start, stop = ray_tile(n_rays, dist.get_rank(process_group), dist.get_world_size(process_group))
        if kwargs.get("u_vals") is not None:
            kwargs = dict(kwargs, u_vals=kwargs["u_vals"][start:stop])
        if stop > start:
            local = batchify_rays(rays_linear[start:stop], chunk, **kwargs)
        else:  # more ranks than rays: an empty tile with the right keys / trailing shapes
            if kwargs.get("u_vals") is not None:
                kwargs = dict(kwargs, u_vals=None)
            local = {k: v[:0] for k, v in batchify_rays(rays_linear[:1], chunk, **kwargs).items()}
        results_batched = {k: gather_tiles(v.contiguous(), n_rays, process_group) for k, v in local.items()}
    for key, val in results_batched.items():
        results_batched[key] = val.reshape(tuple(list(rays_shape[:-1]) + list(val.shape[1:])))
    k_extract = ["rgb_map", "disp_map", "acc_map"]
    ret_list = [results_batched[k] for k in k_extract]
    ret_dict = {k: v for k, v in results_batched.items() if k not in k_extract}
    return ret_list + [ret_dict]
