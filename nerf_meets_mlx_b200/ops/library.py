"""The hot-path ops as PyTorch custom operators (`torch.ops.nmx.*`), the boundary BASELINE.json's north_star names:
"Python host code drives the work through PyTorch custom ops over a thin C-ABI layer into hand-written sm_100a CUDA
kernels".  Each operator
  * calls one C-ABI entry of libnmx.so (include/nmx.h) through the tensor-level wrappers of `nerf_meets_mlx_b200.ops`,
  * has a FAKE (meta) kernel, so shape propagation / `torch.compile` tracing / `torch.library.opcheck` work without
    running CUDA code,
  * where the reference differentiates through it (raw2outputs, the hash grid), has its backward registered as another
    custom operator call -- no Python autograd.Function in between.
There is no CPU implementation: a CPU tensor reaches `require_cuda` and raises."""
from typing import Optional, Tuple

import torch
from torch.library import custom_op

from .. import ops as _ops

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------- sampling (K1, K5)
@custom_op("nmx::sample_z", mutates_args=())
def sample_z(near: Tensor, far: Tensor, n_samples: int, lindisp: bool = False) -> Tensor:
    """uniform.sample_z / linear_disparity.sample_z (sampling/uniform.py:7-18, linear_disparity.py:8-19)."""
    return _ops.sample_z(near, far, n_samples, lindisp)


@sample_z.register_fake
def _(near, far, n_samples, lindisp=False):
    return near.new_empty((near.numel(), n_samples), dtype=torch.float32)


@custom_op("nmx::add_noise_z", mutates_args=())
def add_noise_z(z_vals: Tensor, t_rand: Tensor, strength: float) -> Tensor:
    """add_noise_z (sampling/__init__.py:10-31) with the uniform draw explicit."""
    return _ops.add_noise_z(z_vals, t_rand, strength)


@add_noise_z.register_fake
def _(z_vals, t_rand, strength):
    return torch.empty_like(z_vals, dtype=torch.float32)


@custom_op("nmx::sample_pdf", mutates_args=())
def sample_pdf(z: Tensor, weights: Tensor, u: Tensor, eps: float = 1e-5) -> Tuple[Tensor, Tensor]:
    """sample_from_inverse_cdf_torch (sampling/__init__.py:101-178) + sort-merge (render.py:225):
    -> (z_imp [B, N] unsorted, z_merged [B, n + N] ascending).  Detached by construction (the reference runs it under
    no_grad on host copies)."""
    r = _ops.sample_pdf(z, weights, u, eps)
    return r["z_imp"], r["z_merged"]


@sample_pdf.register_fake
def _(z, weights, u, eps=1e-5):
    B, n = z.shape
    return z.new_empty((B, u.shape[-1])), z.new_empty((B, n + u.shape[-1]))


# ----------------------------------------------------------------------------------------- encodings (K2)
@custom_op("nmx::pe_embedder", mutates_args=())
def pe_embedder(x: Tensor, n_freqs: int, include_input: bool = True) -> Tensor:
    """Embedder.embed (models/embedding.py:35-71)."""
    return _ops.pe_embedder(x, n_freqs, include_input)


@pe_embedder.register_fake
def _(x, n_freqs, include_input=True):
    d = x.shape[-1]
    return x.new_empty(tuple(x.shape[:-1]) + ((d if include_input else 0) + 2 * d * n_freqs,), dtype=torch.float32)


@custom_op("nmx::pe_sinusoidal", mutates_args=())
def pe_sinusoidal(x: Tensor, bands: Tensor, include_input: bool = False) -> Tensor:
    """SinusoidalEncoding.__call__ (encoding/sinusoidal.py:39-66)."""
    return _ops.pe_sinusoidal(x, bands, include_input)


@pe_sinusoidal.register_fake
def _(x, bands, include_input=False):
    d = x.shape[-1]
    return x.new_empty((x.numel() // d, 2 * d * bands.numel() + (d if include_input else 0)), dtype=torch.float32)


@custom_op("nmx::sh_encode", mutates_args=())
def sh_encode(dirs: Tensor, n_degrees: int) -> Tensor:
    """SphericalHarmonicsEncoding.__call__ (encoding/spherical_harmonics.py:33-94)."""
    return _ops.sh_encode(dirs, n_degrees)


@sh_encode.register_fake
def _(dirs, n_degrees):
    return dirs.new_empty(tuple(dirs.shape[:-1]) + ((n_degrees + 1) ** 2,), dtype=torch.float32)


@custom_op("nmx::hashgrid_fwd", mutates_args=())
def hashgrid_fwd(x: Tensor, tables: Tensor, scaled_res: Tensor, log2_T: int) -> Tensor:
    """MultiHashEncoding.__call__ (encoding/multi_hash.py:79-137), canonical per-level semantics."""
    return _ops.hashgrid_fwd(x, tables, scaled_res, log2_T)


@hashgrid_fwd.register_fake
def _(x, tables, scaled_res, log2_T):
    return x.new_empty((x.shape[0], tables.shape[0] * tables.shape[2]), dtype=torch.float32)


@custom_op("nmx::hashgrid_bwd", mutates_args=())
def hashgrid_bwd(x: Tensor, scaled_res: Tensor, d_out: Tensor, L: int, F: int, log2_T: int) -> Tensor:
    """Gradient of hashgrid_fwd w.r.t. the tables (vector-atomic scatter)."""
    return _ops.hashgrid_bwd(x, scaled_res, d_out, L, F, log2_T)


@hashgrid_bwd.register_fake
def _(x, scaled_res, d_out, L, F, log2_T):
    return x.new_empty((L, 1 << log2_T, F), dtype=torch.float32)


def _hashgrid_setup(ctx, inputs, output):
    x, tables, scaled_res, log2_T = inputs
    ctx.save_for_backward(x, scaled_res)
    ctx.shape, ctx.log2_T = tuple(tables.shape), log2_T


def _hashgrid_backward(ctx, d_out):
    x, scaled_res = ctx.saved_tensors
    L, _, F = ctx.shape
    return None, hashgrid_bwd(x, scaled_res, d_out.contiguous(), L, F, ctx.log2_T), None, None


hashgrid_fwd.register_autograd(_hashgrid_backward, setup_context=_hashgrid_setup)


# ----------------------------------------------------------------------------------------- compositing (K4)
@custom_op("nmx::composite_fwd", mutates_args=())
def composite_fwd(raw: Tensor, z: Tensor, rays_d: Tensor, noise: Optional[Tensor] = None, raw_noise_std: float = 0.0,
                  white_bkgd: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """raw2outputs (rendering/render.py:20-96) -> (rgb [B,3], disp [B,1], acc [B,1], weights [B,n,1], depth [B,1])."""
    return _ops.composite_fwd(raw, z, rays_d, noise, raw_noise_std, white_bkgd)


@composite_fwd.register_fake
def _(raw, z, rays_d, noise=None, raw_noise_std=0.0, white_bkgd=False):
    B, n = z.shape
    e = lambda *s: z.new_empty(s, dtype=torch.float32)
    return e(B, 3), e(B, 1), e(B, 1), e(B, n, 1), e(B, 1)


@custom_op("nmx::composite_bwd", mutates_args=())
def composite_bwd(raw: Tensor, z: Tensor, rays_d: Tensor, d_rgb: Tensor, d_disp: Optional[Tensor] = None,
                  d_acc: Optional[Tensor] = None, d_depth: Optional[Tensor] = None, d_weights: Optional[Tensor] = None,
                  noise: Optional[Tensor] = None, raw_noise_std: float = 0.0, white_bkgd: bool = False) -> Tensor:
    """Gradient of raw2outputs w.r.t. raw [B, n, 4] (no gradient w.r.t. z is ever needed: z is a detached input)."""
    return _ops.composite_bwd(raw, z, rays_d, d_rgb, d_disp, d_acc, d_depth, d_weights, noise, raw_noise_std, white_bkgd)


@composite_bwd.register_fake
def _(raw, z, rays_d, d_rgb, d_disp=None, d_acc=None, d_depth=None, d_weights=None, noise=None, raw_noise_std=0.0,
      white_bkgd=False):
    return torch.empty_like(raw, dtype=torch.float32)


def _composite_setup(ctx, inputs, output):
    raw, z, rays_d, noise, std, wb = inputs
    ctx.save_for_backward(raw, z, rays_d, noise)
    ctx.cfg = (std, wb)


def _composite_backward(ctx, d_rgb, d_disp, d_acc, d_weights, d_depth):
    raw, z, rays_d, noise = ctx.saved_tensors
    std, wb = ctx.cfg
    opt = lambda t: None if t is None else t.contiguous()
    if d_rgb is None:
        d_rgb = raw.new_zeros((z.shape[0], 3))
    d_raw = composite_bwd(raw, z, rays_d, d_rgb.contiguous(), opt(d_disp), opt(d_acc), opt(d_depth), opt(d_weights),
                          noise, std, wb)
    return d_raw, None, None, None, None, None


composite_fwd.register_autograd(_composite_backward, setup_context=_composite_setup)


# ----------------------------------------------------------------------------------------- rays, loss, optimiser (K6)
@custom_op("nmx::assemble_rays", mutates_args=())
def assemble_rays(rays_o: Tensor, rays_d: Tensor, near: float, far: float) -> Tensor:
    """[o, d, near, far, d/||d||] rows (__test_nerf.py:57-82)."""
    return _ops.assemble_rays(rays_o, rays_d, near, far)


@assemble_rays.register_fake
def _(rays_o, rays_d, near, far):
    return rays_o.new_empty((rays_o.numel() // 3, 11), dtype=torch.float32)


@custom_op("nmx::mse_fwd_bwd", mutates_args=())
def mse_fwd_bwd(pred: Tensor, target: Tensor) -> Tuple[Tensor, Tensor]:
    """mean((pred - target)^2) and its gradient w.r.t. pred (__test_nerf.py:88,124)."""
    return _ops.mse_fwd_bwd(pred, target)


@mse_fwd_bwd.register_fake
def _(pred, target):
    return pred.new_empty((1,), dtype=torch.float32), torch.empty_like(pred, dtype=torch.float32)


@custom_op("nmx::adam_step", mutates_args=("p", "m", "v"))
def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, b1: float = 0.9, b2: float = 0.999,
              eps: float = 1e-8) -> None:
    """optim.Adam of MLX 0.7.0 (models/NeRF.py:120), no bias correction, in place."""
    _ops.adam_step(p, g, m, v, lr, b1, b2, eps)


# ----------------------------------------------------------------------------------------- NeRF MLP (K3)
@custom_op("nmx::mlp_fwd", mutates_args=("workspace",))
def mlp_fwd(plan: int, workspace: Tensor, params: Tensor, enc_kind: int, x_or_rays: Tensor, z: Optional[Tensor],
            bands: Optional[Tensor], B: int, n: int, out_cols: int, save: bool) -> Tensor:
    """NeRF.forward via run_model (models/NeRF.py:25-48, 201-243) on an nmx_mlp_plan handle (`plan` = its address).
    enc_kind as in include/nmx.h; save=True keeps the activations nmx::mlp_bwd needs in `workspace`."""
    from .._lib_loader import call, i32, i64, ptr, require_cuda, stream
    import ctypes
    require_cuda(workspace, params, x_or_rays, z, bands)
    out = torch.empty((B * n, out_cols), dtype=torch.float32, device=params.device)
    stride = x_or_rays.shape[-1] if enc_kind == 1 else 0
    call("nmx_mlp_fwd", ctypes.c_void_p(plan), ptr(workspace), ptr(params), i32(enc_kind), ptr(x_or_rays), i32(stride),
         ptr(z), ptr(bands), ptr(out), i64(B), i32(n), i32(1 if save else 0), stream())
    return out


@mlp_fwd.register_fake
def _(plan, workspace, params, enc_kind, x_or_rays, z, bands, B, n, out_cols, save):
    return params.new_empty((B * n, out_cols), dtype=torch.float32)


@custom_op("nmx::mlp_bwd", mutates_args=("workspace",))
def mlp_bwd(plan: int, workspace: Tensor, params: Tensor, d_out: Tensor, P: int) -> Tensor:
    """Gradient of all parameters (packed like `params`) from d_out [P, out_cols], using the activations saved by the
    preceding nmx::mlp_fwd(save=True) on the same workspace."""
    from .._lib_loader import call, i64, ptr, require_cuda, stream
    import ctypes
    require_cuda(workspace, params, d_out)
    g = torch.empty_like(params)
    call("nmx_mlp_bwd", ctypes.c_void_p(plan), ptr(workspace), ptr(params), ptr(d_out), ptr(g), i64(P), stream())
    return g


@mlp_bwd.register_fake
def _(plan, workspace, params, d_out, P):
    return torch.empty_like(params)


ALL_OPS = ("sample_z", "add_noise_z", "sample_pdf", "pe_embedder", "pe_sinusoidal", "sh_encode", "hashgrid_fwd",
           "hashgrid_bwd", "composite_fwd", "composite_bwd", "assemble_rays", "mse_fwd_bwd", "adam_step", "mlp_fwd",
           "mlp_bwd")
