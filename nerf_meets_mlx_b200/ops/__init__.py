"""Tensor-level wrappers over the C ABI (include/nmx.h): allocate outputs, pass raw device pointers and the
current CUDA stream.  No CPU path exists; every function raises if the extension is missing or inputs are not
contiguous CUDA tensors."""
import torch

from .._lib_loader import NmxError, call, f32, f64, i32, i64, ptr, require_cuda, stream


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------------------- sampling
def sample_z(near, far, n_samples, lindisp=False):
    near, far = _f32c(near).reshape(-1), _f32c(far).reshape(-1)
    require_cuda(near, far)
    B = near.numel()
    z = torch.empty((B, n_samples), dtype=torch.float32, device=near.device)
    call("nmx_sample_z_fwd", ptr(near), ptr(far), ptr(z), i64(B), i32(n_samples), i32(1 if lindisp else 0), stream())
    return z


def sample_z_rays(rays, n_samples, lindisp=False):
    """sample_z with near / far read from columns 6 / 7 of the assembled ray rows [B, >=8] (no column copies)."""
    rays = _f32c(rays)
    require_cuda(rays)
    B = rays.shape[0]
    z = torch.empty((B, n_samples), dtype=torch.float32, device=rays.device)
    call("nmx_sample_z_rays", ptr(rays), i32(rays.shape[1]), ptr(z), i64(B), i32(n_samples), i32(1 if lindisp else 0), stream())
    return z


def add_noise_z(z_vals, t_rand, strength):
    z = _f32c(z_vals)
    require_cuda(z)
    n = z.shape[-1]
    B = z.numel() // n
    t = _f32c(t_rand)
    out = torch.empty_like(z)
    call("nmx_add_noise_z_fwd", ptr(z), ptr(t), ptr(out), i64(B), i32(n), f32(strength), stream())
    return out


def ray_points(rays, z):
    rays, z = _f32c(rays), _f32c(z)
    require_cuda(rays, z)
    B, n = z.shape
    pos = torch.empty((B, n, 3), dtype=torch.float32, device=z.device)
    call("nmx_ray_points_fwd", ptr(rays), i32(rays.shape[1]), ptr(z), ptr(pos), i64(B), i32(n), stream())
    return pos


def assemble_rays(rays_o, rays_d, near, far):
    """[o, d, near, far, d/||d||] rows [B, 11] from explicit origins / directions (__test_nerf.py:57-82)."""
    o, d = _f32c(rays_o).reshape(-1, 3), _f32c(rays_d).reshape(-1, 3)
    require_cuda(o, d)
    B = o.shape[0]
    rays = torch.empty((B, 11), dtype=torch.float32, device=o.device)
    call("nmx_assemble_rays", ptr(o), ptr(d), i64(B), f32(near), f32(far), ptr(rays), stream())
    return rays


def _rows3(t):
    """A [B, >=3] fp32 CUDA matrix whose rows are contiguous (a column slice of the ray batch qualifies): the tensor
    and its row stride, without copying."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.stride(1) != 1:
        t = t.contiguous()
    if not t.is_cuda:
        raise NmxError("nerf_meets_mlx_b200 ops need CUDA tensors (no CPU fallback exists)")
    return t, int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1])


def gen_rays(H, W, K, c2w, pix=None, near=0.0, far=1.0, n_cols=11, image=None):
    """Pixel ids -> ray batch [B, n_cols] (o, d, near, far, viewdirs) and, with `image` [H*W, C>=3], the target
    pixels [B, 3] (ray.py:7-35 + render.py:283-328 / __test_nerf.py:208-236)."""
    import numpy as np
    c2w = torch.as_tensor(np.asarray(c2w, dtype=np.float32) if not isinstance(c2w, torch.Tensor) else c2w)
    c2w = _f32c(c2w.to("cuda") if not c2w.is_cuda else c2w)
    if c2w.ndim != 2 or c2w.shape[0] < 3 or c2w.shape[1] < 4:
        raise ValueError("c2w must be [3|4, 4]")
    B = H * W if pix is None else int(pix.numel())
    if pix is not None:
        pix = pix.to(device=c2w.device, dtype=torch.int32).contiguous()
    rays = torch.empty((B, n_cols), dtype=torch.float32, device=c2w.device)
    target = None
    img_ld = 0
    if image is not None:
        image = _f32c(image)
        require_cuda(image)
        img_ld = image.shape[-1]
        if image.numel() != H * W * img_ld:
            raise ValueError("image must be [H*W, C]")
        target = torch.empty((B, 3), dtype=torch.float32, device=c2w.device)
    call("nmx_gen_rays", ptr(c2w), i32(c2w.stride(0)), f64(K[0][0]), f64(K[1][1]), f64(K[0][2]), f64(K[1][2]),
         i32(H), i32(W), ptr(pix), i64(B), f32(near), f32(far), ptr(rays), i32(n_cols), ptr(image), i32(img_ld),
         ptr(target), stream())
    return rays if image is None else (rays, target)


# ----------------------------------------------------------------------------------------- encodings
def pe_embedder(x, n_freqs, include_input=True, bands=None):
    """bands: optional [n_freqs] fp32 frequencies (default: the reference's get_embedder bands k^2)"""
    x = _f32c(x)
    require_cuda(x)
    in_dim = x.shape[-1]
    P = x.numel() // in_dim
    out_dim = (in_dim if include_input else 0) + 2 * in_dim * n_freqs
    out = torch.empty((P, out_dim), dtype=torch.float32, device=x.device)
    if bands is None:
        call("nmx_pe_embedder_fwd", ptr(x), ptr(out), i64(P), i32(in_dim), i32(n_freqs), i32(1 if include_input else 0),
             stream())
    else:
        bands = _f32c(bands).to(x.device)
        assert bands.numel() == n_freqs, "bands must hold n_freqs values"
        call("nmx_pe_embedder_bands_fwd", ptr(x), ptr(bands), ptr(out), i64(P), i32(in_dim), i32(n_freqs),
             i32(1 if include_input else 0), stream())
    return out.reshape(*x.shape[:-1], out_dim)


def pe_sinusoidal(x, bands, include_input=False):
    x, bands = _f32c(x), _f32c(bands)
    require_cuda(x, bands)
    in_dim = x.shape[-1]
    P = x.numel() // in_dim
    nf = bands.numel()
    out_dim = 2 * in_dim * nf + (in_dim if include_input else 0)
    out = torch.empty((P, out_dim), dtype=torch.float32, device=x.device)
    call("nmx_pe_sinusoidal_fwd", ptr(x), ptr(bands), ptr(out), i64(P), i32(in_dim), i32(nf), i32(1 if include_input else 0), stream())
    return out


def sh_encode(dirs, n_degrees):
    d = _f32c(dirs)
    require_cuda(d)
    in_dim = d.shape[-1]
    B = d.numel() // in_dim
    od = (n_degrees + 1) ** 2
    out = torch.empty((B, od), dtype=torch.float32, device=d.device)
    call("nmx_sh_encode_fwd", ptr(d), i32(in_dim), ptr(out), i64(B), i32(n_degrees), stream())
    return out.reshape(*d.shape[:-1], od)


def hashgrid_hash(coords, log2_T):
    c = coords.to(torch.int32).contiguous()
    require_cuda(c)
    M = c.numel() // 3
    idx = torch.empty(c.shape[:-1], dtype=torch.int32, device=c.device)
    call("nmx_hashgrid_hash", ptr(c), ptr(idx), i64(M), i32(log2_T), stream())
    return idx


def hashgrid_fwd(x, tables, scaled_res, log2_T, return_idx=False):
    x, tables, scaled_res = _f32c(x), _f32c(tables), _f32c(scaled_res)
    require_cuda(x, tables, scaled_res)
    P = x.shape[0]
    L, T, F = tables.shape
    assert T == 1 << log2_T
    out = torch.empty((P, L * F), dtype=torch.float32, device=x.device)
    idx = torch.empty((P, L, 8), dtype=torch.int32, device=x.device) if return_idx else None
    call("nmx_hashgrid_fwd", ptr(x), ptr(tables), ptr(scaled_res), ptr(out), ptr(idx), i64(P), i32(L), i32(F), i32(log2_T), stream())
    return (out, idx) if return_idx else out


def hashgrid_bwd(x, scaled_res, d_out, L, F, log2_T):
    x, scaled_res, d_out = _f32c(x), _f32c(scaled_res), _f32c(d_out)
    require_cuda(x, scaled_res, d_out)
    P = x.shape[0]
    d_tables = torch.zeros((L, 1 << log2_T, F), dtype=torch.float32, device=x.device)
    call("nmx_hashgrid_bwd", ptr(x), ptr(scaled_res), ptr(d_out), ptr(d_tables), i64(P), i32(L), i32(F), i32(log2_T), stream())
    return d_tables


# ----------------------------------------------------------------------------------------- compositing
def composite_fwd(raw, z, rays_d, noise=None, raw_noise_std=0.0, white_bkgd=False):
    raw, z = _f32c(raw), _f32c(z)
    rays_d, d_stride = _rows3(rays_d)
    require_cuda(raw, z)
    B, n = z.shape
    assert raw.shape == (B, n, 4), f"raw must be [B, n, 4], got {tuple(raw.shape)}"
    dev = z.device
    rgb = torch.empty((B, 3), dtype=torch.float32, device=dev)
    disp = torch.empty((B, 1), dtype=torch.float32, device=dev)
    acc = torch.empty((B, 1), dtype=torch.float32, device=dev)
    weights = torch.empty((B, n, 1), dtype=torch.float32, device=dev)
    depth = torch.empty((B, 1), dtype=torch.float32, device=dev)
    nz = _f32c(noise) if (noise is not None and raw_noise_std > 0) else None
    call("nmx_composite_fwd", ptr(raw), ptr(z), ptr(rays_d), i32(d_stride), ptr(nz), f32(raw_noise_std),
         i32(1 if white_bkgd else 0), ptr(rgb), ptr(disp), ptr(acc), ptr(weights), ptr(depth), i64(B), i32(n), stream())
    return rgb, disp, acc, weights, depth


def composite_bwd(raw, z, rays_d, d_rgb, d_disp=None, d_acc=None, d_depth=None, d_weights=None, noise=None,
                  raw_noise_std=0.0, white_bkgd=False):
    raw, z, d_rgb = _f32c(raw), _f32c(z), _f32c(d_rgb)
    rays_d, d_stride = _rows3(rays_d)
    require_cuda(raw, z, d_rgb)
    B, n = z.shape
    opt = [None if t is None else _f32c(t) for t in (d_disp, d_acc, d_depth, d_weights)]
    nz = _f32c(noise) if (noise is not None and raw_noise_std > 0) else None
    d_raw = torch.empty((B, n, 4), dtype=torch.float32, device=z.device)
    call("nmx_composite_bwd", ptr(raw), ptr(z), ptr(rays_d), i32(d_stride), ptr(nz), f32(raw_noise_std),
         i32(1 if white_bkgd else 0), ptr(d_rgb), ptr(opt[0]), ptr(opt[1]), ptr(opt[2]), ptr(opt[3]), ptr(d_raw),
         i64(B), i32(n), stream())
    return d_raw


def composite_loss_fwd_bwd(raw, z, rays_d, target, white_bkgd=False, want_weights=False, want_rgb=False):
    """Fused training pass: raw2outputs -> mean((rgb - target)^2) -> gradient w.r.t. raw, one kernel.
    Returns (loss [1], d_raw [B,n,4], weights [B,n,1] or None, rgb [B,3] or None)."""
    raw, z, target = _f32c(raw), _f32c(z), _f32c(target)
    rays_d, d_stride = _rows3(rays_d)
    require_cuda(raw, z, target)
    B, n = z.shape
    dev = z.device
    loss = torch.zeros((1,), dtype=torch.float32, device=dev)
    d_raw = torch.empty((B, n, 4), dtype=torch.float32, device=dev)
    weights = torch.empty((B, n, 1), dtype=torch.float32, device=dev) if want_weights else None
    rgb = torch.empty((B, 3), dtype=torch.float32, device=dev) if want_rgb else None
    call("nmx_composite_loss_fwd_bwd", ptr(raw), ptr(z), ptr(rays_d), i32(d_stride), i32(1 if white_bkgd else 0), ptr(target),
         ptr(loss), ptr(d_raw), ptr(rgb), ptr(weights), i64(B), i32(n), stream())
    return loss, d_raw, weights, rgb


# ----------------------------------------------------------------------------------------- resampling
def sample_pdf(z, weights, u, eps=1e-5, cdf=None, want_inds=False, want_cdf=False, want_merged=True, want_imp=True):
    z, u = _f32c(z), _f32c(u)
    require_cuda(z, u)
    B, n = z.shape
    N = u.shape[-1]
    w = None if weights is None else _f32c(weights).reshape(B, n)
    c_in = None if cdf is None else _f32c(cdf)
    dev = z.device
    z_imp = torch.empty((B, N), dtype=torch.float32, device=dev) if want_imp else None
    inds = torch.empty((B, N), dtype=torch.int32, device=dev) if want_inds else None
    cdf_out = torch.empty((B, n + 1), dtype=torch.float32, device=dev) if want_cdf else None
    merged = torch.empty((B, n + N), dtype=torch.float32, device=dev) if want_merged else None
    call("nmx_sample_pdf_fwd", ptr(z), ptr(w), ptr(u), ptr(c_in), f32(eps), ptr(z_imp), ptr(inds), ptr(cdf_out),
         ptr(merged), i64(B), i32(n), i32(N), stream())
    return {"z_imp": z_imp, "inds": inds, "cdf": cdf_out, "z_merged": merged}


def sort_merge_z(a, b):
    a, b = _f32c(a), _f32c(b)
    require_cuda(a, b)
    B = a.shape[0]
    out = torch.empty((B, a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    call("nmx_sort_merge_z", ptr(a), ptr(b), ptr(out), i64(B), i32(a.shape[1]), i32(b.shape[1]), stream())
    return out


# ----------------------------------------------------------------------------------------- loss / optimiser
def mse_fwd_bwd(pred, target, want_grad=True, grad_scale=1.0):
    pred, target = _f32c(pred), _f32c(target)
    require_cuda(pred, target)
    loss = torch.zeros((1,), dtype=torch.float32, device=pred.device)
    d = torch.empty_like(pred) if want_grad else None
    call("nmx_mse_fwd_bwd", ptr(pred), ptr(target), ptr(loss), ptr(d), i64(pred.numel()), f32(grad_scale), stream())
    return loss, d


def adam_step(p, g, m, v, lr, b1=0.9, b2=0.999, eps=1e-8, bias_correction=False, t=1, lr_dev=None):
    require_cuda(p, g, m, v)
    if lr_dev is not None:  # learning rate from a device scalar (CUDA-graph replays); MLX-style update only
        if bias_correction:
            raise NmxError("adam_step: lr_dev is only supported without bias correction")
        require_cuda(lr_dev)
        call("nmx_adam_step_lrdev", ptr(p), ptr(g), ptr(m), ptr(v), i64(p.numel()), ptr(lr_dev), f32(b1), f32(b2), f32(eps),
             stream())
        return
    call("nmx_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), i64(p.numel()), f32(lr), f32(b1), f32(b2), f32(eps),
         i32(1 if bias_correction else 0), i64(t), stream())


# ----------------------------------------------------------------------------------------- GEMM building blocks
def gemm_bf16(A, B, bias=None, relu=False, out_fp32=False):
    """D = act(A @ B.T + bias); A [M,K] bf16, B [N,K] bf16 (tcgen05 kernel; unit-test hook)."""
    require_cuda(A, B, bias)
    M, K = A.shape
    N = B.shape[0]
    D = torch.empty((M, N), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=A.device)
    call("nmx_gemm_bf16", ptr(A), ptr(B), ptr(bias), ptr(D), i64(M), i32(N), i32(K), i32(1 if relu else 0),
         i32(1 if out_fp32 else 0), stream())
    return D


def gemm_pair_bf16(A, B, max_pairs=0):
    """D[M,256] fp32 = A[M,K] @ B[256,K].T on CTA pairs (tcgen05 cta_group::2; unit-test hook)."""
    require_cuda(A, B)
    M, K = A.shape
    if B.shape != (256, K):
        raise ValueError("B must be [256, K]")
    D = torch.empty((M, 256), dtype=torch.float32, device=A.device)
    call("nmx_gemm_pair_bf16", ptr(A), ptr(B), ptr(D), i64(M), i32(K), i32(max_pairs), stream())
    return D


def wgrad_bf16(dY, X, want_db=False):
    """dW[M,N] = dY[P,M].T @ X[P,N] in fp32 (tcgen05 kernel, MN-major operands; unit-test hook).
    want_db: also return db[M] = column sums of dY (bias gradient) from the fused ones-column MMA."""
    require_cuda(dY, X)
    P, M = dY.shape
    N = X.shape[1]
    dW = torch.zeros((M, N), dtype=torch.float32, device=dY.device)
    db = torch.zeros((M,), dtype=torch.float32, device=dY.device) if want_db else None
    call("nmx_wgrad_bf16", ptr(dY), ptr(X), ptr(dW), ptr(db), i64(P), i32(M), i32(N), stream())
    return (dW, db) if want_db else dW


def colsum_bf16(Y):
    require_cuda(Y)
    P, N = Y.shape
    out = torch.zeros((N,), dtype=torch.float32, device=Y.device)
    call("nmx_colsum_bf16", ptr(Y), ptr(out), i64(P), i32(N), stream())
    return out
