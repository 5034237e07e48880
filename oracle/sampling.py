"""Oracle (TEST INFRASTRUCTURE): depth sampling + inverse-CDF resampling, NumPy fp32.

Follows mlx_nerf/sampling/uniform.py:7-18, linear_disparity.py:8-19, sampling/__init__.py:10-31
and :101-178 of the reference.
"""
import numpy as np

F32 = np.float32


def linspace_mlx(start, stop, num):
    """mx.linspace as assumed for MLX 0.7.0: arange(num)*step + start, all fp32 (third-party)."""
    if num == 1:  # mx.linspace(a, b, num=1) is [a]
        return np.array([start], dtype=F32)
    seq = np.arange(num, dtype=F32)
    step = F32((float(stop) - float(start)) / (num - 1))
    return seq * step + F32(start)


def sample_z_uniform(near, far, n_samples):
    """uniform.sample_z (sampling/uniform.py:13-16): z = near*(1-t) + far*t, two-product form."""
    near = np.asarray(near, dtype=F32)
    far = np.asarray(far, dtype=F32)
    t = linspace_mlx(0.0, 1.0, n_samples)
    z_from = near * (F32(1.0) - t)
    z_to = far * t
    return (z_from + z_to).astype(F32)


def sample_z_lindisp(near, far, n_samples):
    """linear_disparity.sample_z (sampling/linear_disparity.py:14-17), restated AS WRITTEN
    (1/(1/(near(1-t)) + 1/(far t)); endpoints hit +-inf -> 0)."""
    near = np.asarray(near, dtype=F32)
    far = np.asarray(far, dtype=F32)
    t = linspace_mlx(0.0, 1.0, n_samples)
    with np.errstate(divide="ignore", invalid="ignore"):
        z_from = F32(1.0) / (near * (F32(1.0) - t))
        z_to = F32(1.0) / (far * t)
        return (F32(1.0) / (z_from + z_to)).astype(F32)


def add_noise_z(z_vals, t_rand=None, strength=1.0):
    """add_noise_z (sampling/__init__.py:10-31).  Declared deviations: the
    intended `[..., -1:]` / `[..., :1]` slices are used (the committed code drops a dim and cannot
    concatenate), and the uniform draw `t_rand` in [0,1) is an explicit input."""
    if strength <= 0.0:
        return z_vals
    z = np.asarray(z_vals, dtype=F32)
    t = (np.asarray(t_rand, dtype=F32) * F32(strength)).astype(F32)
    mids = F32(0.5) * (z[..., :-1] + z[..., 1:])
    upper = np.concatenate([mids, z[..., -1:]], axis=-1)
    lower = np.concatenate([z[..., :1], mids], axis=-1)
    return (lower + (upper - lower) * t).astype(F32)


def build_cdf(weights, eps=1e-5):
    """CDF construction of sample_from_inverse_cdf_torch (sampling/__init__.py:113-131).

    Canonical arithmetic: the row sum and the running prefix are taken in fp64 and
    rounded to fp32 (torch-CPU `cumsum` on fp32 does exactly this; torch-CPU `sum` uses an
    ISA-dependent vectorised fp32 order that differs from it by <= 3 ulp on some rows).
    Returns cdf [B, n+1] fp32.
    """
    w = (np.asarray(weights, dtype=F32)[..., 0] + F32(0.01)).astype(F32)
    n = w.shape[-1]
    w_sum = np.sum(w.astype(np.float64), axis=-1, keepdims=True).astype(F32)
    padding = np.maximum(F32(eps) - w_sum, F32(0.0)).astype(F32)
    w = (w + padding / F32(n)).astype(F32)
    w_sum = (w_sum + padding).astype(F32)
    pdf = (w / w_sum).astype(F32)
    cdf = np.minimum(F32(1.0), np.cumsum(pdf.astype(np.float64), axis=-1).astype(F32))
    cdf = np.concatenate([np.zeros_like(cdf[..., :1]), cdf], axis=-1)
    return cdf.astype(F32)


def sample_pdf(z_vals, weights, u_vals, eps=1e-5, cdf=None, return_inds=False):
    """sample_from_inverse_cdf_torch (sampling/__init__.py:101-178) with `u_vals` [B, N] explicit
    (the reference draws torch.rand from the global RNG, :139-141).  Output is UNSORTED [B, N]."""
    z = np.asarray(z_vals, dtype=F32)
    u = np.asarray(u_vals, dtype=F32)
    if cdf is None:
        cdf = build_cdf(weights, eps)
    B, n1 = cdf.shape
    # torch.searchsorted(cdf, u, side="right"): number of edges <= u
    inds = np.stack([np.searchsorted(cdf[b], u[b], side="right") for b in range(B)]).astype(np.int64)
    below = np.clip(inds - 1, 0, n1 - 1)
    above = np.clip(inds, 0, n1 - 1)
    cdf_from = np.take_along_axis(cdf, below, axis=-1)
    cdf_to = np.take_along_axis(cdf, above, axis=-1)
    z_mid = ((z[..., 1:] + z[..., :-1]) / F32(2)).astype(F32)
    z_mid = np.concatenate([z_mid[..., :1], z_mid, z_mid[..., -1:]], axis=-1)  # :153-159
    z_from = np.take_along_axis(z_mid, below, axis=-1)
    z_to = np.take_along_axis(z_mid, above, axis=-1)
    t_num = (u - cdf_from).astype(F32)
    t_den = (cdf_to - cdf_from).astype(F32)
    t_den = np.where(t_den < F32(eps), F32(1.0), t_den).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        t = (t_num / t_den).astype(F32)
    t = np.nan_to_num(t, nan=0.0).astype(F32)
    t = np.clip(t, F32(0.0), F32(1.0))
    out = (z_from + t * (z_to - z_from)).astype(F32)
    if return_inds:
        return out, inds
    return out


def merge_sorted(z_vals, z_importance):
    """render.py:225 / __test_nerf.py:288: sort(concat([z_vals, z_imp], -1), -1)."""
    return np.sort(np.concatenate([z_vals, z_importance], axis=-1).astype(F32), axis=-1, kind="stable")
