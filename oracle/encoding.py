"""Oracle (TEST INFRASTRUCTURE): positional encodings and the multiresolution hash grid, NumPy fp32.

Follows mlx_nerf/models/embedding.py:4-90 (PE flavour A, volume path),
mlx_nerf/encoding/sinusoidal.py:13-66 (PE flavour B, image path) and
mlx_nerf/encoding/multi_hash.py:13-137 (hash grid, canonical semantics of SURVEY 8a row 9),
mlx_nerf/encoding/spherical_harmonics.py:13-94 and mlx_nerf/encoding/identity.py:13-32.
"""
import math

import numpy as np

from .sampling import linspace_mlx

F32 = np.float32


# ----------------------------------------------------------------------------- PE flavour A
def embedder_freq_bands(n_freqs, max_freq_log2=None):
    """embedding.py:47-49: linspace(0, max_freq_log2, n_freqs) ** 2.0 -- SQUARES; with get_embedder's
    max_freq_log2 = n_freqs - 1 (embedding.py:81, the default here): [0,1,4,9,...], not powers of two (reference quirk,
    replicated)."""
    hi = n_freqs - 1 if max_freq_log2 is None else max_freq_log2
    return (linspace_mlx(0.0, hi, n_freqs) ** F32(2.0)).astype(F32)


def embedder_out_dim(n_freqs, n_input_dims=3):
    if n_freqs == -1:
        return 3
    include = 0 if n_input_dims == 2 else n_input_dims  # embedding.py:79 (`False if 2 == dims else 3`)
    return include + n_input_dims * 2 * n_freqs


def embedder_embed(x, n_freqs, n_input_dims=3, max_freq_log2=None, include_input=None):
    """Embedder.embed (embedding.py:65-71): [x, sin(f0 x), cos(f0 x), sin(f1 x), cos(f1 x), ...].  include_input defaults
    to get_embedder's choice (embedding.py:79: off for 2-D inputs)."""
    x = np.asarray(x, dtype=F32)
    if n_freqs == -1:
        return x
    outs = []
    if (n_input_dims != 2) if include_input is None else include_input:
        outs.append(x)
    for f in embedder_freq_bands(n_freqs, max_freq_log2):
        xf = (x * f).astype(F32)
        outs.append(np.sin(xf).astype(F32))
        outs.append(np.cos(xf).astype(F32))
    return np.concatenate(outs, axis=-1).astype(F32)


def embed(pos, n_freqs_pos, dirs, n_freqs_dir):
    """embedding.embed (embedding.py:4-21): flatten pos [B,n,3]; repeat dir per sample; concat."""
    pos = np.asarray(pos, dtype=F32)
    pos_flat = pos.reshape(-1, pos.shape[-1])
    e_pos = embedder_embed(pos_flat, n_freqs_pos)
    if dirs is None:
        return e_pos
    d = np.repeat(np.asarray(dirs, dtype=F32)[:, None, :], pos.shape[1], axis=1).reshape(-1, dirs.shape[-1])
    e_dir = embedder_embed(d, n_freqs_dir)
    return np.concatenate([e_pos, e_dir], axis=-1)


# ----------------------------------------------------------------------------- PE flavour B
def sinusoidal_freq_bands(n_freqs, min_freq_exp=None, max_freq_exp=None):
    """sinusoidal.py:25-26,49-51: falsy exps default to 0.0 / n_freqs-1; bands = 2 ** linspace."""
    lo = min_freq_exp if min_freq_exp else 0.0
    hi = max_freq_exp if max_freq_exp else float(n_freqs - 1)
    return np.power(F32(2.0), linspace_mlx(lo, hi, n_freqs)).astype(F32)


def sinusoidal_out_dim(in_dim, n_freqs, is_include_input=False):
    return in_dim * n_freqs * 2 + (in_dim if is_include_input else 0)


def sinusoidal_encode(x, n_freqs, min_freq_exp=None, max_freq_exp=None, is_include_input=False):
    """SinusoidalEncoding.__call__ (sinusoidal.py:39-66): dim-major / freq-minor scaled inputs,
    out = sin([s, s + fp32(pi/2)]), optional input appended AT THE END."""
    x = np.asarray(x)
    xin = x
    bands = sinusoidal_freq_bands(n_freqs, min_freq_exp, max_freq_exp)
    s = (x[..., None].astype(F32) * bands).astype(F32)
    s = s.reshape(s.shape[0], -1)
    half_pi = F32(np.pi / 2.0)
    out = np.sin(np.concatenate([s, (s + half_pi).astype(F32)], axis=-1)).astype(F32)
    if is_include_input:
        out = np.concatenate([out, xin.astype(F32)], axis=-1)
    return out


# ----------------------------------------------------------------------------- hash grid
PRIMES = (1, 2654435761, 805459861)  # multi_hash.py:66-70


def hashgrid_scaled_res(n_levels, min_res, max_res):
    """multi_hash.py:32-40: b = exp((ln Nmax - ln Nmin)/(L-1)); N_l = floor(Nmin * b**l), fp32."""
    if n_levels > 1:
        b = np.exp((np.log(F32(max_res)) - np.log(F32(min_res))) / F32(n_levels - 1)).astype(F32)
    else:
        b = F32(1.0)
    levels = np.arange(n_levels).astype(F32)
    return np.floor(F32(min_res) * np.power(b, levels).astype(F32)).astype(F32)


def hashgrid_hash(coords, log2_T):
    """MultiHashEncoding.hash (multi_hash.py:61-77), canonical semantics: int32 coordinates are
    reinterpreted as uint32, products and XOR wrap mod 2^32, `% T` == `& (T-1)` (T power of two)."""
    c = np.asarray(coords).astype(np.int32).view(np.uint32).astype(np.uint64)
    h = np.zeros(c.shape[:-1], dtype=np.uint64)
    for i in range(c.shape[-1]):
        h ^= (c[..., i] * np.uint64(PRIMES[i])) & np.uint64(0xFFFFFFFF)
    return (h & np.uint64((1 << log2_T) - 1)).astype(np.int64)


# corner order of multi_hash.py:102-109: per corner, which of (x,y,z) takes the CEIL coordinate
CORNERS_CEIL = (
    (1, 1, 1),  # grid_0
    (1, 0, 1),  # grid_1
    (0, 0, 1),  # grid_2
    (0, 1, 1),  # grid_3
    (1, 1, 0),  # grid_4
    (1, 0, 0),  # grid_5
    (0, 0, 0),  # grid_6
    (0, 1, 0),  # grid_7
)


def hashgrid_corner_indices(x, scaled_res, log2_T):
    """Table indices of the 8 corners: [B, L, 8] int64, and interpolation offsets [B, L, 3] fp32."""
    x = np.asarray(x, dtype=F32)
    p = (x[:, None, :] * np.asarray(scaled_res, dtype=F32)[:, None]).astype(F32)  # [B, L, 3]
    pc = np.ceil(p).astype(np.int32)
    pf = np.floor(p).astype(np.int32)
    idx = []
    for cx, cy, cz in CORNERS_CEIL:
        g = np.stack([pc[..., 0] if cx else pf[..., 0],
                      pc[..., 1] if cy else pf[..., 1],
                      pc[..., 2] if cz else pf[..., 2]], axis=-1)
        idx.append(hashgrid_hash(g, log2_T))
    offset = (p - pf.astype(F32)).astype(F32)
    return np.stack(idx, axis=-1), offset


def hashgrid_encode(x, tables, scaled_res, log2_T):
    """MultiHashEncoding.__call__ (multi_hash.py:79-137) with per-level tables [L, T, F].
    Interpolation order exactly as :123-131 (x pairs 03,12,56,47; then y; then z)."""
    tables = np.asarray(tables, dtype=F32)
    L = tables.shape[0]
    idx, off = hashgrid_corner_indices(x, scaled_res, log2_T)  # [B,L,8], [B,L,3]
    lv = np.arange(L)[None, :]
    h = [tables[lv, idx[..., k]] for k in range(8)]  # each [B, L, F]
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    one = F32(1.0)
    h03 = h[0] * ox + h[3] * (one - ox)
    h12 = h[1] * ox + h[2] * (one - ox)
    h56 = h[5] * ox + h[6] * (one - ox)
    h47 = h[4] * ox + h[7] * (one - ox)
    h0312 = h03 * oy + h12 * (one - oy)
    h4756 = h47 * oy + h56 * (one - oy)
    out = h0312 * oz + h4756 * (one - oz)
    return out.reshape(out.shape[0], -1).astype(F32)


def hashgrid_backward(x, d_out, n_levels, n_feat, scaled_res, log2_T):
    """Gradient of `hashgrid_encode` w.r.t. the tables (fp64 accumulation, for tolerance tests)."""
    idx, off = hashgrid_corner_indices(x, scaled_res, log2_T)
    B = idx.shape[0]
    d = np.asarray(d_out, dtype=np.float64).reshape(B, n_levels, n_feat)
    ox, oy, oz = (off[..., i].astype(np.float64) for i in range(3))
    g = np.zeros((n_levels, 1 << log2_T, n_feat), dtype=np.float64)
    for k, (cx, cy, cz) in enumerate(CORNERS_CEIL):
        w = (ox if cx else 1 - ox) * (oy if cy else 1 - oy) * (oz if cz else 1 - oz)
        for l in range(n_levels):
            np.add.at(g[l], idx[:, l, k], w[:, l, None] * d[:, l, :])
    return g


# ----------------------------------------------------------------------------- SH / identity
def sh_out_dim(n_degrees):
    """spherical_harmonics.py:28-31."""
    return (n_degrees + 1) ** 2


def sh_encode(dirs, n_degrees):
    """spherical_harmonics.py:33-94: real SH basis (r = 1) of the first three components, fp32, Python scalars stay
    weak (fp32 products), evaluation order as written."""
    assert 0 <= n_degrees <= 4
    d = np.asarray(dirs, dtype=F32)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    out = np.zeros((*d.shape[:-1], sh_out_dim(n_degrees)), dtype=F32)
    c = lambda v: F32(v)
    out[..., 0] = c(0.28209479177387814)
    if n_degrees >= 1:
        out[..., 1] = c(0.4886025119029199) * y
        out[..., 2] = c(0.4886025119029199) * z
        out[..., 3] = c(0.4886025119029199) * x
    if n_degrees >= 2:
        out[..., 4] = c(1.0925484305920792) * xy
        out[..., 5] = c(1.0925484305920792) * yz
        out[..., 6] = c(0.9461746957575601) * zz - c(0.31539156525251999)
        out[..., 7] = c(1.0925484305920792) * xz
        out[..., 8] = c(0.5462742152960396) * (xx - yy)
    if n_degrees >= 3:
        out[..., 9] = c(0.5900435899266435) * y * (c(3) * xx - yy)
        out[..., 10] = c(2.890611442640554) * xy * z
        out[..., 11] = c(0.4570457994644658) * y * (c(5) * zz - c(1))
        out[..., 12] = c(0.3731763325901154) * z * (c(5) * zz - c(3))
        out[..., 13] = c(0.4570457994644658) * x * (c(5) * zz - c(1))
        out[..., 14] = c(1.445305721320277) * z * (xx - yy)
        out[..., 15] = c(0.5900435899266435) * x * (xx - c(3) * yy)
    if n_degrees >= 4:
        out[..., 16] = c(2.5033429417967046) * xy * (xx - yy)
        out[..., 17] = c(1.7701307697799304) * yz * (c(3) * xx - yy)
        out[..., 18] = c(0.9461746957575601) * xy * (c(7) * zz - c(1))
        out[..., 19] = c(0.6690465435572892) * yz * (c(7) * zz - c(3))
        out[..., 20] = c(0.10578554691520431) * (c(35) * zz * zz - c(30) * zz + c(3))
        out[..., 21] = c(0.6690465435572892) * xz * (c(7) * zz - c(3))
        out[..., 22] = c(0.47308734787878004) * (xx - yy) * (c(7) * zz - c(1))
        out[..., 23] = c(1.7701307697799304) * xz * (xx - c(3) * yy)
        out[..., 24] = c(0.6258357354491761) * (xx * (xx - c(3) * yy) - yy * (c(3) * xx - yy))
    return out


def identity_encode(x):
    """identity.py:26-31."""
    return x
