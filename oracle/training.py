"""Oracle (TEST INFRASTRUCTURE): the reference's training iteration, torch-CPU fp32.
Follows mlx_nerf/entrypoints/__test_nerf.py:47-145,200-305 and the image-learning step of
mlx_nerf/entrypoints/__viser_image_learning.py:211-236.
"""
import numpy as np
import torch

from . import encoding as enc
from . import rendering as orend
from . import sampling as osamp


class AdamMLX:
    """optim.Adam of MLX 0.7.0 (third-party, assumed): NO bias correction, eps outside the sqrt:
        m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p = p - lr * m / (sqrt(v) + eps).
    State is keyed by parameter NAME only, so two models with the same tree that are updated by
    the same optimizer instance share moments (reference quirk, __test_nerf.py:134,144;
    SURVEY 8a row 16 (i)).  `shared_state=False` keys the state by (model id, name) instead."""

    def __init__(self, learning_rate, betas=(0.9, 0.999), eps=1e-8, shared_state=True):
        self.learning_rate = learning_rate
        self.betas = betas
        self.eps = eps
        self.shared_state = shared_state
        self.state = {}

    def update(self, model, grads):
        b1, b2 = self.betas
        with torch.no_grad():
            for name, p in model.params.items():
                key = name if self.shared_state else (id(model), name)
                g = grads[name]
                m, v = self.state.get(key, (torch.zeros_like(p), torch.zeros_like(p)))
                m = b1 * m + (1 - b1) * g
                v = b2 * v + (1 - b2) * g * g
                self.state[key] = (m, v)
                p -= self.learning_rate * m / (torch.sqrt(v) + self.eps)


def normalize_dirs(rays_d):
    rays_d = orend._t(rays_d)
    return rays_d / torch.sqrt(torch.sum(rays_d * rays_d, dim=-1, keepdim=True))


def assemble_rays(rays_o, rays_d, near, far):
    """__test_nerf.py:57-82: [o, d, near, far, viewdirs]."""
    rays_o, rays_d = orend._t(rays_o), orend._t(rays_d)
    vd = normalize_dirs(rays_d)
    ones = torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, near * ones, far * ones, vd], dim=-1)


def _grads(model, loss):
    names = list(model.params.keys())
    gs = torch.autograd.grad(loss, [model.params[n] for n in names], allow_unused=True)
    return {n: (g if g is not None else torch.zeros_like(model.params[n])) for n, g in zip(names, gs)}


def loss_coarse(model, rays, target, query_fn, n_samples, white_bkgd):
    """mlx_mse_coarse (__test_nerf.py:47-90)."""
    res = orend.render_rays(rays, model, query_fn, n_samples, white_bkgd=white_bkgd)
    return torch.mean((res["rgb_coarse"] - orend._t(target)) ** 2), res


def loss_fine(model, rays_o, rays_d, z_fine, target, query_fn):
    """mlx_mse_fine (__test_nerf.py:93-126): white_bkgd is hard-wired False here (quirk)."""
    rays_o, rays_d, z_fine = orend._t(rays_o), orend._t(rays_d), orend._t(z_fine)
    vd = normalize_dirs(rays_d)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_fine[..., :, None]
    raw = query_fn(pts, vd, model)
    rgb, _, _, _, _ = orend.raw2outputs(raw, z_fine, rays_d, 0.0, False)
    return torch.mean((rgb - orend._t(target)) ** 2), rgb


def train_iteration(coarse, fine, opt, rays_o, rays_d, target, u_vals, query_fn, n_samples=64,
                    near=2.0, far=6.0, white_bkgd=True):
    """One pass of the loop body (__test_nerf.py:240-293): coarse step, coarse RE-forward with the
    updated net, detached inverse-CDF resample + sort-merge, fine step.  Returns a dict."""
    rays = assemble_rays(rays_o, rays_d, near, far)
    coarse.requires_grad_(True)
    lc, _ = loss_coarse(coarse, rays, target, query_fn, n_samples, white_bkgd)
    gc = _grads(coarse, lc)
    coarse.requires_grad_(False)
    opt.update(coarse, gc)
    out = {"loss_coarse": float(lc.detach()), "grads_coarse": gc}
    if fine is None:
        return out
    with torch.no_grad():
        res = orend.render_rays(rays, coarse, query_fn, n_samples, white_bkgd=white_bkgd)
    z = res["z_vals"].numpy()
    w = res["weights"].numpy()
    z_imp = osamp.sample_pdf(z, w, u_vals)
    z_fine = osamp.merge_sorted(z, z_imp)
    fine.requires_grad_(True)
    lf, _ = loss_fine(fine, rays_o, rays_d, z_fine, target, query_fn)
    gf = _grads(fine, lf)
    fine.requires_grad_(False)
    opt.update(fine, gf)
    out.update(loss_fine=float(lf.detach()), grads_fine=gf, z_fine=z_fine, z_imp=z_imp)
    return out


def lr_schedule(i, lrate=5e-4, lrate_decay=250):
    """__test_nerf.py:302-305."""
    return lrate * (0.1 ** (i / (lrate_decay * 1000)))


def image_step(model, opt, X_int, y, n_freqs=10, min_exp=0.0, max_exp=8.0):
    """mlx_mse + step of the image-learning demo (__viser_image_learning.py:211-236)."""
    emb = torch.from_numpy(enc.sinusoidal_encode(np.asarray(X_int), n_freqs, min_exp, max_exp))
    model.requires_grad_(True)
    loss = torch.mean((model.forward(emb) - orend._t(y)) ** 2)
    g = _grads(model, loss)
    model.requires_grad_(False)
    opt.update(model, g)
    return float(loss.detach()), g


def psnr(mse):
    """ops/metric.py:16-18: 10 log10(1/MSE)."""
    return 10.0 * np.log10(1.0 / mse)


def metric_mse(pred, gt):
    """ops/metric.py:12-14 (fp32 mean of squared differences)."""
    pred, gt = np.asarray(pred, dtype=np.float32), np.asarray(gt, dtype=np.float32)
    return np.mean((pred - gt) ** 2, dtype=np.float32)


def metric_psnr(pred, gt):
    """ops/metric.py:16-18."""
    return np.float32(10) * np.log10(np.float32(1) / metric_mse(pred, gt))
