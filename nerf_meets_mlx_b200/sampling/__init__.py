"""Mirror of mlx_nerf/sampling/__init__.py (reference file:line in each docstring) on CUDA tensors."""
import torch

from .. import ops

__all__ = ["add_noise_z", "sample_from_inverse_cdf", "sample_from_inverse_cdf_torch", "sample_and_merge"]


def add_noise_z(z_vals, strength=1.0, t_rand=None):
    """add_noise_z (sampling/__init__.py:10-31).  Returns z_vals unchanged when strength <= 0 (:14-15).
    Declared deviations: the intended `[..., -1:]`/`[..., :1]` slices (the committed code cannot
    concatenate), and the uniform draw may be passed explicitly as `t_rand` (default torch.rand on device)."""
    if strength <= 0.0:
        return z_vals
    if t_rand is None:
        t_rand = torch.rand(z_vals.shape, device=z_vals.device, dtype=torch.float32)
    return ops.add_noise_z(z_vals, t_rand, float(strength))


def _draw_u(B, N, device, is_stratified_sampling):
    if is_stratified_sampling:
        # the reference's stratified branch raises TypeError (:135-136); intended semantics: linspace(0, 1, N) per ray
        return torch.linspace(0.0, 1.0, N, device=device).expand(B, N).contiguous()
    return torch.rand((B, N), device=device, dtype=torch.float32)


def sample_from_inverse_cdf_torch(z_vals, weights, n_importance_samples, eps=1e-5, is_stratified_sampling=False,
                                  u_vals=None):
    """sample_from_inverse_cdf_torch (sampling/__init__.py:101-178) -> [B, n_importance_samples], UNSORTED.
    z_vals [B, n], weights [B, n, 1].  `u_vals` [B, N] may be passed explicitly (the reference draws torch.rand)."""
    B = z_vals.shape[0]
    if u_vals is None:
        u_vals = _draw_u(B, n_importance_samples, z_vals.device, is_stratified_sampling)
    r = ops.sample_pdf(z_vals, weights, u_vals, eps=eps, want_merged=False)
    return r["z_imp"]


def sample_from_inverse_cdf(z_vals, weights, n_importance_samples, eps=1e-5, is_stratified_sampling=False, u_vals=None):
    """sample_from_inverse_cdf (sampling/__init__.py:34-99).  The reference's MLX version is dead code that cannot
    run (normal-distributed u, 1-D searchsorted on a 2-D CDF, unpadded mid-points); name and signature are kept and
    routed to the same kernel as the torch version."""
    return sample_from_inverse_cdf_torch(z_vals, weights, n_importance_samples, eps, is_stratified_sampling, u_vals)


def sample_and_merge(z_vals, weights, n_importance_samples, eps=1e-5, u_vals=None):
    """Fused resample + sort(concat([z_vals, z_imp])) (render.py:215-225 / __test_nerf.py:275-288) in one kernel,
    no host round trip.  Returns (z_merged [B, n+N], z_imp [B, N])."""
    B = z_vals.shape[0]
    if u_vals is None:
        u_vals = _draw_u(B, n_importance_samples, z_vals.device, False)
    r = ops.sample_pdf(z_vals, weights, u_vals, eps=eps, want_merged=True)
    return r["z_merged"], r["z_imp"]
