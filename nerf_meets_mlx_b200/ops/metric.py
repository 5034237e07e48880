"""Mirror of mlx_nerf/ops/metric.py: MSE and PSNR on device tensors.

SSIM is unfinished in the reference (its __call__ returns None after building the windows, metric.py:20-47) and LPIPS
needs the `lpips` package and VGG weights (metric.py:66-75); both raise NotImplementedError here, as does
`loss_to_PSNR`, which the reference itself stubs (metric.py:8-10)."""
import torch

from . import mse_fwd_bwd


def loss_to_PSNR(loss):
    raise NotImplementedError  # metric.py:8-10 returns the NotImplementedError class; raising is the intent


class MSE:
    """metric.py:12-14: mean((pred - gt)^2) over all elements -> device scalar (nmx_mse_fwd_bwd, no gradient)."""

    def __call__(self, pred, gt):
        pred, gt = torch.broadcast_tensors(pred, gt)
        loss, _ = mse_fwd_bwd(pred, gt, want_grad=False)
        return loss[0]


class PSNR:
    """metric.py:16-18: 10 * log10(1 / MSE)."""

    def __call__(self, pred, gt):
        return 10 * torch.log10(1 / MSE()(pred, gt))


class SSIM:
    def __call__(self, pred, gt, w_size=11, size_average=True, full=False):
        raise NotImplementedError("SSIM is unfinished in the reference (ops/metric.py:20-47)")


class LPIPS:
    def __init__(self) -> None:
        raise NotImplementedError("LPIPS needs the external `lpips` package (ops/metric.py:66-75)")
