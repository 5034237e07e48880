"""nerf_meets_mlx_b200 -- B200 (sm_100a) implementation of the volume-learning hot path of
piljoong-jeong/nerf_meets_mlx behind the reference's own call signatures.

Sub-packages mirror the reference: `sampling`, `encoding`, `models`, `rendering`.  All arithmetic runs in
hand-written CUDA kernels (libnmx.so, C ABI in include/nmx.h); there is no CPU path."""
from . import _lib_loader  # noqa: F401

__all__ = ["sampling", "encoding", "models", "rendering", "ops"]
