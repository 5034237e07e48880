"""Mirror of mlx_nerf/encoding/__init__.py: `Encoding` base (encoding/__init__.py:10-24) + the encoders."""
from abc import abstractmethod

import torch


class Encoding(torch.nn.Module):
    def __init__(self, in_dim: int) -> None:
        super().__init__()
        self.in_dim = in_dim

    @abstractmethod
    def forward(self, in_array):
        raise NotImplementedError

    @abstractmethod
    def get_out_dim(self):
        raise NotImplementedError


from .sinusoidal import SinusoidalEncoding  # noqa: E402,F401
from .multi_hash import MultiHashEncoding  # noqa: E402,F401
from .spherical_harmonics import SphericalHarmonicsEncoding  # noqa: E402,F401
from .identity import IdentityEncoding  # noqa: E402,F401
