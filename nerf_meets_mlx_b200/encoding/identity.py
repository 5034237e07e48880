"""Mirror of mlx_nerf/encoding/identity.py."""
from . import Encoding


class IdentityEncoding(Encoding):
    """IdentityEncoding (encoding/identity.py:13-32): no encoding; returns its input (the same tensor, no copy)."""

    def __init__(self, in_dim: int) -> None:
        super().__init__(in_dim)

    def get_out_dim(self):
        return self.in_dim

    def forward(self, in_array):
        return in_array
