"""Mirror of mlx_nerf/encoding/spherical_harmonics.py."""
from .. import ops
from . import Encoding


class SphericalHarmonicsEncoding(Encoding):
    """SphericalHarmonicsEncoding (encoding/spherical_harmonics.py:13-94): real SH basis of unit directions up to
    degree `n_degrees` in [0, 4]; out_dim = (n_degrees + 1)^2.  Same assertion as the reference (:22)."""

    def __init__(self, in_dim: int, n_degrees: int) -> None:
        super().__init__(in_dim)
        assert 0 <= n_degrees <= 4, f"[ERROR] {n_degrees=} must be in range [0, 4]!"
        self.n_degrees = n_degrees

    def get_out_dim(self):
        return (self.n_degrees + 1) ** 2

    def forward(self, in_dirs):
        return ops.sh_encode(in_dirs, self.n_degrees)
