"""Mirror of mlx_nerf/encoding/sinusoidal.py."""
import torch

from .. import ops
from . import Encoding


def mlx_linspace(start, stop, num):
    """mx.linspace as assumed for MLX 0.7.0: arange(num) * fp32((stop-start)/(num-1)) + start, all fp32."""
    if num == 1:  # mx.linspace(a, b, num=1) is [a] (no division by num - 1)
        return torch.tensor([float(start)], dtype=torch.float32)
    seq = torch.arange(num, dtype=torch.float32)
    step = torch.tensor((float(stop) - float(start)) / (num - 1), dtype=torch.float32)
    return seq * step + torch.tensor(float(start), dtype=torch.float32)


class SinusoidalEncoding(Encoding):
    """SinusoidalEncoding (encoding/sinusoidal.py:13-66): s = x[..., None] * 2**linspace(min, max, N) (dim-major,
    freq-minor), out = sin([s, s + fp32(pi/2)]) (the reference's cos), optional input appended at the END."""

    def __init__(self, in_dim: int, n_freqs: int, min_freq_exp: float = None, max_freq_exp: float = None,
                 is_include_input: bool = False) -> None:
        super().__init__(in_dim)
        self.n_freqs = n_freqs
        self.min_freq_exp = min_freq_exp if min_freq_exp else 0.0          # sinusoidal.py:25 (falsy -> default)
        self.max_freq_exp = max_freq_exp if max_freq_exp else float(n_freqs - 1)
        self.is_include_input = is_include_input
        self._bands = {}

    def get_out_dim(self):
        out_dim = self.in_dim * self.n_freqs * 2
        if self.is_include_input:
            out_dim += self.in_dim
        return out_dim

    def freq_bands(self, device):
        key = str(device)
        if key not in self._bands:
            b = torch.pow(torch.tensor(2.0, dtype=torch.float32), mlx_linspace(self.min_freq_exp, self.max_freq_exp, self.n_freqs))
            self._bands[key] = b.to(device)
        return self._bands[key]

    def forward(self, in_array):
        x = in_array.to(torch.float32)  # integer pixel coordinates promote to fp32 (MLX int32*float32 -> float32)
        return ops.pe_sinusoidal(x.reshape(-1, self.in_dim), self.freq_bands(x.device), self.is_include_input)
