"""Mirror of mlx_nerf/encoding/multi_hash.py with the canonical semantics declared in DESIGN.md (the reference
code is WIP and cannot run as committed)."""
import math

import torch

from .. import ops
from ..ops import library as _library  # noqa: F401  (registers torch.ops.nmx.*)
from . import Encoding


class MultiHashEncoding(Encoding):
    """MultiHashEncoding (encoding/multi_hash.py:13-137).
      * N_l = floor(N_min * b**l), b = exp((ln N_max - ln N_min)/(L-1))   (:32-40), computed in fp32 on the host
        and handed to the kernel as `scaled_res [L]`;
      * per-level tables `hash_table [L, T, F]` (the reference builds L nn.Embedding(T, F), :46-49), init U(+-1e-4);
      * hash = (x*1 ^ y*2654435761 ^ z*805459861) mod T in uint32 wraparound (:61-77), every level hashed;
      * corners (floor, ceil), offset = p - floor(p) on the ceil corner, interpolation order of :122-131."""

    def __init__(self, in_dim: int, n_levels: int, min_res: int, max_res: int, n_features_per_level: int,
                 log2_hashmap_size: int, hash_init_scale: float = 0.0001, device="cuda", seed=0) -> None:
        super().__init__(in_dim)
        if in_dim != 3:
            raise NotImplementedError("the reference's corner construction is 3-D only (multi_hash.py:97-109)")
        self.n_levels = n_levels
        self.min_res = min_res
        self.max_res = max_res
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        f32 = torch.float32
        if n_levels > 1:
            b = torch.exp((torch.log(torch.tensor(float(max_res), dtype=f32)) - torch.log(torch.tensor(float(min_res), dtype=f32)))
                          / torch.tensor(float(n_levels - 1), dtype=f32))
        else:
            b = torch.tensor(1.0, dtype=f32)
        self.growing_factor = b
        levels = torch.arange(n_levels, dtype=f32)
        self.register_buffer("scaled_res", torch.floor(torch.tensor(float(min_res), dtype=f32) * torch.pow(b, levels)).to(device))
        self.hash_table_size = 2 ** log2_hashmap_size
        g = torch.Generator().manual_seed(seed)
        init = (torch.rand(n_levels, self.hash_table_size, n_features_per_level, generator=g) * 2 - 1) * hash_init_scale
        self.hash_table = torch.nn.Parameter(init.to(device))

    def get_out_dim(self):
        return self.n_levels * self.n_features_per_level

    def hash(self, in_array):
        """hash (multi_hash.py:61-77): integer grid coordinates [..., 3] -> table index [...] (int32)."""
        return ops.hashgrid_hash(in_array, self.log2_hashmap_size)

    def forward(self, in_array):
        x = in_array.to(torch.float32).reshape(-1, 3).contiguous()
        # custom operator torch.ops.nmx.hashgrid_fwd; its backward (table scatter) is torch.ops.nmx.hashgrid_bwd
        return torch.ops.nmx.hashgrid_fwd(x, self.hash_table, self.scaled_res, self.log2_hashmap_size)
