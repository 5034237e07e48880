"""Oracle (TEST INFRASTRUCTURE): the NeRF MLP in torch-CPU fp32 (autograd gives the gradients
the CUDA backward is checked against).  Follows mlx_nerf/models/NeRF.py:10-48,160-243 and
mlx_nerf/models/embedding.py:4-21.
"""
import math

import numpy as np
import torch

from . import encoding as enc


def linear_init(out_f, in_f, gen):
    """nn.Linear of MLX 0.7.0 (third-party): weight [out, in], weight and bias ~ U(+-1/sqrt(in))."""
    s = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=torch.float32) * 2 - 1) * s
    b = (torch.rand(out_f, generator=gen, dtype=torch.float32) * 2 - 1) * s
    return w, b


class NeRF:
    """Parameter tree and forward of reference `NeRF` (models/NeRF.py:160-243).

    params: dict name -> torch tensor, names as the reference's tree:
      list_linears_pos.{i}.weight/bias, list_linears_dir.0.*, feature_linear.*, alpha_linear.*,
      rgb_linear.*  (view-dir head)  or  output_linear.* (no-view head).
    """

    def __init__(self, n_layers=8, width_layers=256, channel_input=3, channel_input_views=3,
                 channel_output=4, list_skip_connection_layers=(4,), is_use_view_directions=False,
                 seed=0):
        self.D = n_layers
        self.W = width_layers
        self.channel_input_pos = channel_input
        self.channel_input_dir = channel_input_views
        self.channel_output = channel_output
        self.skips = list(list_skip_connection_layers)
        self.use_dirs = is_use_view_directions
        g = torch.Generator().manual_seed(seed)
        p = {}
        W = width_layers
        # NeRF.py:182-188: layer i+1 takes W+channel_input when i is a skip layer
        dims_in = [channel_input] + [W + channel_input if i in self.skips else W for i in range(n_layers - 1)]
        for i, din in enumerate(dims_in):
            p[f"list_linears_pos.{i}.weight"], p[f"list_linears_pos.{i}.bias"] = linear_init(W, din, g)
        if self.use_dirs:
            p["list_linears_dir.0.weight"], p["list_linears_dir.0.bias"] = linear_init(W // 2, W + channel_input_views, g)
            p["feature_linear.weight"], p["feature_linear.bias"] = linear_init(W, W, g)
            p["alpha_linear.weight"], p["alpha_linear.bias"] = linear_init(1, W, g)
            p["rgb_linear.weight"], p["rgb_linear.bias"] = linear_init(3, W // 2, g)
        else:
            p["output_linear.weight"], p["output_linear.bias"] = linear_init(channel_output, W, g)
        self.params = p

    def requires_grad_(self, flag=True):
        for v in self.params.values():
            v.requires_grad_(flag)
        return self

    def forward(self, x):
        """NeRF.forward (NeRF.py:201-243).  ReLU after EVERY trunk layer, skip concat [input_pos, h]
        after layer idx in skips, no activation on feature/alpha/rgb."""
        p = self.params
        if self.use_dirs:
            input_pos = x[..., : self.channel_input_pos]
            input_dir = x[..., self.channel_input_pos:]
        else:
            input_pos = x
        h = input_pos
        for i in range(self.D):
            h = torch.relu(h @ p[f"list_linears_pos.{i}.weight"].T + p[f"list_linears_pos.{i}.bias"])
            if i in self.skips:
                h = torch.cat([input_pos, h], dim=-1)
        if self.use_dirs:
            alpha = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
            feat = h @ p["feature_linear.weight"].T + p["feature_linear.bias"]
            h = torch.cat([feat, input_dir], dim=-1)
            h = torch.relu(h @ p["list_linears_dir.0.weight"].T + p["list_linears_dir.0.bias"])
            rgb = h @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
            return torch.cat([rgb, alpha], dim=-1)
        return h @ p["output_linear.weight"].T + p["output_linear.bias"]


def run_model(pos, n_freqs_pos, dirs, n_freqs_dir, model, netchunk=64 * 1024):
    """run_model (NeRF.py:25-48): flatten -> embed -> chunked forward -> reshape [B, n, C].
    pos: torch/np [B,n,3]; dirs: [B,3] or None.  Returns torch tensor."""
    pos_np = pos.detach().numpy() if isinstance(pos, torch.Tensor) else np.asarray(pos, dtype=np.float32)
    assert pos_np.ndim == 3
    B, n = pos_np.shape[:2]
    dirs_np = None if dirs is None else (dirs.detach().numpy() if isinstance(dirs, torch.Tensor) else np.asarray(dirs, np.float32))
    emb = torch.from_numpy(enc.embed(pos_np, n_freqs_pos, dirs_np, n_freqs_dir))
    outs = [model.forward(emb[i:i + netchunk]) for i in range(0, emb.shape[0], netchunk)]
    out = torch.cat(outs, dim=0)
    return out.reshape(B, n, out.shape[-1])
