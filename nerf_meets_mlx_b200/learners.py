"""The two other training steps BASELINE.json's configs name, driven through the raw C-ABI ops (no autograd graph) and
replayable from a CUDA graph:

  * `ImageLearner`    -- C1, the image-learning step of mlx_nerf/entrypoints/__viser_image_learning.py:186-236:
                         SinusoidalEncoding(2, 10, 0, 8) of INTEGER pixel coordinates -> NeRF(40 -> 8x256 -> 3) -> MSE
                         -> Adam(1e-3, betas (0.9, 0.99)), MLX-style (no bias correction);
  * `HashGridLearner` -- C4, MultiHashEncoding (encoding/multi_hash.py:13-137) -> tiny MLP -> MSE, gradients to the MLP
                         and (through nmx_mlp_bwd_input + the atomic scatter) to the hash tables, Adam on both.
                         The reference never wires this model into an entrypoint; the step is the
                         image step's structure with the hash grid as the encoder.
"""
import torch

from . import ops
from .encoding import MultiHashEncoding, SinusoidalEncoding
from .models.NeRF import AdamMLX, NeRF


class _GraphedStep:
    """Capture `body(*static_inputs)` once per input shape and replay it; the first call runs eagerly (warm-up)."""

    def __init__(self):
        self._graphs = {}
        self._calls = 0

    def run(self, body, inputs, use_graph):
        self._calls += 1
        if not use_graph or self._calls == 1:
            return body(*inputs)
        key = tuple(tuple(t.shape) for t in inputs)
        if key not in self._graphs:
            static = [t.clone() for t in inputs]
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = body(*static)
            self._graphs[key] = (g, static, out)
        g, static, out = self._graphs[key]
        for s, t in zip(static, inputs):
            s.copy_(t)
        g.replay()
        return out  # static buffer: overwritten by the next replay


class ImageLearner:
    """C1.  `step(X, y)`: X [B, 2] integer (row, col) pixel coordinates (un-normalised, as batch_iterate yields them,
    __viser_image_learning.py:92-116), y [B, 3] target colours -> device scalar loss."""

    def __init__(self, device="cuda", lr=1e-3, betas=(0.9, 0.99), seed=0, use_cuda_graph=True, max_points=1 << 16):
        self.embed = SinusoidalEncoding(2, 10, min_freq_exp=0.0, max_freq_exp=8.0, is_include_input=False)
        self.model = NeRF(channel_input=self.embed.get_out_dim(), channel_input_views=0, channel_output=3,
                          is_use_view_directions=False, device=device, seed=seed, n_freqs_pos=10, max_points=max_points)
        self.optimizer = AdamMLX(lr, betas=betas)
        self.bands = self.embed.freq_bands(torch.device(device))
        self._g = torch.empty_like(self.model.flat.data)
        self.model.reserve(max_points, training=True)
        self.use_cuda_graph = use_cuda_graph
        self._graph = _GraphedStep()

    def _body(self, X, y):
        B = X.shape[0]
        pred = self.model._fwd_raw(2, X, None, self.bands, B, 1, save=True)  # PE evaluated in the operand producer
        loss, d_pred = ops.mse_fwd_bwd(pred, y)
        self.model._bwd_raw(d_pred, B, out=self._g)
        self.optimizer.update(self.model, self._g)
        return loss

    def step(self, X, y):
        X = X.to(torch.float32).contiguous()  # integer pixel coordinates promote to fp32 (MLX int32 * float32)
        return self._graph.run(self._body, (X, y.float().contiguous()), self.use_cuda_graph)

    @torch.no_grad()
    def predict(self, X):
        X = X.to(torch.float32).contiguous()
        return self.model._fwd_raw(2, X, None, self.bands, X.shape[0], 1, save=False)


class HashGridLearner:
    """C4.  `step(x, y)`: x [P, 3] positions, y [P, C] targets -> device scalar loss; updates the MLP and the tables."""

    def __init__(self, n_levels=16, min_res=16, max_res=2048, n_features=2, log2_T=19, mlp_layers=2, mlp_width=64,
                 out_ch=4, device="cuda", lr=1e-2, seed=0, hash_init_scale=1e-4, use_cuda_graph=True, max_points=1 << 18):
        self.enc = MultiHashEncoding(3, n_levels, min_res, max_res, n_features, log2_T, hash_init_scale=hash_init_scale,
                                     device=device, seed=seed)
        self.model = NeRF(n_layers=mlp_layers, width_layers=mlp_width, channel_input=self.enc.get_out_dim(),
                          channel_input_views=0, channel_output=out_ch, list_skip_connection_layers=[],
                          is_use_view_directions=False, device=device, seed=seed + 1, max_points=max_points)
        self.opt_mlp = AdamMLX(lr, betas=(0.9, 0.99), shared_state=False)
        self.tab_m = torch.zeros_like(self.enc.hash_table.data)
        self.tab_v = torch.zeros_like(self.enc.hash_table.data)
        self.lr = lr
        self._g = torch.empty_like(self.model.flat.data)
        self.model.reserve(max_points, training=True)
        self.use_cuda_graph = use_cuda_graph
        self._graph = _GraphedStep()
        self.L, self.F, self.log2_T = n_levels, n_features, log2_T

    def _body(self, x, y):
        P = x.shape[0]
        tables = self.enc.hash_table.data
        feat = ops.hashgrid_fwd(x, tables, self.enc.scaled_res, self.log2_T)
        pred = self.model._fwd_raw(0, feat, None, None, P, 1, save=True)
        loss, d_pred = ops.mse_fwd_bwd(pred, y)
        _, d_feat = self.model._bwd_raw(d_pred, P, out=self._g, want_input_grad=True)
        d_tab = ops.hashgrid_bwd(x, self.enc.scaled_res, d_feat, self.L, self.F, self.log2_T)
        self.opt_mlp.update(self.model, self._g)
        ops.adam_step(tables.view(-1), d_tab.view(-1), self.tab_m.view(-1), self.tab_v.view(-1), self.lr, 0.9, 0.99, 1e-8)
        return loss

    def step(self, x, y):
        return self._graph.run(self._body, (x.float().contiguous(), y.float().contiguous()), self.use_cuda_graph)
