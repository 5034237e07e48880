"""Isolated timing of the fused MLP forward chain (inference and training-forward) at the fine-pass size."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models.NeRF import NeRF

B, n = 8192, 192
net = NeRF(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
           list_skip_connection_layers=[4], is_use_view_directions=True, n_freqs_pos=10, n_freqs_dir=4)
torch.manual_seed(0)
o = torch.randn(B, 3, device="cuda")
d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda"), dim=-1)
rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
net.reserve(B * n, training=True)
FLOP = 2 * 593408 * B * n


def bench(fn, name, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.3f} ms  {FLOP / ms / 1e9:.1f} TFLOP/s (algorithmic)", flush=True)


bench(lambda: net._fwd_raw(1, rays, z, None, B, n, save=False), "chain fwd inference (incl. encode kernels)")
bench(lambda: net._fwd_raw(1, rays, z, None, B, n, save=True), "chain fwd training  (incl. encode kernels, saves activations)")
