"""CUDA-event time of the training forward (8192 rays x n samples) and of the backward pass; optional ncu target."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models import NeRF
KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 192
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
P = B * n
net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, seed=7, **KW)
o = torch.randn(B, 3, device="cuda") * 0.3 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda") - torch.tensor([0.0, 0.0, 2.0], device="cuda"), dim=-1)
rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
d_out = torch.randn(P, 4, device="cuda") * 1e-3
net.reserve(P, training=True)
g = torch.empty_like(net.flat.data)


def t(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tf = t(lambda: net._fwd_raw(1, rays, z, None, B, n, save=True))
ti = t(lambda: net._fwd_raw(1, rays, z, None, B, n, save=False))
net._fwd_raw(1, rays, z, None, B, n, save=True)
tb = t(lambda: net._bwd_raw(d_out, P, out=g))
fl = 2 * 593408 * P
print(f"B={B} n={n}: fwd train {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s, {4.9e3 * P / tf / 1e6:.0f} GB/s stored) | fwd infer {ti:.3f} ms "
      f"({fl / ti / 1e9:.0f} TFLOP/s) | bwd (chain + wgrad + heads) {tb:.3f} ms")
