"""chain2 (CTA pairs, two tiles in ping-pong; NMX_DISABLE_CHAIN2=1 turns it off) vs the one-tile-per-CTA chain: the inference
forward (chain2 when enabled) against the training forward (NMX_DISABLE_CHAIN2T=1: the one-tile chain) on the same inputs, and timing.

    python scripts/chain2_check.py      # NMX_CHAIN2_DBG selects the timing-only experiments (experiments build only:
                                        # python -m nerf_meets_mlx_b200.build --experiments; the default build ignores it)
"""
import sys, time
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models import NeRF
KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)
torch.manual_seed(3)
net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, **KW)
for B, n in ((2, 128), (9, 64), (37, 64), (75, 64), (2368, 64), (8192, 192)):  # 1, 3, 10, 19, 592, 6144 pair tiles
    o = torch.randn(B, 3, device="cuda")
    d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda"), dim=-1)
    rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
    z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
    net.reserve(B * n, training=True)
    raw_t = net._fwd_raw(1, rays, z, None, B, n, save=True).clone()
    torch.cuda.synchronize()
    print(f"B={B} n={n}: training-chain forward done", flush=True)
    raw_i = net._fwd_raw(1, rays, z, None, B, n, save=False)
    torch.cuda.synchronize()
    err = float((raw_i - raw_t).abs().max() / raw_t.abs().max())
    print(f"   inference vs training chain: rel max err {err:.3e}  finite={bool(torch.isfinite(raw_i).all())}", flush=True)
    if err > 1e-2:
        bad = torch.nonzero((raw_i - raw_t).abs().max(dim=1).values > 1e-2 * raw_t.abs().max()).flatten()
        print("   bad rows:", bad.numel(), bad[:8].tolist(), bad[-4:].tolist())
B, n = 8192, 192
for _ in range(3): net._fwd_raw(1, rays, z, None, B, n, save=False)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): net._fwd_raw(1, rays, z, None, B, n, save=False)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"inference {B}x{n}: {ms:.3f} ms  {B * n / ms / 1e3:.1f} Msamples/s  {B * n * 1186816 / ms / 1e9:.0f} TFLOP/s algorithmic")
