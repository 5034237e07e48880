"""Pair training forward (nmx_chain2t.cu) vs the one-tile training chain (nmx_chain.cu): outputs, every saved tensor
(h_0..h_7, hd, X0, ReLU sign bits) and the parameter gradient of the following backward pass.  Run twice:
    NMX_DISABLE_CHAIN2T=1 python scripts/chain2t_check.py ref [B n]   # writes gpurun_out/chain2t_ref.pt
    python scripts/chain2t_check.py cmp [B n]                        # compares the pair kernels against it
"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L
from nerf_meets_mlx_b200.models import NeRF

KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)
mode = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 192
P = B * n
torch.manual_seed(3)
net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, seed=7, **KW)
with torch.no_grad():
    net.alpha_linear.bias.fill_(0.3)
    net.mark_params_updated()
o = torch.randn(B, 3, device="cuda") * 0.3 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda") - torch.tensor([0.0, 0.0, 2.0], device="cuda"), dim=-1)
rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
d_out = torch.randn(P, 4, device="cuda") * 1e-3
net.reserve(P, training=True)
raw = net._fwd_raw(1, rays, z, None, B, n, save=True).clone()
torch.cuda.synchronize()
out = (ctypes.c_int64 * 12)()
L.call("nmx_mlp_debug_layout", net._plan, out, L.i32(12))
base, x0, h0, hs, _, hd, _, _, _, bits, cap, x0_cols = list(out)
ws = net._ws


def view(off, rows, cols):
    return ws[base + off: base + off + rows * cols * 2].view(torch.bfloat16).view(rows, cols)


saved = {"raw": raw.cpu(), "x0": view(x0, P, x0_cols).float().cpu(), "hd": view(hd, P, 128).float().cpu()}
for l in range(8):
    saved[f"h{l}"] = view(h0 + l * hs, P, 256).float().cpu()
b = ws[base + bits: base + bits + 9 * cap * 32].view(torch.int32).view(9, cap, 8)
import os
if P >= 4096 and not os.environ.get("NMX_DISABLE_CHAIN2T"):  # pair forward: each 4 KB tile is word-major [8][128]
    b = b.reshape(9, cap // 128, 8, 128).permute(0, 1, 3, 2).reshape(9, cap, 8)
b = b[:, :P].cpu().numpy().astype(np.uint32)
# sign bits must describe the activations saved by the SAME run
for slot in range(9):
    cols = 256 if slot < 8 else 128
    h = saved[f"h{slot}" if slot < 8 else "hd"].numpy()
    got = np.zeros((P, cols), dtype=bool)
    for w in range(cols // 32):
        for e in range(16):
            got[:, 32 * w + 2 * e] = (b[slot, :, w] >> e) & 1
            got[:, 32 * w + 2 * e + 1] = (b[slot, :, w] >> (16 + e)) & 1
    assert np.array_equal(got, h > 0), f"sign bits of slot {slot} do not match the saved activation"
g = net._bwd_raw(d_out, P).clone()
saved["grad"] = g.cpu()
assert all(torch.isfinite(v).all() for v in saved.values())
path = f"gpurun_out/chain2t_ref_{B}_{n}.pt"
if mode == "ref":
    torch.save(saved, path)
    print("reference written", path)
else:
    ref = torch.load(path)
    worst = 0.0
    for k, v in saved.items():
        r = ref[k]
        e = float((v - r).abs().max() / r.abs().max().clamp_min(1e-30))
        en = float((v - r).norm() / r.norm().clamp_min(1e-30))
        print(f"{k:5s} max-rel {e:.3e}  norm-rel {en:.3e}  exact {bool(torch.equal(v, r))}")
        if k == "x0":
            assert torch.equal(v, r), "X0 must be bit-identical"
        else:
            worst = max(worst, en)
    assert worst < 5e-3, worst
    print("CHAIN2T_CHECK_OK", B, n)
