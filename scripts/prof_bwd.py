"""Backward pass split: total, wgrad launches (library event timers), rest (heads + fused dgrad chain)."""
import ctypes
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L
from nerf_meets_mlx_b200.models.NeRF import NeRF

B, n = 8192, int(os.environ.get("NS", "192"))
net = NeRF(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
           list_skip_connection_layers=[4], is_use_view_directions=True, n_freqs_pos=10, n_freqs_dir=4)
torch.manual_seed(0)
o = torch.randn(B, 3, device="cuda")
d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda"), dim=-1)
rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
net.reserve(B * n, training=True)
raw = net._fwd_raw(1, rays, z, None, B, n, save=True)
d_raw = torch.randn(B * n, 4, device="cuda") * 1e-3
g = torch.empty_like(net.flat.data)
lib = L.lib()
for _ in range(2):
    net._bwd_raw(d_raw, B * n, out=g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = 5
e0.record()
for _ in range(R):
    net._bwd_raw(d_raw, B * n, out=g)
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1) / R
lib.nmx_profile_enable(1)
for _ in range(R):
    net._bwd_raw(d_raw, B * n, out=g)
ms_w, fl_w, n_w = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
lib.nmx_profile_report(1, ctypes.byref(ms_w), ctypes.byref(fl_w), ctypes.byref(n_w))
ms_g, fl_g, n_g = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
lib.nmx_profile_report(0, ctypes.byref(ms_g), ctypes.byref(fl_g), ctypes.byref(n_g))
lib.nmx_profile_enable(0)
P = B * n
print(f"bwd n={n}: total {tot:.3f} ms | wgrad {ms_w.value / R:.3f} ms ({n_w.value // R} launches, {fl_w.value / ms_w.value / 1e9:.0f} TFLOP/s, "
      f"{P * 9.5e3 / (ms_w.value / R) / 1e9:.0f} GB/s est.) | layer GEMMs {ms_g.value / R:.3f} ms ({n_g.value // R}) | rest (heads + dgrad chain) "
      f"{tot - ms_w.value / R - ms_g.value / R:.3f} ms")
if os.environ.get("NMX_CHAIN_DBG", "0") != "0":
    T, NLAY = 6, 10
    buf = np.zeros(2 * T * NLAY * 4 * 2, dtype=np.int64)
    L.call("nmx_chain_trace_read", buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), L.i32(buf.size))
    tr = buf.reshape(2, T, NLAY, 4, 2)
    t0 = tr[0, 0, 0, 0, 0]
    print("tile layer |  M:tempty_ok  M:full_ok  M:act_ok(issue)  M:committed |  E:wake  E:layer_done | period")
    prev = None
    for it in range(1, 3):
        for l in range(9):
            m = tr[0, it, l, :, 0] - t0
            e = tr[1, it, l, :, 0] - t0
            per = (m[3] - prev) if prev is not None else 0
            prev = m[3]
            print(f"{it:4d} {l:5d} | {m[0]:11d} {m[1]:9d} {m[2]:15d} {m[3]:12d} | {e[0]:7d} {e[2]:12d} | {per:6d}")
