"""MMA issue-rate probe + event trace of the fused forward chain (CTA 0).
The event trace (NMX_CHAIN_DBG=4) exists only in an experiments build: python -m nerf_meets_mlx_b200.build --experiments"""
import ctypes
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L
from nerf_meets_mlx_b200.models.NeRF import NeRF

lib = L.lib()
out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
for ctas in (1, 148):
    for N, slabs in ((256, 1), (256, 3), (128, 3)):
        for _ in range(2):
            L.call("nmx_diag_mma_rate", L.i32(N), L.i32(20000), L.i32(slabs), L.i32(ctas), L.ptr(out), L.stream(), L.i32(0))
        torch.cuda.synchronize()
        o = out.cpu().numpy().reshape(-1, 2)[:ctas]
        clk, ns = o[:, 0].mean() / 20000, o[:, 1].mean() / 20000
        print(f"mma_rate ctas={ctas:3d} N={N} slabs={slabs}: {clk:7.1f} clk/MMA  {ns:7.1f} ns/MMA  -> {clk / ns * 1e3:6.0f} MHz, "
              f"{ctas * 2 * 128 * N * 16 / ns / 1e3:7.1f} TFLOP/s", flush=True)

if os.environ.get("NMX_CHAIN_DBG", "0") != "0":
    B, n = 2368, 64  # 148 CTAs x 8 tiles
    net = NeRF(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
               list_skip_connection_layers=[4], is_use_view_directions=True, n_freqs_pos=10, n_freqs_dir=4)
    torch.manual_seed(0)
    o3 = torch.randn(B, 3, device="cuda")
    d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda"), dim=-1)
    rays = torch.cat([o3, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
    z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
    for save in (False, True):
        net.reserve(B * n, training=True)
        for _ in range(3):
            net._fwd_raw(1, rays, z, None, B, n, save=save)
        torch.cuda.synchronize()
        T, NLAY = 6, 10
        buf = np.zeros(2 * T * NLAY * 4 * 2, dtype=np.int64)
        L.call("nmx_chain_trace_read", buf.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), L.i32(buf.size))
        tr = buf.reshape(2, T, NLAY, 4, 2)
        t0 = tr[0, 0, 0, 0, 0]
        g0 = tr[0, 0, 0, 0, 1]
        print(f"--- chain trace save={save} (clocks rel. to start; M* = MMA thread, E* = epilogue warp 2)")
        print("tile layer |  M:tempty_ok  M:full_ok  M:act_ok(issue)  M:committed |  E:wake  E:ld_done  E:layer_done  E:store_done_ok(start) | period")
        prev = None
        for it in range(1, 4):
            for l in range(NLAY):
                m = tr[0, it, l, :, 0] - t0
                e = tr[1, it, l, :, 0] - t0
                per = (m[3] - prev) if prev is not None else 0
                prev = m[3]
                print(f"{it:4d} {l:5d} | {m[0]:11d} {m[1]:9d} {m[2]:15d} {m[3]:12d} | {e[0]:7d} {e[1]:9d} {e[2]:12d} {e[3]:12d} | {per:6d}")
        dc = tr[0, T - 1, NLAY - 1, 3, 0] - t0
        dg = tr[0, T - 1, NLAY - 1, 3, 1] - g0
        print(f"clock: {dc} clks in {dg} ns -> {dc / dg * 1e3:.0f} MHz; per tile {dc / (T - 1 + 1):.0f} clks")
