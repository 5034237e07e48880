"""tcgen05 issue-rate probe: plain / with a commit per 4 MMAs / with commit + barrier poll."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L
import ctypes
out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
for ctas in (148,):
    for N in (256, 128, 64):
        for mode in (0, 1, 2, 5, 6):
            for _ in range(2):
                L.call("nmx_diag_mma_rate", L.i32(N), L.i32(20000), L.i32(3), L.i32(ctas), L.ptr(out), L.stream(), L.i32(mode))
            torch.cuda.synchronize()
            o = out.cpu().numpy().reshape(-1, 2)[:ctas]
            clk, ns = o[:, 0].mean() / 20000, o[:, 1].mean() / 20000
            print(f"mma_rate ctas={ctas:3d} N={N:3d} mode={mode}: {clk:7.1f} clk/MMA ({4 * clk:6.1f} per 4)  {clk / ns * 1e3:6.0f} MHz "
                  f"{ctas * 2 * 128 * N * 16 / ns / 1e3:7.1f} TFLOP/s", flush=True)
