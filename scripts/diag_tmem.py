"""TMEM read-rate probe: how many bytes per clock tcgen05.ld delivers per SM (the chain's epilogue floor)."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L

out = torch.zeros(2 * 148, dtype=torch.int64, device="cuda")
for ctas in (1, 148):
    for warps in (4, 8, 16):
        for mode in (0, 1):
            for batch in (1, 2, 4):
                for _ in range(2):
                    L.call("nmx_diag_tmem_ld_rate", L.i32(4096), L.i32(warps), L.i32(mode), L.i32(batch), L.i32(ctas), L.ptr(out), L.stream())
                torch.cuda.synchronize()
                o = out.cpu().numpy().reshape(-1, 2)[:ctas]
                print(f"ctas={ctas:3d} warps={warps:2d} ld=x{32 if mode else 16} batch={batch}: {o[:, 1].mean() / o[:, 0].mean():7.1f} B/clk/SM"
                      f"  ({o[:, 0].mean() / 4096:6.1f} clk per load round)", flush=True)
