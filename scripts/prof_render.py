"""One 800x800 coarse+fine frame through NeRFTrainer.render_rays_eval (the C5 path of bench.py), for ncu launch lists."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models.NeRF import default_args
from nerf_meets_mlx_b200.training import NeRFTrainer
tr = NeRFTrainer(default_args(N_importance=128, n_depth_samples=64), device="cuda", max_rays=32768)
n = 640000
g = torch.Generator(device="cuda").manual_seed(0)
o = torch.randn(n, 3, device="cuda", generator=g) * 0.1 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
d = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g) - torch.tensor([0.0, 0.0, 3.0], device="cuda"), dim=-1)
rays = torch.cat([o, d, torch.full((n, 1), 2.0, device="cuda"), torch.full((n, 1), 6.0, device="cuda"), d], -1).contiguous()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for r in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in range(0, n, 32768):
            tr.render_rays_eval(rays[s:s + 32768])
        b.record(); torch.cuda.synchronize()
        print(f"frame {r}: {a.elapsed_time(b):.2f} ms  {n * 256 / a.elapsed_time(b) / 1e3:.1f} Msamples/s", flush=True)
