"""Phase-level CUDA-event timing of one C3 training iteration (8192 rays, 64 + 128 samples)."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import ops
from nerf_meets_mlx_b200.models.NeRF import default_args
from nerf_meets_mlx_b200.training import NeRFTrainer, assemble_rays

B, n, N = (int(sys.argv[1]) if len(sys.argv) > 1 else 8192), 64, 128
tr = NeRFTrainer(default_args(N_importance=N, n_depth_samples=n), device="cuda", max_rays=B)
torch.manual_seed(0)
o = torch.randn(B, 3, device="cuda") * 0.1 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda") * 0.2 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
tgt = torch.rand(B, 3, device="cuda")
u = torch.rand(B, N, device="cuda")
rays = assemble_rays(o, d, 2.0, 6.0)
rd = rays[:, 3:6].contiguous()

marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def step(model, z, wb, g):
    Bn, nn = z.shape
    raw = model._fwd_raw(1, rays, z, None, Bn, nn, save=True); mark(f"fwd_train n={nn}")
    rgb, _, _, w, _ = ops.composite_fwd(raw.view(Bn, nn, 4), z, rd, white_bkgd=wb)
    loss, d_rgb = ops.mse_fwd_bwd(rgb, tgt)
    d_raw = ops.composite_bwd(raw.view(Bn, nn, 4), z, rd, d_rgb, white_bkgd=wb); mark(f"composite+mse n={nn}")
    model._bwd_raw(d_raw.view(Bn * nn, 4), Bn * nn, out=g); mark(f"bwd n={nn}")
    tr.optimizer.update(model, g); mark(f"adam n={nn}")
    return w


def iteration():
    marks.clear()
    mark("start")
    z = ops.sample_z(rays[:, 6], rays[:, 7], n); mark("sample_z")
    step(tr.coarse, z, True, tr._g_coarse)
    raw = tr.coarse._fwd_raw(1, rays, z, None, B, n, save=False); mark("fwd_infer n=64 (incl. weight repack)")
    _, _, _, w, _ = ops.composite_fwd(raw.view(B, n, 4), z, rd, white_bkgd=True)
    zf = ops.sample_pdf(z, w, u, want_imp=False)["z_merged"]; mark("composite+sample_pdf")
    step(tr.fine, zf, False, tr._g_fine)


for _ in range(3):
    iteration()
torch.cuda.synchronize()
acc = {}
R = 5
for _ in range(R):
    iteration()
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
tot = sum(acc.values()) / R
for k, v in acc.items():
    print(f"{k:45s} {v / R:8.3f} ms  {100 * v / R / tot:5.1f}%")
print(f"{'total':45s} {tot:8.3f} ms  -> {B / tot * 1e3:.0f} rays/s")
