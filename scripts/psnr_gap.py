"""PSNR gap between the bf16 tensor-core path and the fp32 oracle (north_star: "... with the PSNR gap stated").

Both trainers start from the SAME weights and see the SAME rays, targets and resampling draws every iteration of the
reference's coarse+fine loop (64 + 128 samples); afterwards both render the same held-out rays and are scored against
the synthetic ground truth.  Scene: an analytic radiance field (Gaussian density blob with a position-dependent colour)
rendered through the oracle's own compositing with 192 samples.

    python scripts/psnr_gap.py [--iters 200] [--rays 1024]
Prints one JSON line; the oracle side is CPU torch fp32 (about a second per iteration at 1024 rays).
"""
import argparse
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import models as omodels, rendering as orend, training as otrain  # noqa: E402
from nerf_meets_mlx_b200.models.NeRF import default_args  # noqa: E402
from nerf_meets_mlx_b200.training import NeRFTrainer, assemble_rays  # noqa: E402

N_S, N_I = 64, 128


def analytic_targets(o, d, n=192, near=2.0, far=6.0):
    """Ground truth: sigma(x) = 12 exp(-|x|^2 / 0.6^2), c(x) = 0.5 + 0.5 sin(2.5 x + phase); white background."""
    z = np.linspace(near, far, n, dtype=np.float32)[None, :].repeat(o.shape[0], 0)
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    sigma = 12.0 * np.exp(-np.sum(pts * pts, -1) / 0.36)
    rgb = 0.5 + 0.5 * np.sin(2.5 * pts + np.array([0.0, 2.0, 4.0], np.float32))
    raw = torch.from_numpy(np.concatenate([rgb, sigma[..., None]], -1).astype(np.float32))
    out, _, _, _, _ = orend.raw2outputs(raw, torch.from_numpy(z), torch.from_numpy(d), 0.0, True)
    return out.numpy().astype(np.float32)


def rays_for(rng, B):
    theta = rng.uniform(0, 2 * np.pi, size=B)
    phi = rng.uniform(-0.6, 0.6, size=B)
    cam = 4.0 * np.stack([np.cos(phi) * np.sin(theta), np.cos(phi) * np.cos(theta), np.sin(phi)], -1)
    look = -cam / np.linalg.norm(cam, axis=-1, keepdims=True)
    d = look + rng.normal(scale=0.12, size=(B, 3))
    return cam.astype(np.float32), d.astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--eval-rays", type=int, default=4096)
    a = ap.parse_args()
    torch.set_num_threads(torch.get_num_threads())
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=1, **kw), omodels.NeRF(seed=2, **kw)
    tr = NeRFTrainer(default_args(N_importance=N_I, n_depth_samples=N_S), device="cuda", max_rays=max(a.rays, a.eval_rays))
    tr.coarse.load_reference_parameters(oc.params)
    tr.fine.load_reference_parameters(of.params)
    opt = otrain.AdamMLX(5e-4)
    qf = orend.make_query_fn(10, 4)
    rng = np.random.default_rng(0)
    t_cpu = t_gpu = 0.0
    hist = []
    for it in range(1, a.iters + 1):
        o, d = rays_for(rng, a.rays)
        tgt = analytic_targets(o, d)
        u = rng.random(size=(a.rays, N_I), dtype=np.float32)
        opt.learning_rate = otrain.lr_schedule(it - 1)
        t0 = time.perf_counter()
        ro = otrain.train_iteration(oc, of, opt, o, d, tgt, u, qf, n_samples=N_S)
        t_cpu += time.perf_counter() - t0
        t0 = time.perf_counter()
        rg = tr.train_iteration(*(torch.from_numpy(x).cuda() for x in (o, d, tgt)), u_vals=torch.from_numpy(u).cuda())
        lg = float(rg["loss_fine"])
        t_gpu += time.perf_counter() - t0
        if it % 25 == 0 or it == 1:
            hist.append({"iter": it, "loss_fine_oracle": ro["loss_fine"], "loss_fine_b200": lg})
            print(f"iter {it}: fine loss oracle {ro['loss_fine']:.5f}  b200 {lg:.5f}", file=sys.stderr, flush=True)
    # held-out evaluation through each side's own coarse+fine renderer
    o, d = rays_for(np.random.default_rng(123), a.eval_rays)
    tgt = analytic_targets(o, d)
    u = np.random.default_rng(7).random(size=(a.eval_rays, N_I), dtype=np.float32)
    with torch.no_grad():
        rays_o = otrain.assemble_rays(o, d, 2.0, 6.0)
        res_o = orend.render_rays_eval(rays_o, oc, qf, N_S, network_fine=of, white_bkgd=True, u_vals=u)
        rgb_o = res_o["rgb_map"].numpy()
        rays_g = assemble_rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 2.0, 6.0)
        rgb_g = tr.render_rays_eval(rays_g, u_vals=torch.from_numpy(u).cuda())["rgb_map"].cpu().numpy()
    mse_o, mse_g = float(np.mean((rgb_o - tgt) ** 2)), float(np.mean((rgb_g - tgt) ** 2))
    line = {"iters": a.iters, "rays_per_iter": a.rays, "eval_rays": a.eval_rays,
            "psnr_oracle_fp32_db": float(otrain.psnr(mse_o)), "psnr_b200_bf16_db": float(otrain.psnr(mse_g)),
            "psnr_gap_db": float(otrain.psnr(mse_g) - otrain.psnr(mse_o)),
            "max_abs_rgb_diff": float(np.abs(rgb_o - rgb_g).max()), "cpu_s": t_cpu, "gpu_s": t_gpu, "history": hist}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
