"""End-to-end: one reference training iteration (coarse step, coarse re-forward, detached resample + merge, fine step)
through the CUDA path vs the fp32 oracle, and the mirrored render API vs golden vectors of the reference's own code."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models as omodels, rendering as orend, training as otrain  # noqa: E402


def _rays(B, seed=0):
    rng = np.random.default_rng(seed)
    o = (rng.uniform(-0.5, 0.5, size=(B, 3)) + np.array([0, 0, 4.0])).astype(np.float32)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - 1.0
    d /= np.linalg.norm(d, axis=-1, keepdims=True) * 0.9
    target = rng.random(size=(B, 3)).astype(np.float32)
    return o, d.astype(np.float32), target


def test_train_iteration_vs_oracle(measured):
    from nerf_meets_mlx_b200.training import NeRFTrainer
    from nerf_meets_mlx_b200.models.NeRF import default_args
    B, n, N = 96, 64, 128
    args = default_args(N_importance=N, n_depth_samples=n)
    tr = NeRFTrainer(args, max_rays=B)
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=1, **kw), omodels.NeRF(seed=2, **kw)
    # The reference's last-bin delta of 1e10 makes alpha_last = step(sigma_last): a density of ~0 (random init) flips
    # whole rays on a 1-ulp change.  Keep sigma away from 0 so the loss is a smooth function of the bf16 rounding.
    for m in (oc, of):
        m.params["alpha_linear.bias"] = m.params["alpha_linear.bias"] * 0 + 0.5
    tr.coarse.load_reference_parameters(oc.params)
    tr.fine.load_reference_parameters(of.params)
    o, d, target = _rays(B)
    u = np.random.default_rng(5).random(size=(B, N), dtype=np.float32)
    opt = otrain.AdamMLX(args.lrate, shared_state=True)
    qf = orend.make_query_fn(10, 4)
    ref = otrain.train_iteration(oc, of, opt, o, d, target, u, qf, n_samples=n, near=2.0, far=6.0, white_bkgd=True)
    out = tr.train_iteration(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(target).cuda(),
                             u_vals=torch.from_numpy(u).cuda())
    assert abs(out["loss_coarse"].item() - ref["loss_coarse"]) / ref["loss_coarse"] < 1e-2
    assert abs(out["loss_fine"].item() - ref["loss_fine"]) / ref["loss_fine"] < 1e-2
    # gradients (bf16 operands vs the fp32 oracle; the tight check lives in test_mlp_gpu.py)
    gc = tr.coarse.split_flat(tr._g_coarse)
    for name, g in ref["grads_coarse"].items():
        e = measured("train_iteration_96rays/coarse_grad_vs_fp32_oracle", float((gc[name].cpu() - g).norm() / (g.norm() + 1e-20)))
        assert e < 5.5e-2, (name, e)  # 2 x the measured 2.7e-2 (profiles/r2_test_measurements.json); full size: 6.4e-3
    # fine depths: sorted, same count, close to the oracle's (weights come from a bf16 coarse net)
    zf = out["z_fine"].cpu().numpy()
    assert zf.shape == ref["z_fine"].shape and np.all(np.diff(zf, axis=-1) >= 0)
    close = np.mean(np.abs(zf - ref["z_fine"]) < 2e-2)
    assert close > 0.97, close
    # parameters moved by the MLX-style Adam (no bias correction): first step has |dp| ~ lr*0.1/sqrt(0.001)
    lr_eff = args.lrate * 0.1 / np.sqrt(0.001)
    p_now = tr.coarse.named_reference_parameters()["list_linears_pos.3.weight"].cpu()
    p_ref = oc.params["list_linears_pos.3.weight"]  # already updated by the oracle's Adam
    assert float((p_now - p_ref).abs().max()) <= 2.1 * lr_eff


def test_losses_decrease_over_iterations():
    from nerf_meets_mlx_b200.training import NeRFTrainer
    from nerf_meets_mlx_b200.models.NeRF import default_args
    tr = NeRFTrainer(default_args(N_importance=64, n_depth_samples=32), max_rays=256, shared_adam_state=False)
    with torch.no_grad():  # avoid the dead-density start (relu(sigma) == 0 everywhere has zero gradient by construction)
        for m in (tr.coarse, tr.fine):
            m.alpha_linear.bias.fill_(0.5)
            m.mark_params_updated()
    o, d, target = _rays(256, seed=3)
    target[:] = np.array([0.2, 0.5, 0.8], np.float32)
    O, D, T = (torch.from_numpy(a).cuda() for a in (o, d, target))
    first = last = None
    for i in range(60):
        out = tr.train_iteration(O, D, T)
        if i == 0:
            first = (out["loss_coarse"].item(), out["loss_fine"].item())
        last = (out["loss_coarse"].item(), out["loss_fine"].item())
    assert last[0] < 0.5 * first[0] and last[1] < 0.5 * first[1], (first, last)


def _load(net, g, prefix):
    net.load_reference_parameters({k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)})


def test_render_api_shapes_and_coarse_fine(measured):
    """Mirror API (create_NeRF -> render) end to end at small size vs the oracle with identical weights."""
    from nerf_meets_mlx_b200.models.NeRF import create_NeRF, default_args
    from nerf_meets_mlx_b200.rendering import render as R
    args = default_args(N_importance=128)
    kw_train, kw_test, _, _ = create_NeRF(args)
    assert kw_train is kw_test and kw_test["perturb"] is False  # reference alias quirk
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=11, **kw), omodels.NeRF(seed=12, **kw)
    for m in (oc, of):
        m.params["alpha_linear.bias"] = m.params["alpha_linear.bias"] * 0 + 0.5
    kw_test["network_coarse"].load_reference_parameters(oc.params)
    kw_test["network_fine"].load_reference_parameters(of.params)
    H, W, focal = 12, 16, 20.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = orend.pose_spherical(40.0, -30.0, 4.0)
    u = np.random.default_rng(1).random(size=(H * W, 128), dtype=np.float32)
    with torch.no_grad():
        rgb, disp, acc, extras = R.render(H, W, K, chunk=64, c2w=torch.from_numpy(c2w[:3, :4]).cuda(), near=2.0,
                                          far=6.0, u_vals=torch.from_numpy(u).cuda(), **kw_test)
    assert rgb.shape == (H, W, 3) and disp.shape == (H, W, 1) and acc.shape == (H, W, 1)
    assert extras["z_vals"].shape == (H, W, 64) and extras["weights"].shape == (H, W, 64, 1)
    qf = orend.make_query_fn(10, 4)
    ref = orend.render(H, W, K, chunk=64, c2w=c2w[:3, :4], near=2.0, far=6.0, use_viewdirs=True,
                       render_rays_func=orend.render_rays_eval, network_coarse=oc, network_fine=of,
                       network_query_fn=qf, n_depth_samples=64, N_importance=128, white_bkgd=True, u_vals=u)
    for got, want, name in ((rgb, ref[0], "rgb"), (acc, ref[2], "acc"), (extras["rgb_coarse"], ref[3]["rgb_coarse"], "rgb_coarse")):
        want = want.numpy()
        err = measured(f"render_12x16/{name}", np.abs(got.cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-6))
        assert err < 1e-2, (name, err)  # north_star: 1e-2 for bf16 MLP outputs


@pytest.mark.gpu
def test_cuda_graph_replay_matches_eager():
    """The captured-and-replayed iteration (device-resident learning rate) follows the eager one step for step."""
    import torch
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer
    B = 256
    torch.manual_seed(0)
    data = [(torch.randn(B, 3, device="cuda") * 0.1 + torch.tensor([0.0, 0.0, 4.0], device="cuda"),
             torch.nn.functional.normalize(torch.randn(B, 3, device="cuda") * 0.2 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1),
             torch.rand(B, 3, device="cuda"), torch.rand(B, 128, device="cuda")) for _ in range(4)]
    losses = []
    for graphed in (False, True):
        tr = NeRFTrainer(default_args(N_importance=128, n_depth_samples=64, lrate_decay=0.001), device="cuda", max_rays=B,
                         use_cuda_graph=graphed)
        ls = []
        for o, d, t, u in data:
            out = tr.train_iteration(o, d, t, u_vals=u)
            ls.append((float(out["loss_coarse"]), float(out["loss_fine"])))
        assert (tr._graph is not None) == graphed
        losses.append(ls)
    for (ce, fe), (cg, fg) in zip(*losses):
        assert abs(ce - cg) <= 2e-3 * abs(ce) and abs(fe - fg) <= 2e-3 * abs(fe), (losses[0], losses[1])


def test_render_from_pose_golden(golden):
    """render(H, W, K, c2w=...) end to end -- ray generation kernel, ray assembly, coarse pass -- against the
    reference's own render() output (small 64-wide net, fp32 reference vs bf16 tensor path: 1e-2)."""
    from nerf_meets_mlx_b200.models.NeRF import create_NeRF, default_args
    from nerf_meets_mlx_b200.rendering import render
    g = golden("render_full")
    kw, _, _, _ = create_NeRF(default_args(N_importance=0, netwidth=128, n_depth_samples=16))
    net = kw["network_coarse"]
    net.load_reference_parameters({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")})
    H, W = int(g["H"]), int(g["W"])
    with torch.no_grad():
        out = render.render(H, W, g["K"], chunk=20, c2w=g["c2w"][:3, :4], ndc=False, near=2.0, far=6.0, use_viewdirs=True,
                            network_query_fn=kw["network_query_fn"], network_coarse=net, n_depth_samples=16,
                            white_bkgd=True, render_rays_func=render.render_rays)
    assert out[0].shape == (H, W, 3)
    assert float(np.abs(out[0].cpu().numpy() - g["rgb"]).max()) < 1e-2 * max(1.0, float(np.abs(g["rgb"]).max()))
    assert float(np.abs(out[2].cpu().numpy().reshape(H, W) - g["acc"].reshape(H, W)).max()) < 1e-2


def test_train_iteration_from_pixels_matches_ray_inputs():
    """train_iteration_pixels (device-side ray selection) == train_iteration on the same rays assembled on the host."""
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer
    from oracle import rendering as orend
    rng = np.random.default_rng(9)
    H = W = 64
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = orend.pose_spherical(10.0, -30.0, 4.0)
    img = rng.random(size=(H * W, 3)).astype(np.float32)
    pix = rng.choice(H * W, size=512, replace=False).astype(np.int32)
    u = torch.rand(512, 32, generator=torch.Generator().manual_seed(0)).cuda()
    losses = []
    for mode in (0, 1):
        torch.manual_seed(0)
        tr = NeRFTrainer(default_args(N_importance=32, n_depth_samples=16), device="cuda", max_rays=512)
        if mode == 0:
            r = tr.train_iteration_pixels(H, W, K, c2w, torch.from_numpy(pix).cuda(), torch.from_numpy(img).cuda(), u_vals=u)
        else:
            ro, rd = orend.get_rays(H, W, K, c2w[:3, :4])
            ro = torch.from_numpy(np.reshape(ro, (-1, 3))[pix].astype(np.float32)).cuda()
            rd = torch.from_numpy(np.reshape(rd, (-1, 3))[pix].astype(np.float32)).cuda()
            r = tr.train_iteration(ro, rd, torch.from_numpy(img[pix]).cuda(), u_vals=u)
        losses.append((float(r["loss_coarse"]), float(r["loss_fine"])))
    # same rays and targets: the coarse loss is a deterministic forward; the fine loss sees the coarse net after an
    # update whose weight gradients were reduced with fp32 atomics (order-dependent in the last bits)
    assert abs(losses[0][0] - losses[1][0]) <= 1e-6 * abs(losses[1][0])
    assert abs(losses[0][1] - losses[1][1]) <= 1e-3 * abs(losses[1][1])


def test_turntable_loop_matches_single_renders():
    """render_turntable (__test_nerf.py:326-341) == rendering each pose of the loader's render_poses circle on its own."""
    from nerf_meets_mlx_b200.models.NeRF import create_NeRF, default_args
    from nerf_meets_mlx_b200.ops.pose import pose_spherical
    from nerf_meets_mlx_b200.rendering import render as R
    from nerf_meets_mlx_b200.rendering.turntable import render_turntable, to8b
    kw_train, kw_test, _, _ = create_NeRF(default_args(N_importance=16, n_depth_samples=16))
    H, W, focal = 10, 12, 15.0
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    poses = torch.stack([torch.as_tensor(np.asarray(pose_spherical(float(a), -30.0, 4.0), dtype=np.float32)) for a in (-180.0, -60.0, 60.0)])
    u = torch.rand(H * W, 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    kw = dict(kw_test, near=2.0, far=6.0, u_vals=u)
    got = []
    frames = render_turntable(H, W, K, poses, kw, writer=got.append)
    assert frames.shape == (3, H, W, 3) and frames.dtype == np.uint8 and len(got) == 3
    for i in range(3):
        with torch.no_grad():
            rgb = R.render(H, W, K, c2w=poses[i][:3, :4].cuda(), **kw)[0]
        np.testing.assert_array_equal(frames[i], to8b(rgb))
        np.testing.assert_array_equal(got[i], frames[i])
