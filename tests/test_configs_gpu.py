"""Config-level parity (BASELINE.json `configs`): the whole training step of C1 (image learning), C2 (coarse-only NeRF)
and C4 (multiresolution hash grid + tiny MLP) through the CUDA path against the fp32 oracle on the same seeded inputs.
C3 (coarse+fine) is tests/test_training_gpu.py; C5 (render) is test_render_api_shapes_and_coarse_fine."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoding as oenc, models as omodels, rendering as orend, training as otrain  # noqa: E402


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


def test_c1_image_learning_steps_vs_oracle():
    """C1 (__viser_image_learning.py:188-236): SinusoidalEncoding(2, 10, 0, 8) of INTEGER pixel coordinates -> NeRF(40 ->
    8x256 -> 3, skip at 4), loss = MSE, Adam(lr 1e-3, betas (0.9, 0.99)) without bias correction; 1024 pixels per step of a
    synthetic 256x256 image.  Four consecutive steps: losses and the updated parameters track the oracle."""
    from nerf_meets_mlx_b200.encoding import SinusoidalEncoding
    from nerf_meets_mlx_b200.models import NeRF
    from nerf_meets_mlx_b200.models.NeRF import AdamMLX
    kw = dict(n_layers=8, width_layers=256, channel_input=40, channel_input_views=0, channel_output=3,
              list_skip_connection_layers=[4], is_use_view_directions=False)
    ref = omodels.NeRF(seed=21, **kw)
    net = NeRF(device="cuda", **kw)
    net.load_reference_parameters(ref.params)
    enc = SinusoidalEncoding(2, 10, min_freq_exp=0.0, max_freq_exp=8.0)
    opt_ref = otrain.AdamMLX(1e-3, betas=(0.9, 0.99))
    opt = AdamMLX(1e-3, betas=(0.9, 0.99))
    rng = np.random.default_rng(0)
    yy, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    img = np.stack([0.5 + 0.5 * np.sin(xx / 17.0), 0.5 + 0.5 * np.cos(yy / 23.0), ((xx // 32 + yy // 32) % 2)], -1).astype(np.float32)
    coords = np.stack([yy, xx], -1).reshape(-1, 2)
    perm = rng.permutation(256 * 256)
    for step in range(4):
        sel = perm[step * 1024:(step + 1) * 1024]
        X, y = coords[sel], img.reshape(-1, 3)[sel]
        loss_ref, _ = otrain.image_step(ref, opt_ref, X, y)
        net.flat.requires_grad_(True)
        net.flat.grad = None
        pred = net.forward(enc(torch.from_numpy(X).cuda()))
        loss = torch.mean((pred - torch.from_numpy(y).cuda()) ** 2)
        loss.backward()
        opt.update(net, net.flat.grad)
        assert _rel(float(loss), loss_ref) < 1e-2, (step, float(loss), loss_ref)
    got = net.split_flat(net.flat.data)
    for name, p_ref in ref.params.items():
        d = (got[name].cpu() - p_ref.detach()).norm() / p_ref.detach().norm().clamp_min(1e-12)
        assert float(d) < 2e-2, (name, float(d))


def test_c2_coarse_only_iteration_vs_oracle():
    """C2: coarse NeRF only (N_importance = 0 -> output_ch 4, no fine net, one optimiser step per iteration)."""
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer
    B, n = 192, 64
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=4,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc = omodels.NeRF(seed=31, **kw)
    tr = NeRFTrainer(default_args(N_importance=0, n_depth_samples=n), device="cuda", max_rays=B)
    assert tr.fine is None
    tr.coarse.load_reference_parameters(oc.params)
    opt = otrain.AdamMLX(5e-4)
    qf = orend.make_query_fn(10, 4)
    rng = np.random.default_rng(4)
    for it in range(2):
        o = (rng.uniform(-0.5, 0.5, size=(B, 3)) + np.array([0, 0, 4.0])).astype(np.float32)
        d = rng.standard_normal(size=(B, 3)).astype(np.float32)
        d[:, 2] = -np.abs(d[:, 2]) - 1.0
        tgt = rng.random(size=(B, 3)).astype(np.float32)
        opt.learning_rate = otrain.lr_schedule(it)
        r_ref = otrain.train_iteration(oc, None, opt, o, d, tgt, None, qf, n_samples=n)
        r = tr.train_iteration(*(torch.from_numpy(a).cuda() for a in (o, d, tgt)))
        assert "loss_fine" not in r
        assert _rel(float(r["loss_coarse"]), r_ref["loss_coarse"]) < 1e-2
    got = tr.coarse.split_flat(tr.coarse.flat.data)
    for name, p_ref in oc.params.items():
        d = (got[name].cpu() - p_ref.detach()).norm() / p_ref.detach().norm().clamp_min(1e-12)
        assert float(d) < 2e-2, (name, float(d))


def test_c4_hashgrid_tiny_mlp_step_vs_oracle():
    """C4: MultiHashEncoding(L=16, T=2^19, F=2, 16..2048) -> NeRF(n_layers=2, width 64, 32 -> 4) (the "tiny MLP" of
    SURVEY 8d), loss = MSE on the raw outputs; one training step: loss, the table gradient (scatter with vector atomics)
    and the MLP gradients against the fp32/fp64 oracle.  The hash indices are bit-exact (test_hash_bit_exact)."""
    from nerf_meets_mlx_b200.encoding import MultiHashEncoding
    from nerf_meets_mlx_b200.models import NeRF
    L, F, T = 16, 2, 19
    kw = dict(n_layers=2, width_layers=64, channel_input=L * F, channel_input_views=0, channel_output=4,
              list_skip_connection_layers=[], is_use_view_directions=False)
    enc = MultiHashEncoding(3, L, 16, 2048, F, T, hash_init_scale=1e-1, device="cuda", seed=3)  # large init: visible signal
    ref = omodels.NeRF(seed=41, **kw)
    net = NeRF(device="cuda", **kw)
    net.load_reference_parameters(ref.params)
    P = 8192
    rng = np.random.default_rng(5)
    x = rng.random(size=(P, 3), dtype=np.float32)
    tgt = rng.random(size=(P, 4)).astype(np.float32)
    tables = enc.hash_table.detach().cpu().numpy()
    res = enc.scaled_res.cpu().numpy()
    # oracle: encode (fp32, reference interpolation order) -> MLP -> MSE, gradients by autograd + the table scatter
    feat_ref = torch.from_numpy(oenc.hashgrid_encode(x, tables, res, T)).requires_grad_(True)
    ref.requires_grad_(True)
    y_ref = ref.forward(feat_ref)
    loss_ref = torch.mean((y_ref - torch.from_numpy(tgt)) ** 2)
    names = list(ref.params.keys())
    grads_ref = torch.autograd.grad(loss_ref, [feat_ref] + [ref.params[k] for k in names])
    g_tab_ref = oenc.hashgrid_backward(x, grads_ref[0].numpy(), L, F, res, T)
    # CUDA path
    enc.hash_table.grad = None
    net.flat.requires_grad_(True)
    feat = enc(torch.from_numpy(x).cuda())
    np.testing.assert_allclose(feat.detach().cpu().numpy(), feat_ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    y = net.forward(feat)
    loss = torch.mean((y - torch.from_numpy(tgt).cuda()) ** 2)
    loss.backward()
    assert _rel(float(loss), float(loss_ref)) < 1e-2
    got = net.split_flat(net.flat.grad)
    for k, g in zip(names, grads_ref[1:]):
        d = (got[k].cpu() - g).norm() / g.norm().clamp_min(1e-20)
        assert float(d) < 2e-2, (k, float(d))
    g_tab = enc.hash_table.grad.cpu().numpy().astype(np.float64)
    d = np.linalg.norm(g_tab - g_tab_ref) / np.linalg.norm(g_tab_ref)
    assert d < 5e-2, d  # bf16 data gradient through the MLP (ReLU masks of near-zero units flip); the scatter itself is fp32-exact (test_hashgrid_fwd_bwd)
    # only table entries some query actually touches receive gradient (a few may round to exactly zero in bf16)
    nz, nz_ref = np.abs(g_tab).sum(-1) > 0, np.abs(g_tab_ref).sum(-1) > 0
    assert not np.any(nz & ~nz_ref)
    assert nz.sum() > 0.98 * nz_ref.sum()


@pytest.mark.parametrize("cfg_id", [0, 2, 3])
def test_mlp_input_gradient_vs_oracle(cfg_id, measured):
    """nmx_mlp_bwd_input: d loss / d (encoded position inputs) through the fused chain (view-dir 8x256 net, skip
    connection: two contributions), the layer-by-layer path with a skip (image net) and the tiny 3x64 net."""
    from test_mlp_gpu import CFGS, emulated_forward, make_pair
    cfg = CFGS[cfg_id]
    torch.manual_seed(cfg_id)
    ref, net = make_pair(**cfg)
    P = 777
    n_pos = cfg["channel_input"]
    cin = n_pos + (cfg["channel_input_views"] if cfg["is_use_view_directions"] else 0)
    x = torch.randn(P, cin).clamp(-1, 1)
    g_out = torch.randn(P, 4 if cfg["is_use_view_directions"] else cfg["channel_output"])
    x_ref = x.clone().requires_grad_(True)
    (ref.forward(x_ref) * g_out).sum().backward()
    x_emu = x.clone().requires_grad_(True)   # same arithmetic as the kernels (bf16 storage points): tight tolerance
    (emulated_forward(ref, x_emu) * g_out).sum().backward()
    x_dev = x.cuda().requires_grad_(True)
    net.flat.requires_grad_(True)
    (net.forward(x_dev) * g_out.cuda()).sum().backward()
    got = x_dev.grad.cpu()[:, :n_pos]
    assert float((got - x_emu.grad[:, :n_pos]).norm() / x_emu.grad[:, :n_pos].norm()) < 1e-2
    # vs the fp32 oracle: units whose pre-activation rounds across zero flip their ReLU mask (same bound as the
    # parameter gradients in test_mlp_gpu.py)
    e = measured(f"mlp_input_grad_vs_fp32_oracle/cfg{cfg_id}", float((got - x_ref.grad[:, :n_pos]).norm() / x_ref.grad[:, :n_pos].norm()))
    assert e < 1.5e-1
    if cin > n_pos:  # declared: view-direction inputs receive no gradient (nothing learnable feeds them)
        assert float(x_dev.grad[:, n_pos:].abs().max()) == 0.0
