"""NeRF MLP (bf16 tcgen05 GEMMs, fp32 accumulate) vs the fp32 oracle (torch-CPU restatement of NeRF.forward).
north_star tolerance: 1e-2 relative for bf16 MLP outputs.  Gradients: <= 1e-2 normwise against a reference that
emulates the kernels' bf16 storage points; against the fp32 oracle the measured error is recorded per case (see
GRAD_VS_FP32 below) -- the parity claim for gradients is the full-size test (tests/test_fullsize_gpu.py: <= 2.4e-2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import models as omodels, encoding as oenc  # noqa: E402


def rel_max(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def rel_norm(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


class _RoundBF16(torch.autograd.Function):
    """bf16 rounding of a value on the way forward AND of its gradient on the way back: exactly where the kernels
    store bf16 (activations h_l / feature / hd, and the back-propagated dY buffers)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    """bf16 rounding of a weight operand with a straight-through gradient (master weights stay fp32)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def emulated_forward(ref, x):
    """The oracle's NeRF.forward with bf16 rounding inserted at the kernel's storage points (fp32 accumulate)."""
    p = ref.params
    rb, rw = _RoundBF16.apply, _RoundFwd.apply
    x = x.bfloat16().float()
    if ref.use_dirs:
        input_pos, input_dir = x[..., :ref.channel_input_pos], x[..., ref.channel_input_pos:]
    else:
        input_pos = x
    h = input_pos
    for i in range(ref.D):
        h = rb(torch.relu(h @ rw(p[f"list_linears_pos.{i}.weight"]).T + p[f"list_linears_pos.{i}.bias"]))
        if i in ref.skips:
            h = torch.cat([input_pos, h], dim=-1)
    if ref.use_dirs:
        alpha = h @ p["alpha_linear.weight"].T + p["alpha_linear.bias"]
        feat = rb(h @ rw(p["feature_linear.weight"]).T + p["feature_linear.bias"])
        h = torch.cat([feat, input_dir], dim=-1)
        h = rb(torch.relu(h @ rw(p["list_linears_dir.0.weight"]).T + p["list_linears_dir.0.bias"]))
        rgb = h @ p["rgb_linear.weight"].T + p["rgb_linear.bias"]
        return torch.cat([rgb, alpha], dim=-1)
    return h @ p["output_linear.weight"].T + p["output_linear.bias"]


def make_pair(**kw):
    from nerf_meets_mlx_b200.models import NeRF
    ref = omodels.NeRF(seed=3, **kw)
    net = NeRF(device="cuda", **kw)
    net.load_reference_parameters(ref.params)
    return ref, net


CFGS = [
    dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
         list_skip_connection_layers=[4], is_use_view_directions=True),
    dict(n_layers=8, width_layers=128, channel_input=63, channel_input_views=27, channel_output=5,
         list_skip_connection_layers=[4], is_use_view_directions=True),
    dict(n_layers=8, width_layers=256, channel_input=40, channel_input_views=0, channel_output=3,
         list_skip_connection_layers=[4], is_use_view_directions=False),
    dict(n_layers=3, width_layers=64, channel_input=32, channel_input_views=0, channel_output=4,
         list_skip_connection_layers=[], is_use_view_directions=False),
]


# Normwise gradient error vs the fp32 ORACLE on these micro cases (random weights, random +-1 inputs, random output
# gradients, 128 / 1000 points): measured 0.06-0.14 on B200 (profiles/r2_test_measurements.json) -- almost all of it ReLU
# masks of near-zero pre-activations flipping under bf16 rounding, which a handful of points cannot average out (the same
# kernels give <= 1.2e-2 on a real 8192-ray step, tests/test_fullsize_gpu.py, and <= 1e-2 against the bf16-emulating
# reference below).  The bound is 1.1x / 1.4x the measured maxima; it is a regression guard, not the parity claim.
GRAD_VS_FP32 = {1000: 1.5e-1, 128: 1.5e-1}


@pytest.mark.parametrize("cfg", CFGS)
@pytest.mark.parametrize("P", [1000, 128])
def test_forward_backward_vs_oracle(cfg, P, measured):
    torch.manual_seed(P)
    ref, net = make_pair(**cfg)
    cin = cfg["channel_input"] + (cfg["channel_input_views"] if cfg["is_use_view_directions"] else 0)
    x = torch.randn(P, cin).clamp(-1, 1)
    ref.requires_grad_(True)
    y_ref = ref.forward(x)
    y = net.forward(x.cuda())
    assert y.shape == y_ref.shape
    e = rel_max(y.detach().cpu(), y_ref.detach())
    assert e < 1e-2, f"forward rel err {e}"
    g_out = torch.randn_like(y_ref)
    names = list(ref.params.keys())
    g_ref = torch.autograd.grad((y_ref * g_out).sum(), [ref.params[n] for n in names])
    # same arithmetic as the kernels (bf16 storage points, fp32 accumulate): tight tolerance
    y_emu = emulated_forward(ref, x)
    g_emu = torch.autograd.grad((y_emu * g_out).sum(), [ref.params[n] for n in names])
    assert rel_max(y.detach().cpu(), y_emu.detach()) < 2e-3
    (y * g_out.cuda()).sum().backward()
    got = net.split_flat(net.flat.grad)
    for n, gr, ge in zip(names, g_ref, g_emu):
        en = rel_norm(got[n].cpu(), ge)
        assert en < 1e-2, f"grad {n}: normwise rel err vs bf16-emulating reference {en}"
        ef = measured(f"mlp_grad_vs_fp32_oracle/W{cfg['width_layers']}_P{P}", rel_norm(got[n].cpu(), gr))
        assert ef < GRAD_VS_FP32[P], f"grad {n}: normwise rel err vs fp32 oracle {ef}"


def test_forward_golden_noview(golden):
    from nerf_meets_mlx_b200.models import NeRF
    g = golden("nerf_forward")
    net = NeRF(n_layers=8, width_layers=64, channel_input=63, channel_input_views=27, channel_output=5,
               list_skip_connection_layers=[4], is_use_view_directions=False)
    net.load_reference_parameters({k[2:]: g[k] for k in g.files if k.startswith("n/")})
    with torch.no_grad():
        y = net.forward(torch.from_numpy(g["x"][:, :63]).cuda()).cpu().numpy()
    ref = g["y_n"]
    assert np.abs(y - ref).max() / np.abs(ref).max() < 1e-2


def test_fused_rays_path_matches_oracle_run_model():
    from nerf_meets_mlx_b200.models import NeRF
    cfg = CFGS[0]
    ref = omodels.NeRF(seed=5, **cfg)
    net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, **cfg)
    net.load_reference_parameters(ref.params)
    rng = np.random.default_rng(0)
    B, n = 37, 64
    o = rng.uniform(-1, 1, size=(B, 3)).astype(np.float32)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    vd = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.concatenate([o, d, np.full((B, 1), 2, np.float32), np.full((B, 1), 6, np.float32), vd], -1).astype(np.float32)
    z = np.sort(rng.uniform(2, 6, size=(B, n)).astype(np.float32), -1)
    pos = o[:, None, :] + z[:, :, None] * d[:, None, :]
    y_ref = omodels.run_model(pos, 10, vd, 4, ref)
    with torch.no_grad():
        y = net.forward_rays(torch.from_numpy(rays).cuda(), torch.from_numpy(z).cuda())
    assert y.shape == y_ref.shape
    e = rel_max(y.cpu(), y_ref)
    assert e < 1e-2, e
    # inference chunking (P > 65536) gives the same numbers as one training-mode pass
    B2 = 1100
    rays2 = torch.from_numpy(np.tile(rays, (B2 // B + 1, 1))[:B2]).cuda()
    z2 = torch.from_numpy(np.tile(z, (B2 // B + 1, 1))[:B2]).cuda()
    with torch.no_grad():
        y2 = net.forward_rays(rays2, z2)
    assert torch.equal(y2[:B], y)


def test_sinusoidal_fused_image_net():
    from nerf_meets_mlx_b200.models import NeRF
    from nerf_meets_mlx_b200.encoding import SinusoidalEncoding
    cfg = CFGS[2]
    ref = omodels.NeRF(seed=9, **cfg)
    net = NeRF(device="cuda", n_freqs_pos=10, **cfg)
    net.load_reference_parameters(ref.params)
    X = torch.stack(torch.meshgrid(torch.arange(0, 256, 9), torch.arange(0, 256, 11), indexing="ij"), -1).reshape(-1, 2)
    enc = SinusoidalEncoding(2, 10, min_freq_exp=0.0, max_freq_exp=8.0)
    emb_ref = torch.from_numpy(oenc.sinusoidal_encode(X.numpy(), 10, 0.0, 8.0))
    y_ref = ref.forward(emb_ref)
    with torch.no_grad():
        y_a = net.forward(enc(X.cuda()))
        y_b = net.forward_sinusoidal(X.cuda().float(), enc.freq_bands("cuda"))
    assert rel_max(y_a.cpu(), y_ref) < 1e-2
    assert rel_max(y_b.cpu(), y_ref) < 1e-2


def test_two_forwards_before_backward_and_chunked_run_model():
    """the saved activations live in ONE workspace per model.  A second saving forward before the first
    one's backward (chunked callers: run_model's netchunk loop, batchify_rays, render_rays_eval with network_fine=None)
    must not corrupt the first one's gradient: its backward recomputes what was overwritten."""
    from nerf_meets_mlx_b200.models import NeRF
    from nerf_meets_mlx_b200.models.NeRF import run_model
    from nerf_meets_mlx_b200.models import embedding
    cfg = CFGS[0]
    torch.manual_seed(0)
    net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, **cfg)
    cin = cfg["channel_input"] + cfg["channel_input_views"]
    xa, xb = torch.randn(300, cin, device="cuda").clamp(-1, 1), torch.randn(500, cin, device="cuda").clamp(-1, 1)
    ga, gb = torch.randn(300, 4, device="cuda"), torch.randn(500, 4, device="cuda")

    def grad_of(x, g):
        net.flat.grad = None
        (net.forward(x) * g).sum().backward()
        return net.flat.grad.clone()
    ref = grad_of(xa, ga) + grad_of(xb, gb)
    n0 = net.recomputed_backwards
    net.flat.grad = None
    ya, yb = net.forward(xa), net.forward(xb)     # the second forward overwrites the first one's activations
    ((ya * ga).sum() + (yb * gb).sum()).backward()
    assert net.recomputed_backwards == n0 + 1
    assert float((net.flat.grad - ref).norm() / ref.norm()) < 1e-5
    # chunked run_model (netchunk < points: several model.forward calls inside one autograd graph)
    e_pos, _ = embedding.get_embedder(10)
    e_dir, _ = embedding.get_embedder(4)
    pos = torch.randn(40, 16, 3, device="cuda")
    dirs = torch.nn.functional.normalize(torch.randn(40, 3, device="cuda"), dim=-1)
    g = torch.randn(40, 16, 4, device="cuda")
    net.flat.grad = None
    (run_model(pos, e_pos, dirs, e_dir, net, netchunk=1 << 20) * g).sum().backward()
    one = net.flat.grad.clone()
    net.flat.grad = None
    n0 = net.recomputed_backwards
    (run_model(pos, e_pos, dirs, e_dir, net, netchunk=256) * g).sum().backward()   # 3 chunks
    assert net.recomputed_backwards >= n0 + 2
    assert float((net.flat.grad - one).norm() / one.norm()) < 1e-5
    # no-view-dir net (use_viewdirs=False configuration of the reference)
    net2 = NeRF(device="cuda", **CFGS[2])
    x1, x2 = torch.randn(200, CFGS[2]["channel_input"], device="cuda"), torch.randn(333, CFGS[2]["channel_input"], device="cuda")
    net2.flat.grad = None
    (net2.forward(x1).sum() * 1.0).backward()
    r1 = net2.flat.grad.clone()
    net2.flat.grad = None
    (net2.forward(x2).sum() * 1.0).backward()
    r2 = net2.flat.grad.clone()
    net2.flat.grad = None
    (net2.forward(x1).sum() + net2.forward(x2).sum()).backward()
    assert float((net2.flat.grad - (r1 + r2)).norm() / (r1 + r2).norm()) < 1e-5
