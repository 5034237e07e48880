"""GPU parity of the fused MLP chains (nmx_chain.cu) beyond one tile per CTA, the ReLU sign-bit store, and
size-independent properties at the BASELINE.json sizes (8192 rays x 192 samples = 1.57 M points)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import models as omodels

pytestmark = pytest.mark.gpu

KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)


def _pair(seed=5):
    from nerf_meets_mlx_b200.models import NeRF
    ref = omodels.NeRF(seed=seed, **KW)
    net = NeRF(device="cuda", **KW)
    net.load_reference_parameters(ref.params)
    return ref, net


def _rel_norm(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_multi_tile_per_cta_forward_backward_vs_emulated_reference():
    """38 605 points = 302 tiles: every CTA walks 2-3 tiles and the last tile is ragged (77 valid rows)."""
    from test_mlp_gpu import emulated_forward
    P = 148 * 128 * 2 + 128 * 5 + 77
    torch.manual_seed(1)
    ref, net = _pair()
    x = torch.randn(P, 90).clamp(-1, 1)
    ref.requires_grad_(True)
    y = net.forward(x.cuda())
    y_emu = emulated_forward(ref, x)
    assert float((y.detach().cpu() - y_emu.detach()).abs().max() / y_emu.detach().abs().max()) < 2e-3
    # inference (no saved activations) and training forward give the same numbers
    with torch.no_grad():
        y_inf = net.forward(x.cuda())
    assert torch.equal(y_inf, y.detach())
    g_out = torch.randn(P, 4)
    names = list(ref.params.keys())
    g_emu = torch.autograd.grad((y_emu * g_out).sum(), [ref.params[n] for n in names])
    (y * g_out.cuda()).sum().backward()
    got = net.split_flat(net.flat.grad)
    for n, ge in zip(names, g_emu):
        assert _rel_norm(got[n].cpu(), ge) < 1e-2, n


def test_relu_sign_bits_match_saved_activations():
    """The training forward's 1-bit-per-activation store (what the backward chain masks with) == (saved h > 0),
    bit e / 16+e of word w = columns 32w+2e / 32w+2e+1."""
    from nerf_meets_mlx_b200 import _lib_loader as L
    P = 148 * 128 + 300
    torch.manual_seed(2)
    _, net = _pair()
    net.flat.requires_grad_(True)
    net.forward(torch.randn(P, 90, device="cuda").clamp(-1, 1))
    torch.cuda.synchronize()
    out = (ctypes.c_int64 * 12)()
    L.call("nmx_mlp_debug_layout", net._plan, out, L.i32(12))
    base, _, h0, hs, _, hd, _, _, _, bits, cap, _ = list(out)
    ws = net._ws
    assert bits > 0
    b = ws[base + bits: base + bits + 9 * cap * 32].view(torch.int32).view(9, cap, 8)[:, :P].cpu().numpy().astype(np.uint32)
    for slot in range(9):
        cols = 256 if slot < 8 else 128
        off = h0 + slot * hs if slot < 8 else hd
        h = ws[base + off: base + off + P * cols * 2].view(torch.bfloat16).view(P, cols).float().cpu().numpy()
        got = np.zeros((P, cols), dtype=bool)
        for w in range(cols // 32):
            for e in range(16):
                got[:, 32 * w + 2 * e] = (b[slot, :, w] >> e) & 1
                got[:, 32 * w + 2 * e + 1] = (b[slot, :, w] >> (16 + e)) & 1
        assert np.array_equal(got, h > 0), f"slot {slot}"


def test_full_size_properties():
    """BASELINE size (8192 x 192): training and inference forwards agree (2e-3: different kernels), the backward is linear in d_out
    (bwd(2 d) == 2 bwd(d) exactly: powers of two commute with every rounding), and repeated runs are deterministic
    up to the fp32 atomic order of the weight-gradient reduction."""
    from nerf_meets_mlx_b200.models import NeRF
    B, n = 8192, 192
    torch.manual_seed(3)
    net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, **KW)
    o = torch.randn(B, 3, device="cuda")
    d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda"), dim=-1)
    rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
    z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
    net.reserve(B * n, training=True)
    raw_t = net._fwd_raw(1, rays, z, None, B, n, save=True).clone()
    raw_i = net._fwd_raw(1, rays, z, None, B, n, save=False)
    assert torch.isfinite(raw_t).all()
    # inference runs on CTA pairs (nmx_chain2.cu): same arithmetic except that the per-ray view-dir term of the dir layer
    # is summed separately (fp32) instead of inside the tensor-core accumulation
    assert float((raw_t - raw_i).abs().max() / raw_t.abs().max()) < 2e-3
    net._fwd_raw(1, rays, z, None, B, n, save=True)
    d_raw = torch.randn(B * n, 4, device="cuda") * 1e-3
    g1 = net._bwd_raw(d_raw, B * n).clone()
    g2 = net._bwd_raw((2.0 * d_raw).contiguous(), B * n).clone()
    g1b = net._bwd_raw(d_raw, B * n).clone()
    assert torch.isfinite(g1).all() and float(g1.abs().max()) > 0
    assert _rel_norm(g2, 2.0 * g1) < 1e-5
    assert _rel_norm(g1b, g1) < 1e-5


def test_cta_pair_two_tile_chain_matches_one_tile_chain():
    """nmx_chain2.cu (tcgen05 cta_group::2, two pair tiles in ping-pong, per-ray view-dir term in the dir layer's
    epilogue) gives the one-tile chain's raw outputs up to the re-ordered view-dir sum: 2 / 9 / 37 / 75 / 2368 / 8192 rays (1, 3,
    10, 19, ... pair tiles: odd counts leave a cluster's second slot empty; ragged last tiles) (the script compares the inference forward = chain2 with the training forward = one-tile chain)."""
    import os
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.pop("NMX_CHAIN2_DBG", None)
    env.pop("NMX_DISABLE_CHAIN2", None)
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "chain2_check.py")], cwd=root, env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    errs = [float(x) for x in re.findall(r"rel max err ([0-9.e+-]+)", r.stdout)]
    assert len(errs) == 6 and all(e < 2e-3 for e in errs), r.stdout
    assert r.stdout.count("finite=True") == 6


@pytest.mark.parametrize("B,n", [(525, 8), (545, 8), (33, 128), (1000, 8), (64, 64), (700, 24)])
def test_pair_training_chains_match_one_tile_chains(B, n):
    """The CTA-pair training forward / backward (nmx_chain2t.cu, taken for rays-path passes of >= 4096 points with >= 8
    samples per ray) against the one-tile training forward (nmx_chain.cu) on the SAME points: the one-tile forward is
    selected in-process by presenting every ray several times with < 8 of its samples each (below the pair forward's
    limit); the pair backward then runs on the one-tile forward's row-major sign bits, so both sign-bit layouts are
    exercised and must give the same gradient.
    Sizes cover an odd number of 256-point pair tiles (one cluster runs a single tile from the start), a ragged last
    tile and whole multiples."""
    from nerf_meets_mlx_b200.models import NeRF
    torch.manual_seed(B * 1000 + n)
    net = NeRF(device="cuda", n_freqs_pos=10, n_freqs_dir=4, seed=11, **KW)
    with torch.no_grad():
        net.alpha_linear.bias.fill_(0.3)
        net.mark_params_updated()
    o = torch.randn(B, 3, device="cuda") * 0.3 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
    d = torch.nn.functional.normalize(torch.randn(B, 3, device="cuda") - torch.tensor([0.0, 0.0, 2.0], device="cuda"), dim=-1)
    rays = torch.cat([o, d, torch.full((B, 1), 2.0, device="cuda"), torch.full((B, 1), 6.0, device="cuda"), d], -1).contiguous()
    z = torch.sort(torch.rand(B, n, device="cuda") * 4 + 2, -1).values.contiguous()
    P = B * n
    d_out = torch.randn(P, 4, device="cuda") * 1e-3
    net.reserve(P, training=True)
    # pair kernels
    raw_p = net._fwd_raw(1, rays, z, None, B, n, save=True).clone()
    g_p = net._bwd_raw(d_out, P).clone()
    # one-tile forward: rep * B "rays" of k < 8 samples each (same points in the same order)
    k = max(dv for dv in range(1, 8) if n % dv == 0)
    rep = n // k
    rays1 = rays.repeat_interleave(rep, dim=0).contiguous()
    z1 = z.reshape(B * rep, k).contiguous()
    raw_1 = net._fwd_raw(1, rays1, z1, None, B * rep, k, save=True).clone()
    g_1 = net._bwd_raw(d_out, P).clone()
    assert torch.isfinite(raw_p).all() and torch.isfinite(g_p).all()
    assert float((raw_p - raw_1).norm() / raw_1.norm()) < 1e-4
    assert float((g_p - g_1).norm() / g_1.norm()) < 1e-4
