"""Parity of the CUDA kernels (through the C ABI) against the CPU oracle and the golden vectors produced by the
reference's own code.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import sampling as osamp, encoding as oenc, rendering as orend  # noqa: E402


def dev(x, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to("cuda").to(dtype)


@pytest.fixture(scope="module")
def ops():
    from nerf_meets_mlx_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------ K1 sampling
def test_sample_z_bit_exact(ops, golden):
    g = golden("sample_z")
    near, far = dev(g["near"]), dev(g["far"])
    for n in (2, 64, 192):
        np.testing.assert_array_equal(ops.sample_z(near, far, n).cpu().numpy(), g[f"uniform_{n}"])
        np.testing.assert_array_equal(ops.sample_z(near, far, n, lindisp=True).cpu().numpy(), g[f"lindisp_{n}"])


def test_add_noise_z(ops):
    rng = np.random.default_rng(0)
    z = np.sort(rng.uniform(2, 6, size=(37, 64)).astype(np.float32), -1)
    t = rng.random(size=(37, 64), dtype=np.float32)
    for s in (1.0, 0.5):
        np.testing.assert_array_equal(ops.add_noise_z(dev(z), dev(t), s).cpu().numpy(), osamp.add_noise_z(z, t, s))
    np.testing.assert_array_equal(ops.add_noise_z(dev(z), dev(t), 0.0).cpu().numpy(), z)


def test_ray_points(ops):
    rng = np.random.default_rng(1)
    rays = rng.standard_normal((19, 11)).astype(np.float32)
    z = rng.uniform(2, 6, size=(19, 7)).astype(np.float32)
    ref = rays[:, None, 0:3] + z[:, :, None] * rays[:, None, 3:6]
    np.testing.assert_array_equal(ops.ray_points(dev(rays), dev(z)).cpu().numpy(), ref.astype(np.float32))


# ------------------------------------------------------------------ K2a / K2b positional encodings
def test_pe_embedder_golden(ops, golden):
    g = golden("pe_embedder")
    out = ops.pe_embedder(dev(g["pos"].reshape(-1, 3)), 10).cpu().numpy()
    np.testing.assert_allclose(out, g["pe_pos"], rtol=0, atol=2e-6)
    out = ops.pe_embedder(dev(g["dirs"]), 4).cpu().numpy()
    np.testing.assert_allclose(out, g["pe_dir"], rtol=0, atol=2e-6)
    out = ops.pe_embedder(dev(g["xy"]), 6, include_input=False).cpu().numpy()
    np.testing.assert_allclose(out, g["pe_xy"], rtol=0, atol=2e-6)


def test_pe_embedder_explicit_bands_golden(ops, golden):
    """The mirrored Embedder class with max_freq_log2 != num_freqs - 1: explicit bands through nmx_pe_embedder_bands_fwd."""
    from nerf_meets_mlx_b200.models.embedding import Embedder
    g = golden("pe_embedder_bands")
    for tag, x in (("a", "x3"), ("b", "x3"), ("c", "x2"), ("d", "x3")):
        e = Embedder(include_input=bool(int(g[f"{tag}_inc"])), input_dims=g[x].shape[-1], max_freq_log2=float(g[f"{tag}_max"]),
                     num_freqs=int(g[f"{tag}_n"]), log_sampling=True, periodic_funcs=["sin", "cos"])
        assert e.out_dim == int(g[f"{tag}_dim"])
        out = e.embed(dev(g[x])).cpu().numpy()
        np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=0, atol=2e-6)


def test_pe_sinusoidal_golden(ops, golden):
    g = golden("pe_sinusoidal")
    bands = oenc.sinusoidal_freq_bands(10, 0.0, 8.0)
    out = ops.pe_sinusoidal(dev(g["X"].astype(np.float32)), dev(bands)).cpu().numpy()
    # arguments reach 2.5e2*2.56e2 = 6.5e4: fp32 sin of the SAME fp32 argument, full-range reduction
    np.testing.assert_allclose(out, g["enc"], rtol=0, atol=3e-6)
    b3 = oenc.sinusoidal_freq_bands(4)
    out = ops.pe_sinusoidal(dev(g["x3"]), dev(b3), include_input=True).cpu().numpy()
    np.testing.assert_allclose(out, g["enc3"], rtol=0, atol=2e-6)


# ------------------------------------------------------------------ K2c hash grid
def test_hash_bit_exact(ops):
    rng = np.random.default_rng(2)
    c = rng.integers(-5000, 5000, size=(4096, 3)).astype(np.int32)
    for T in (14, 19, 24):
        got = ops.hashgrid_hash(dev(c, torch.int32), T).cpu().numpy().astype(np.int64)
        np.testing.assert_array_equal(got, oenc.hashgrid_hash(c, T))


def test_hashgrid_golden_reference_lines(ops, golden):
    """Kernel == what the reference's own lines produce (golden/hashgrid.npz, see test_oracle_golden.py)."""
    from nerf_meets_mlx_b200.encoding.multi_hash import MultiHashEncoding
    g = golden("hashgrid")
    c = g["hash_coords"]
    c32 = c.astype(np.int32)  # every golden coordinate fits int32
    assert np.array_equal(c32.astype(np.int64), c)
    for T in (10, 14, 19, 24):
        got = ops.hashgrid_hash(dev(c32, torch.int32), T).cpu().numpy().astype(np.int64)
        np.testing.assert_array_equal(got, g[f"hash_T{T}"])
    for tag in "abcd":
        L_, nmin, nmax, F_, T = (int(v) for v in g[f"{tag}_cfg"])
        enc = MultiHashEncoding(3, L_, nmin, nmax, F_, T, device="cuda")
        np.testing.assert_array_equal(enc.scaled_res.cpu().numpy(), g[f"{tag}_scaled_res"])
        assert enc.get_out_dim() == int(g[f"{tag}_out_dim"]) and enc.hash_table_size == int(g[f"{tag}_table_size"])
    for tag in "ab":
        L_, nmin, nmax, F_, T = (int(v) for v in g[f"call_{tag}_cfg"])
        tables = np.random.default_rng(900 + L_).uniform(-1.0, 1.0, size=(L_, 2 ** T, F_)).astype(np.float32)
        np.testing.assert_array_equal(tables[:, :64], g[f"call_{tag}_tables_sample"])
        out, idx = ops.hashgrid_fwd(dev(g[f"call_{tag}_x"]), dev(tables), dev(g[f"{tag}_scaled_res"]), T, return_idx=True)
        np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), g[f"call_{tag}_idx"])
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"call_{tag}_out"])
        enc = MultiHashEncoding(3, L_, nmin, nmax, F_, T, device="cuda")
        with torch.no_grad():
            enc.hash_table.copy_(dev(tables))
            np.testing.assert_array_equal(enc(dev(g[f"call_{tag}_x"])).cpu().numpy(), g[f"call_{tag}_out"])


@pytest.mark.parametrize("L,F,T", [(16, 2, 19), (4, 4, 12), (8, 1, 10)])
def test_hashgrid_fwd_bwd(ops, L, F, T):
    rng = np.random.default_rng(3)
    P = 5000
    x = rng.random(size=(P, 3), dtype=np.float32)
    x[:7] = np.array([0.0, 0.25, 0.5], np.float32)  # exact grid hits: floor == ceil, offset 0
    x[7] = -0.3
    tables = rng.uniform(-1e-4, 1e-4, size=(L, 1 << T, F)).astype(np.float32)
    res = oenc.hashgrid_scaled_res(L, 16, 2048)
    out, idx = ops.hashgrid_fwd(dev(x), dev(tables), dev(res), T, return_idx=True)
    ref_idx, _ = oenc.hashgrid_corner_indices(x, res, T)
    np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)  # cell indices: bit-exact
    ref = oenc.hashgrid_encode(x, tables, res, T)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)  # same fp32 op order, no FMA contraction
    d_out = rng.standard_normal(size=(P, L * F)).astype(np.float32)
    g = ops.hashgrid_bwd(dev(x), dev(res), dev(d_out), L, F, T).cpu().numpy()
    gref = oenc.hashgrid_backward(x, d_out, L, F, res, T)
    np.testing.assert_allclose(g, gref, rtol=2e-4, atol=2e-5)


# ------------------------------------------------------------------ K4 compositing
def test_composite_golden(ops, golden):
    g = golden("raw2outputs")
    for tag in "abc":
        for wb in (0, 1):
            outs = ops.composite_fwd(dev(g[f"{tag}_raw"]), dev(g[f"{tag}_z"]), dev(g[f"{tag}_d"]), white_bkgd=bool(wb))
            for name, o in zip(("rgb", "disp", "acc", "weights", "depth"), outs):
                ref = g[f"{tag}_{wb}_{name}"]
                assert tuple(o.shape) == tuple(ref.shape), name
                # north_star: 1e-5 relative for fp32 compositing (atol covers cancellation in sums)
                np.testing.assert_allclose(o.cpu().numpy(), ref, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(ref).max()), err_msg=name)


@pytest.mark.parametrize("n", [1, 5, 64, 100, 192, 256])
def test_composite_fwd_bwd_vs_oracle(ops, n):
    rng = np.random.default_rng(n)
    B = 257
    raw = rng.standard_normal(size=(B, n, 4)).astype(np.float32)
    raw[..., 3] = raw[..., 3] * 4.0 + 1.0
    z = np.sort(rng.uniform(2, 6, size=(B, n)).astype(np.float32), -1)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    for wb in (False, True):
        rt = torch.from_numpy(raw).requires_grad_(True)
        outs_ref = orend.raw2outputs(rt, z, d, 0, wb)
        outs = ops.composite_fwd(dev(raw), dev(z), dev(d), white_bkgd=wb)
        for o, r, name in zip(outs, outs_ref, ("rgb", "disp", "acc", "weights", "depth")):
            r = r.detach().numpy()
            if name == "disp":  # 1/max(1e-10, depth/acc): ill-conditioned when acc ~ 0
                continue
            np.testing.assert_allclose(o.cpu().numpy(), r, rtol=1e-5, atol=2e-5 * max(1.0, np.abs(r).max()), err_msg=name)
        g_rgb = rng.standard_normal(size=(B, 3)).astype(np.float32)
        g_acc = rng.standard_normal(size=(B, 1)).astype(np.float32)
        g_depth = rng.standard_normal(size=(B, 1)).astype(np.float32)
        g_w = rng.standard_normal(size=(B, n, 1)).astype(np.float32)
        loss = (outs_ref[0] * torch.from_numpy(g_rgb)).sum() + (outs_ref[2] * torch.from_numpy(g_acc)).sum() \
            + (outs_ref[4] * torch.from_numpy(g_depth)).sum() + (outs_ref[3] * torch.from_numpy(g_w)).sum()
        (g_ref,) = torch.autograd.grad(loss, rt)
        got = ops.composite_bwd(dev(raw), dev(z), dev(d), dev(g_rgb), d_acc=dev(g_acc), d_depth=dev(g_depth),
                                d_weights=dev(g_w), white_bkgd=wb).cpu().numpy()
        g_ref = g_ref.numpy()
        scale = np.abs(g_ref).max()
        np.testing.assert_allclose(got, g_ref, rtol=2e-4, atol=2e-5 * scale)


def test_composite_bwd_rgb_only_and_disp(ops):
    rng = np.random.default_rng(77)
    B, n = 64, 64
    raw = rng.standard_normal(size=(B, n, 4)).astype(np.float32)
    raw[..., 3] = np.abs(raw[..., 3]) * 2
    z = np.sort(rng.uniform(2, 6, size=(B, n)).astype(np.float32), -1)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    rt = torch.from_numpy(raw).requires_grad_(True)
    rgb, disp, acc, w, depth = orend.raw2outputs(rt, z, d, 0, True)
    g_rgb = rng.standard_normal(size=(B, 3)).astype(np.float32)
    g_disp = rng.standard_normal(size=(B, 1)).astype(np.float32)
    (g_ref,) = torch.autograd.grad((rgb * torch.from_numpy(g_rgb)).sum() + (disp * torch.from_numpy(g_disp)).sum(), rt)
    got = ops.composite_bwd(dev(raw), dev(z), dev(d), dev(g_rgb), d_disp=dev(g_disp), white_bkgd=True).cpu().numpy()
    np.testing.assert_allclose(got, g_ref.numpy(), rtol=5e-4, atol=5e-5 * np.abs(g_ref.numpy()).max())


def test_composite_empty(ops):
    e = lambda *s: torch.empty(s, device="cuda")
    outs = ops.composite_fwd(e(0, 64, 4), e(0, 64), e(0, 3))
    assert outs[0].shape == (0, 3)


# ------------------------------------------------------------------ K5 resampling
def test_sample_pdf_indices_bit_exact_given_reference_cdf(ops, golden):
    g = golden("sample_pdf")
    for tag in "abc":
        r = ops.sample_pdf(dev(g[f"{tag}_z"]), None, dev(g[f"{tag}_u"]), cdf=dev(g[f"{tag}_cdf"]), want_inds=True)
        np.testing.assert_array_equal(r["inds"].cpu().numpy().astype(np.int64), g[f"{tag}_inds"])
        np.testing.assert_array_equal(r["z_imp"].cpu().numpy(), g[f"{tag}_out"])
        ref_sorted = np.sort(np.concatenate([g[f"{tag}_z"], g[f"{tag}_out"]], -1), -1)
        np.testing.assert_array_equal(r["z_merged"].cpu().numpy(), ref_sorted)


@pytest.mark.parametrize("B,n,N", [(1, 64, 128), (513, 64, 128), (33, 16, 40), (7, 192, 64), (3, 2, 5), (19, 64, 256),
                                   (5, 64, 300), (9, 33, 97), (40000, 64, 128)])
def test_sample_pdf_vs_oracle_bit_exact(ops, B, n, N):
    rng = np.random.default_rng(B + n)
    z = np.sort(rng.uniform(2, 6, size=(B, n)).astype(np.float32), -1)
    w = (rng.random(size=(B, n, 1)) ** 4).astype(np.float32)
    if B > 2:
        w[0] = 0
        w[1] = 0
        w[1, n // 2] = 5.0
    u = rng.random(size=(B, N), dtype=np.float32)
    u[0, 0] = 0.0
    if N > 1:
        u[0, 1] = np.float32(1.0) - np.float32(2 ** -24)
    r = ops.sample_pdf(dev(z), dev(w), dev(u), want_inds=True, want_cdf=True)
    cdf = osamp.build_cdf(w)
    np.testing.assert_array_equal(r["cdf"].cpu().numpy(), cdf)
    out, inds = osamp.sample_pdf(z, w, u, return_inds=True)
    np.testing.assert_array_equal(r["inds"].cpu().numpy().astype(np.int64), inds)
    np.testing.assert_array_equal(r["z_imp"].cpu().numpy(), out)
    np.testing.assert_array_equal(r["z_merged"].cpu().numpy(), osamp.merge_sorted(z, out))
    m = r["z_merged"].cpu().numpy()
    assert np.all(np.diff(m, axis=-1) >= 0)


def test_sample_pdf_merge_unsorted_coarse_depths_and_ties(ops):
    """sort(concat) must hold for ANY coarse row (rank-counting path) and with heavy ties (degenerate CDF bins collapse
    many samples onto the same mid-point; coarse duplicates)."""
    rng = np.random.default_rng(77)
    B, n, N = 65, 64, 128
    z = rng.uniform(2, 6, size=(B, n)).astype(np.float32)           # not sorted
    z[1::2] = np.sort(z[1::2], -1)                                   # every other row sorted: both paths in one launch
    z[3, 10:20] = z[3, 10]                                           # duplicated coarse depths
    w = np.zeros((B, n, 1), np.float32)
    w[:, 5] = 1.0                                                    # one-hot weights: most samples share a bin
    u = rng.random(size=(B, N), dtype=np.float32)
    u[:, ::3] = u[:, :1]                                             # repeated draws -> exactly equal samples
    r = ops.sample_pdf(dev(z), dev(w), dev(u))
    imp = r["z_imp"].cpu().numpy()
    np.testing.assert_array_equal(r["z_merged"].cpu().numpy(), np.sort(np.concatenate([z, imp], -1), -1))


def test_sort_merge(ops):
    rng = np.random.default_rng(5)
    a = rng.standard_normal(size=(77, 64)).astype(np.float32)
    b = rng.standard_normal(size=(77, 128)).astype(np.float32)
    b[:, :5] = a[:, :5]  # ties
    got = ops.sort_merge_z(dev(a), dev(b)).cpu().numpy()
    np.testing.assert_array_equal(got, np.sort(np.concatenate([a, b], -1), -1))


# ------------------------------------------------------------------ K6 loss / optimiser
def test_mse_adam(ops):
    rng = np.random.default_rng(6)
    p = rng.standard_normal(size=(8192, 3)).astype(np.float32)
    t = rng.random(size=(8192, 3)).astype(np.float32)
    loss, d = ops.mse_fwd_bwd(dev(p), dev(t))
    np.testing.assert_allclose(loss.item(), np.mean((p - t) ** 2), rtol=1e-5)
    np.testing.assert_allclose(d.cpu().numpy(), 2 * (p - t) / p.size, rtol=1e-6, atol=1e-10)
    n = 100003
    w = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    m = rng.standard_normal(n).astype(np.float32) * 0.1
    v = rng.random(n).astype(np.float32) * 0.1
    W, G, M, V = dev(w), dev(g), dev(m), dev(v)
    ops.adam_step(W, G, M, V, 5e-4)
    m2 = 0.9 * m + 0.1 * g
    v2 = 0.999 * v + 0.001 * g * g
    np.testing.assert_allclose(M.cpu().numpy(), m2, rtol=1e-5, atol=2e-7)  # fma contraction vs numpy
    np.testing.assert_allclose(V.cpu().numpy(), v2, rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(W.cpu().numpy(), w - 5e-4 * m2 / (np.sqrt(v2) + 1e-8), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ ray generation, SH, metrics
def test_gen_rays_golden_bit_exact(ops, golden):
    """nmx_gen_rays against the reference's own get_rays output (bit-exact o, d) and the restated ray assembly."""
    g, gf = golden("rays"), golden("render_full")
    H, W = int(g["H"]), int(g["W"])
    from nerf_meets_mlx_b200.rendering import ray
    ro, rd = ray.get_rays(H, W, g["K"], g["c2w"][:3, :4])
    np.testing.assert_array_equal(ro.cpu().numpy(), g["rays_o"].astype(np.float32))
    np.testing.assert_array_equal(rd.cpu().numpy(), g["rays_d"].astype(np.float32))
    rays = ops.gen_rays(H, W, gf["K"], gf["c2w"], None, 2.0, 6.0, 11).cpu().numpy()
    ref = orend.build_rays(H, W, gf["K"], gf["c2w"][:3, :4], 2.0, 6.0, use_viewdirs=True)
    np.testing.assert_array_equal(rays[:, :8], ref[:, :8])
    np.testing.assert_allclose(rays[:, 8:], ref[:, 8:], rtol=0, atol=1.2e-7)


def test_gen_rays_pixel_selection_and_targets(ops):
    """The training loop's ray selection (__test_nerf.py:208-236): random pixel ids of a 400x400 view, rays + targets."""
    rng = np.random.default_rng(3)
    H = W = 400
    focal = 0.5 * W / np.tan(0.5 * 0.6911112)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = orend.pose_spherical(77.0, -30.0, 4.0)
    img = rng.random(size=(H * W, 4)).astype(np.float32)
    pix = rng.choice(H * W, size=4096, replace=False).astype(np.int32)
    rays, tgt = ops.gen_rays(H, W, K, c2w, torch.from_numpy(pix).cuda(), 2.0, 6.0, 11, image=dev(img))
    ro, rd = orend.get_rays(H, W, K, c2w[:3, :4])
    sel_o = np.reshape(ro, (-1, 3))[pix].astype(np.float32)
    sel_d = np.reshape(rd, (-1, 3))[pix].astype(np.float32)
    rays = rays.cpu().numpy()
    np.testing.assert_array_equal(rays[:, 0:3], sel_o)
    np.testing.assert_array_equal(rays[:, 3:6], sel_d)
    np.testing.assert_array_equal(rays[:, 6:8], np.tile(np.array([[2.0, 6.0]], np.float32), (4096, 1)))
    vd = sel_d / np.sqrt(np.sum(sel_d * sel_d, -1, keepdims=True))
    np.testing.assert_allclose(rays[:, 8:11], vd, rtol=0, atol=1.2e-7)
    np.testing.assert_array_equal(tgt.cpu().numpy(), img[pix, :3])
    # empty batch and bad arguments
    assert ops.gen_rays(H, W, K, c2w, torch.zeros(0, dtype=torch.int32, device="cuda"), 2.0, 6.0, 11).shape == (0, 11)
    with pytest.raises(RuntimeError):
        ops.gen_rays(H, W, K, c2w, None, 2.0, 6.0, 7)


def test_sh_encoding_golden_bit_exact(ops, golden):
    from nerf_meets_mlx_b200.encoding import SphericalHarmonicsEncoding, IdentityEncoding
    g = golden("sh")
    d = dev(g["dirs"])
    for lv in range(5):
        enc = SphericalHarmonicsEncoding(3, lv)
        assert enc.get_out_dim() == int(g[f"sh_dim_{lv}"])
        np.testing.assert_array_equal(enc(d).cpu().numpy(), g[f"sh_{lv}"])
    with pytest.raises(AssertionError):
        SphericalHarmonicsEncoding(3, 5)
    big = torch.nn.functional.normalize(torch.randn(100_003, 3, device="cuda"), dim=-1)
    np.testing.assert_array_equal(ops.sh_encode(big, 4).cpu().numpy(), oenc.sh_encode(big.cpu().numpy(), 4))
    ident = IdentityEncoding(3)
    assert ident.get_out_dim() == 3 and ident(d) is d
    assert ops.sh_encode(big[:0], 4).shape == (0, 25)


def test_metric_mirror(golden):
    from nerf_meets_mlx_b200.ops import metric
    g = golden("metric")
    np.testing.assert_allclose(float(metric.MSE()(dev(g["pred"]), dev(g["gt"]))), float(g["mse"]), rtol=2e-6)
    np.testing.assert_allclose(float(metric.PSNR()(dev(g["pred"]), dev(g["gt"]))), float(g["psnr"]), rtol=2e-6)
    with pytest.raises(NotImplementedError):
        metric.SSIM()(dev(g["pred"]), dev(g["gt"]))


# ------------------------------------------------------------------ torch.library custom operators
def test_custom_ops_opcheck(ops):
    """torch.library.opcheck: schema, fake kernel vs real kernel (shapes / dtypes / strides), autograd registration and
    AOT dispatch of the custom operators the mirror API is built on."""
    from torch.library import opcheck
    from nerf_meets_mlx_b200.ops import library  # noqa: F401
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.rand(*s, device="cuda", generator=g)
    B, n, N = 33, 64, 128
    z = torch.sort(2.0 + 4.0 * r(B, n), dim=-1).values
    raw = (r(B, n, 4) - 0.3).requires_grad_(True)
    d = r(B, 3) - 0.5
    opcheck(torch.ops.nmx.composite_fwd.default, (raw, z, d, None, 0.0, True))
    opcheck(torch.ops.nmx.composite_bwd.default, (raw.detach(), z, d, r(B, 3)))
    w = torch.ops.nmx.composite_fwd(raw.detach(), z, d)[3]
    opcheck(torch.ops.nmx.sample_pdf.default, (z, w, r(B, N)))
    opcheck(torch.ops.nmx.sample_z.default, (2.0 + r(B), 6.0 + r(B), n, False))
    opcheck(torch.ops.nmx.pe_embedder.default, (r(50, 3), 10, True))
    opcheck(torch.ops.nmx.sh_encode.default, (r(50, 3), 4))
    tables = (r(4, 1 << 10, 2) - 0.5).requires_grad_(True)
    res = torch.tensor([16.0, 32.0, 64.0, 128.0], device="cuda")
    opcheck(torch.ops.nmx.hashgrid_fwd.default, (r(77, 3), tables, res, 10))
    opcheck(torch.ops.nmx.assemble_rays.default, (r(B, 3), r(B, 3) + 0.1, 2.0, 6.0))
    opcheck(torch.ops.nmx.mse_fwd_bwd.default, (r(B, 3), r(B, 3)))
    # autograd through the custom operators == the explicit backward kernels
    rgb, disp, acc, wts, depth = torch.ops.nmx.composite_fwd(raw, z, d, None, 0.0, False)
    (rgb.sum() + 0.5 * acc.sum()).backward()
    want = ops.composite_bwd(raw.detach(), z, d, torch.ones(B, 3, device="cuda"), d_acc=torch.full((B, 1), 0.5, device="cuda"))
    np.testing.assert_array_equal(raw.grad.cpu().numpy(), want.cpu().numpy())


def test_composite_loss_fused_equals_separate_kernels(ops):
    """nmx_composite_loss_fwd_bwd (compositing + MSE + its gradient in one kernel) == composite_fwd -> mse_fwd_bwd ->
    composite_bwd: rgb and weights bit-identical, d_raw and loss to fp32 contraction / summation order."""
    rng = np.random.default_rng(11)
    for B, n, wb in ((37, 64, True), (50, 192, False), (5, 7, True), (1024, 64, True)):
        raw = rng.standard_normal(size=(B, n, 4)).astype(np.float32)
        raw[..., 3] *= 2.0
        z = np.sort(rng.uniform(2, 6, size=(B, n)).astype(np.float32), -1)
        d = rng.standard_normal(size=(B, 3)).astype(np.float32)
        tgt = rng.random(size=(B, 3)).astype(np.float32)
        R, Z, D, T = dev(raw), dev(z), dev(d), dev(tgt)
        rgb, _, _, w, _ = ops.composite_fwd(R, Z, D, white_bkgd=wb)
        loss, d_rgb = ops.mse_fwd_bwd(rgb, T)
        d_raw = ops.composite_bwd(R, Z, D, d_rgb, white_bkgd=wb)
        loss2, d_raw2, w2, rgb2 = ops.composite_loss_fwd_bwd(R, Z, D, T, white_bkgd=wb, want_weights=True, want_rgb=True)
        assert torch.equal(rgb2, rgb) and torch.equal(w2, w)
        err = float((d_raw2 - d_raw).abs().max() / d_raw.abs().max())
        assert err < 1e-6, err
        assert abs(float(loss2) - float(loss)) <= 1e-5 * abs(float(loss))
