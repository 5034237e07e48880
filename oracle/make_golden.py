"""Generate tests/golden/*.npz by executing the reference's OWN, UNMODIFIED Python files from
/root/reference (TEST INFRASTRUCTURE; runs only in the build container, where /root/reference
exists -- the committed .npz files are what travels).

MLX 0.7.0 cannot be installed here, so `mlx.core` / `mlx.nn` / `mlx.optimizers` resolve to the
NumPy-backed stand-in in oracle/mlx_shim (see its docstring for the assumed MLX semantics).
`sample_from_inverse_cdf_torch` is pure torch and runs for real.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NMX_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def make_tiny_scene(scene, rng):
    """A 3-split Blender-format scene of 8x8 RGBA frames (deterministic), written once and committed."""
    import json
    from PIL import Image
    if os.path.exists(os.path.join(scene, "transforms_train.json")):
        return
    for split, n in (("train", 3), ("val", 2), ("test", 4)):
        os.makedirs(os.path.join(scene, split), exist_ok=True)
        frames = []
        for i in range(n):
            px = rng.integers(0, 256, size=(8, 8, 4), dtype=np.uint8)
            px[:2, :, 3] = 0      # fully transparent rows (white-background compositing path)
            px[2:4, :, 3] = 255
            Image.fromarray(px, "RGBA").save(os.path.join(scene, split, f"r_{i}.png"))
            th = float(rng.uniform(-180, 180))
            c, s_ = float(np.cos(np.deg2rad(th))), float(np.sin(np.deg2rad(th)))
            frames.append({"file_path": f"./{split}/r_{i}", "rotation": 0.1,
                           "transform_matrix": [[c, -s_, 0.0, 4.0 * s_], [s_, c, 0.0, -4.0 * c], [0.0, 0.0, 1.0, 0.5 * i],
                                                [0.0, 0.0, 0.0, 1.0]]})
        with open(os.path.join(scene, f"transforms_{split}.json"), "w") as fp:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, fp, indent=1)


def main():
    sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
    sys.path.insert(0, REF)
    import mlx.core as mx
    import mlx.nn as mnn
    from mlx_nerf import sampling
    from mlx_nerf.sampling import uniform, linear_disparity
    from mlx_nerf.models import NeRF as RN, embedding
    from mlx_nerf.encoding.sinusoidal import SinusoidalEncoding
    from mlx_nerf.rendering import render, ray
    from mlx_nerf.ops import pose

    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- sampling: sample_z (uniform.py:7-18, linear_disparity.py:8-19)
    near = rng.uniform(0.5, 3.0, size=(7, 1)).astype(np.float32)
    far = (near + rng.uniform(1.0, 5.0, size=(7, 1))).astype(np.float32)
    g = {"near": near, "far": far}
    for n in (2, 64, 192):
        g[f"uniform_{n}"] = uniform.sample_z(near, far, n)
        with np.errstate(divide="ignore", invalid="ignore"):
            g[f"lindisp_{n}"] = linear_disparity.sample_z(near, far, n)
    np.savez_compressed(os.path.join(OUT, "sample_z.npz"), **g)

    # ---- sample_from_inverse_cdf_torch (sampling/__init__.py:101-178), real torch
    g = {}
    for tag, B, n, N in (("a", 33, 64, 128), ("b", 5, 16, 40), ("c", 9, 64, 128)):
        z = np.sort(rng.uniform(2.0, 6.0, size=(B, n)).astype(np.float32), axis=-1)
        if tag == "c":  # peaky weights, incl. exact zeros and a one-hot row
            w = (rng.random(size=(B, n, 1)) ** 8).astype(np.float32)
            w[0] = 0.0
            w[1] = 0.0
            w[1, 17] = 1.0
        else:
            w = rng.random(size=(B, n, 1)).astype(np.float32)
        torch.manual_seed(1234 + B)
        u = torch.rand([B, N])  # same draw the function makes first thing from the global RNG
        torch.manual_seed(1234 + B)
        out = sampling.sample_from_inverse_cdf_torch(torch.from_numpy(z), torch.from_numpy(w), N)
        # the CDF the reference built (re-derived with the same torch ops, :113-131)
        wt = torch.from_numpy(w)[..., 0] + 0.01
        ws = torch.sum(wt, dim=-1, keepdim=True)
        pdf = wt / ws
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, axis=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        inds = torch.searchsorted(cdf, u, side="right")
        g.update({f"{tag}_z": z, f"{tag}_w": w, f"{tag}_u": u.numpy(), f"{tag}_out": out.numpy(),
                  f"{tag}_cdf": cdf.numpy(), f"{tag}_inds": inds.numpy()})
    np.savez_compressed(os.path.join(OUT, "sample_pdf.npz"), **g)

    # ---- larger resampling golden: 2048 rays x 128 draws through the REAL reference function, weights shaped
    # like trained compositing weights (a few peaks over a small floor), plus the end-to-end index-mismatch count of the
    # canonical CDF (fp64 sum / prefix, oracle + kernels) against the reference's own torch-CPU CDF over > 10^7 draws
    # (torch-CPU `sum` has an ISA-dependent fp32 order: this container's AVX512 host; SURVEY 7 probed ~1.2e-6).
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import sampling as osamp
    rng_s = np.random.default_rng(20261020)

    def peaky(B, n):
        w = 0.002 * rng_s.random(size=(B, n, 1))
        for _ in range(3):
            c = rng_s.integers(0, n, size=B)
            w[np.arange(B), c, 0] += rng_s.random(size=B) ** 2
            w[np.arange(B), np.minimum(c + 1, n - 1), 0] += 0.5 * rng_s.random(size=B) ** 2
        return w.astype(np.float32)

    def ref_draw(z, w, seed, N):
        torch.manual_seed(seed)
        u = torch.rand([z.shape[0], N])
        torch.manual_seed(seed)
        out = sampling.sample_from_inverse_cdf_torch(torch.from_numpy(z), torch.from_numpy(w), N)
        wt = torch.from_numpy(w)[..., 0] + 0.01
        pdf = wt / torch.sum(wt, dim=-1, keepdim=True)
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, axis=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        inds = torch.searchsorted(cdf, u, side="right")
        return u.numpy(), out.numpy(), inds.numpy()

    B, n, N = 2048, 64, 128
    z = np.sort(rng_s.uniform(2.0, 6.0, size=(B, n)).astype(np.float32), axis=-1)
    w = peaky(B, n)
    u, out, inds = ref_draw(z, w, 777, N)
    o_out, o_inds = osamp.sample_pdf(z, w, u, return_inds=True)
    g = {"z": z, "w": w, "u": u, "out": out, "inds": inds.astype(np.uint8),
         "mismatch_2048": np.int64(np.sum(o_inds != inds))}
    tot = mism = 0
    for k in range(10):  # 10 x 8192 rays x 128 draws = 10.5 M draws
        zb = np.sort(rng_s.uniform(2.0, 6.0, size=(8192, n)).astype(np.float32), axis=-1)
        wb = peaky(8192, n) if k % 2 else rng_s.random(size=(8192, n, 1)).astype(np.float32)
        ub, _, ib = ref_draw(zb, wb, 1000 + k, N)
        _, ob = osamp.sample_pdf(zb, wb, ub, return_inds=True)
        tot += ib.size
        mism += int(np.sum(ob != ib))
    g["big_draws"], g["big_mismatch"] = np.int64(tot), np.int64(mism)
    print(f"sample_pdf: canonical-CDF vs reference-CDF index mismatches: {mism} of {tot} draws ({mism / tot:.2e}); "
          f"2048-ray golden: {int(g['mismatch_2048'])} of {B * N}")
    np.savez_compressed(os.path.join(OUT, "sample_pdf_large.npz"), **g)

    # ---- PE flavour A (models/embedding.py) and embed()
    pos = rng.uniform(-4.0, 4.0, size=(6, 5, 3)).astype(np.float32)
    dirs = rng.standard_normal(size=(6, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    e_pos, d_pos = embedding.get_embedder(10)
    e_dir, d_dir = embedding.get_embedder(4)
    e_2d, d_2d = embedding.get_embedder(6, n_input_dims=2)
    xy = rng.uniform(-1, 1, size=(11, 2)).astype(np.float32)
    np.savez_compressed(
        os.path.join(OUT, "pe_embedder.npz"), pos=pos, dirs=dirs,
        pe_pos=e_pos(pos.reshape(-1, 3)), pe_dir=e_dir(dirs), d_pos=d_pos, d_dir=d_dir,
        embed=embedding.embed(pos, e_pos, dirs, e_dir), xy=xy, pe_xy=e_2d(xy), d_2d=d_2d)

    # ---- PE flavour B (encoding/sinusoidal.py), integer pixel coords as the image demo feeds
    X = np.stack(np.meshgrid(np.arange(0, 256, 37), np.arange(0, 256, 41), indexing="ij"), -1).reshape(-1, 2)
    se = SinusoidalEncoding(2, 10, min_freq_exp=0.0, max_freq_exp=8.0, is_include_input=False)
    se3 = SinusoidalEncoding(3, 4, is_include_input=True)
    x3 = rng.uniform(-2, 2, size=(9, 3)).astype(np.float32)
    # MLX promotes int32 * float32 -> float32 (NumPy would go to float64), so the integer coordinates
    # are pre-cast to float32 before entering the reference code under the NumPy stand-in.
    np.savez_compressed(os.path.join(OUT, "pe_sinusoidal.npz"), X=X.astype(np.int32), enc=se(X.astype(np.float32)),
                        out_dim=se.get_out_dim(), x3=x3, enc3=se3(x3), out_dim3=se3.get_out_dim())

    # ---- NeRF.forward (models/NeRF.py:160-243): view-dir net, no-view net, image net (small widths)
    def dump_params(model, prefix, g):
        for i, l in enumerate(model.list_linears_pos):
            g[f"{prefix}list_linears_pos.{i}.weight"] = l.weight
            g[f"{prefix}list_linears_pos.{i}.bias"] = l.bias
        for name in ("feature_linear", "alpha_linear", "rgb_linear", "output_linear"):
            if hasattr(model, name):
                g[f"{prefix}{name}.weight"] = getattr(model, name).weight
                g[f"{prefix}{name}.bias"] = getattr(model, name).bias
        if hasattr(model, "list_linears_dir"):
            g[f"{prefix}list_linears_dir.0.weight"] = model.list_linears_dir[0].weight
            g[f"{prefix}list_linears_dir.0.bias"] = model.list_linears_dir[0].bias

    mnn.seed(7)
    g = {}
    net_v = RN.NeRF(n_layers=8, width_layers=64, channel_input=63, channel_input_views=27, channel_output=5,
                    list_skip_connection_layers=[4], is_use_view_directions=True)
    net_n = RN.NeRF(n_layers=8, width_layers=64, channel_input=63, channel_input_views=27, channel_output=5,
                    list_skip_connection_layers=[4], is_use_view_directions=False)
    net_i = RN.NeRF(n_layers=8, width_layers=32, channel_input=40, channel_input_views=0, channel_output=3,
                    is_use_view_directions=False)
    xin = rng.standard_normal(size=(19, 90)).astype(np.float32)
    xim = rng.standard_normal(size=(13, 40)).astype(np.float32)
    dump_params(net_v, "v/", g)
    dump_params(net_n, "n/", g)
    dump_params(net_i, "i/", g)
    g.update(x=xin, y_v=net_v.forward(xin), y_n=net_n.forward(xin[:, :63]), x_img=xim, y_i=net_i.forward(xim))
    np.savez_compressed(os.path.join(OUT, "nerf_forward.npz"), **g)

    # ---- raw2outputs (rendering/render.py:20-96)
    g = {}
    for tag, B, n in (("a", 17, 64), ("b", 4, 192), ("c", 3, 5)):
        raw = rng.standard_normal(size=(B, n, 4)).astype(np.float32)
        raw[..., 3] *= 3.0  # negative and positive densities (T can exceed 1: reference quirk)
        z = np.sort(rng.uniform(2.0, 6.0, size=(B, n)).astype(np.float32), axis=-1)
        d = rng.standard_normal(size=(B, 3)).astype(np.float32)
        for wb in (False, True):
            outs = render.raw2outputs(raw, z, d, 0, wb)
            for name, o in zip(("rgb", "disp", "acc", "weights", "depth"), outs):
                g[f"{tag}_{int(wb)}_{name}"] = o
        g.update({f"{tag}_raw": raw, f"{tag}_z": z, f"{tag}_d": d})
    np.savez_compressed(os.path.join(OUT, "raw2outputs.npz"), **g)

    # ---- render_rays / render_rays_eval end to end (render.py:112-241) with the small view-dir net
    B, n, N = 12, 64, 128
    o = rng.uniform(-1, 1, size=(B, 3)).astype(np.float32) + np.array([0, 0, 4], np.float32)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    vd = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays = np.concatenate([o, d, 2.0 * np.ones((B, 1), np.float32), 6.0 * np.ones((B, 1), np.float32), vd], -1).astype(np.float32)
    qf = lambda inputs, viewdirs, model: RN.run_model(inputs, e_pos, viewdirs, e_dir, model, netchunk=256)
    mnn.seed(11)
    net_f = RN.NeRF(n_layers=8, width_layers=64, channel_input=63, channel_input_views=27, channel_output=5,
                    list_skip_connection_layers=[4], is_use_view_directions=True)
    g = {"rays": rays}
    dump_params(net_v, "c/", g)
    dump_params(net_f, "f/", g)
    r1 = render.render_rays(rays, net_v, qf, n, retraw=True, white_bkgd=True)
    for k, v in r1.items():
        g[f"rr_{k}"] = v
    torch.manual_seed(99)
    u = torch.rand([B, N]).numpy()
    torch.manual_seed(99)
    r2 = render.render_rays_eval(rays, net_v, qf, n, white_bkgd=True, N_importance=N, network_fine=net_f)
    for k, v in r2.items():
        g[f"re_{k}"] = v
    g["u"] = u
    np.savez_compressed(os.path.join(OUT, "render_rays.npz"), **g)

    # ---- get_rays / pose_spherical (rendering/ray.py:7-35, ops/pose.py:7-58)
    H, W, focal = 6, 8, 9.5
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    c2w = np.asarray(pose.pose_spherical(30.0, -30.0, 4.0))
    ro, rd = ray.get_rays(H, W, K, c2w[:3, :4])
    np.savez_compressed(os.path.join(OUT, "rays.npz"), K=K, c2w=c2w, rays_o=np.array(ro), rays_d=np.array(rd), H=H, W=W)

    # ---- render() incl. its ray assembly (render.py:268-345: get_rays, viewdir normalisation, near/far columns)
    mnn.seed(13)
    net_w = RN.NeRF(n_layers=8, width_layers=128, channel_input=63, channel_input_views=27, channel_output=5,
                    list_skip_connection_layers=[4], is_use_view_directions=True)  # 128: narrowest view-dir net the CUDA path takes
    rk = dict(network_query_fn=qf, network_coarse=net_w, n_depth_samples=16, white_bkgd=True, retraw=True,
              render_rays_func=render.render_rays)
    r3 = render.render(H, W, K, chunk=20, c2w=c2w[:3, :4], ndc=False, near=2.0, far=6.0, use_viewdirs=True, **rk)
    g = {}
    dump_params(net_w, "w/", g)
    np.savez_compressed(os.path.join(OUT, "render_full.npz"), K=K, c2w=c2w, H=H, W=W, rgb=r3[0], disp=r3[1], acc=r3[2],
                        **{f"x_{k}": v for k, v in r3[3].items()}, **g)

    # ---- SphericalHarmonicsEncoding / IdentityEncoding (encoding/spherical_harmonics.py, identity.py)
    from mlx_nerf.encoding.spherical_harmonics import SphericalHarmonicsEncoding
    from mlx_nerf.encoding.identity import IdentityEncoding
    dsh = rng.standard_normal(size=(64, 3)).astype(np.float32)
    dsh /= np.linalg.norm(dsh, axis=-1, keepdims=True)
    dsh[0] = [0, 0, 1]
    dsh[1] = [1, 0, 0]
    dsh[2] = [0, -1, 0]
    g = {"dirs": dsh, "identity": IdentityEncoding(3)(dsh), "identity_dim": IdentityEncoding(3).get_out_dim()}
    for lv in range(5):
        enc = SphericalHarmonicsEncoding(3, lv)
        g[f"sh_{lv}"] = enc(dsh)
        g[f"sh_dim_{lv}"] = enc.get_out_dim()
    np.savez_compressed(os.path.join(OUT, "sh.npz"), **g)

    # ---- MultiHashEncoding (encoding/multi_hash.py:13-137): the class cannot run as committed, its
    # pieces can.  Everything below executes the reference's OWN source lines, read from the reference tree at
    # generation time:
    #   (1) the real constructor (:14-53) -> growing_factor, scaled_res (:32-40), hash_table_size (:43);
    #   (2) `hash` (:61-77) on int64 coordinate arrays (int32 * 2654435761 has no defined result; on int64 nothing
    #       overflows for |coord| < 2^31 and the low log2_T bits equal the uint32-wraparound result), incl. negatives;
    #   (3) `__call__` (:79-137) with ONLY its eight table-lookup lines (:112-119, which call a Python list) replaced by
    #       the canonical per-level lookup `tables[level, hash(int64(grid_k))]`: the scaling, ceil/floor, corner
    #       construction (:90-109) and the interpolation + flatten (:121-137) are the reference's text, exec'd verbatim.
    import inspect
    import textwrap
    from mlx_nerf.encoding import multi_hash as RMH
    g = {}
    rng_h = np.random.default_rng(20261019)  # own stream: the sections below keep their draws
    for tag, (L_, nmin, nmax, F_, log2T) in {"a": (16, 16, 2048, 2, 19), "b": (8, 16, 512, 4, 14), "c": (1, 4, 4, 1, 10),
                                             "d": (5, 3, 100, 2, 12)}.items():
        mnn.seed(100 + L_)
        enc = RMH.MultiHashEncoding(3, L_, nmin, nmax, F_, log2T)
        g[f"{tag}_cfg"] = np.array([L_, nmin, nmax, F_, log2T], np.int64)
        g[f"{tag}_growing_factor"] = np.asarray(enc.growing_factor, np.float32)
        g[f"{tag}_scaled_res"] = np.asarray(enc.scaled_res, np.float32)
        g[f"{tag}_table_size"] = np.int64(enc.hash_table_size)
        g[f"{tag}_out_dim"] = np.int64(enc.get_out_dim())
        assert len(enc.hash_table) == L_ and enc.hash_table[0].weight.shape == (2 ** log2T, F_)
    # (2) hash on int64 arrays [B, L, 3]
    coords = np.concatenate([
        rng_h.integers(0, 2049, size=(4000, 4, 3)),                   # the C4 range
        rng_h.integers(0, 2 ** 31 - 1, size=(500, 4, 3)),             # large non-negative
        rng_h.integers(-2 ** 31, 0, size=(500, 4, 3)),                # negative (unclamped inputs)
        np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], [[1, 1, 1], [2047, 2048, 2047], [-1, -1, -1], [5, -7, 9]]]),
    ]).astype(np.int64)
    g["hash_coords"] = coords
    for log2T in (10, 14, 19, 24):
        stub = types.SimpleNamespace(hash_table_size=2 ** log2T)
        g[f"hash_T{log2T}"] = np.asarray(RMH.MultiHashEncoding.hash(stub, coords.copy()), np.int64)
    # (3) __call__ with the lookup lines replaced
    src = inspect.getsource(RMH.MultiHashEncoding.__call__).splitlines()
    i_a = next(i for i, ln in enumerate(src) if "in_array_scaled = in_array[..., None, :]" in ln)
    i_b = next(i for i, ln in enumerate(src) if "hashed_0 = self.hash_table(self.hash(grid_0))" in ln)
    i_c = next(i for i, ln in enumerate(src) if "offset = in_array_scaled - in_sf" in ln)
    assert all(f"hashed_{k} = self.hash_table(self.hash(grid_{k}))" in src[i_b + k] for k in range(8))
    body = src[i_a:i_b] + [f"        hashed_{k} = lookup(self.hash(grid_{k}.astype('int64')))" for k in range(8)] + src[i_c:]
    fn_src = "def _call(self, in_array, lookup, mx):\n" + textwrap.dedent("\n".join(body)).replace("\n", "\n    ")
    fn_src = fn_src.replace("def _call(self, in_array, lookup, mx):\n", "def _call(self, in_array, lookup, mx):\n    ")
    ns = {}
    exec(compile(fn_src, "<multi_hash.__call__ :90-109,121-137>", "exec"), ns)
    for tag, (L_, nmin, nmax, F_, log2T, B_, lo, hi) in {"a": (16, 16, 2048, 2, 19, 257, 0.0, 1.0),
                                                          "b": (8, 16, 512, 4, 14, 64, -1.5, 2.5)}.items():
        mnn.seed(200 + L_)
        enc = RMH.MultiHashEncoding(3, L_, nmin, nmax, F_, log2T)
        # tables are large (a: 64 MiB): drawn from their own seeded stream so that the tests can redraw them
        tables = np.random.default_rng(900 + L_).uniform(-1.0, 1.0, size=(L_, 2 ** log2T, F_)).astype(np.float32)
        x = rng_h.uniform(lo, hi, size=(B_, 3)).astype(np.float32)
        x[0] = [0.0, 0.5, 1.0]          # exact-integer scaled coordinates: ceil == floor, offset 0
        x[1] = [0.25, 0.125, 0.75]
        lv = np.arange(L_)[None, :]
        seen = []

        def lookup(idx, tables=tables, lv=lv, seen=seen):
            seen.append(np.asarray(idx, np.int64))
            return tables[lv, np.asarray(idx, np.int64)]
        out = ns["_call"](enc, x, lookup, mx)
        g.update({f"call_{tag}_x": x, f"call_{tag}_out": np.asarray(out, np.float32),
                  f"call_{tag}_idx": np.stack(seen, -1), f"call_{tag}_cfg": np.array([L_, nmin, nmax, F_, log2T], np.int64)})
        g[f"call_{tag}_tables_sample"] = tables[:, :64].copy()
        g[f"call_{tag}_tables_rng"] = np.int64(900 + L_)
    np.savez_compressed(os.path.join(OUT, "hashgrid.npz"), **g)

    # ---- metrics (ops/metric.py:12-18)
    from mlx_nerf.ops import metric
    pa = rng.random(size=(9, 7, 3)).astype(np.float32)
    pb = rng.random(size=(9, 7, 3)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "metric.npz"), pred=pa, gt=pb, mse=np.asarray(metric.MSE()(pa, pb)),
                        psnr=np.asarray(metric.PSNR()(pa, pb)))

    # ---- Blender loader (dataset/dataloader.py:20-113) on the tiny committed scene tests/golden/blender_tiny.
    # imageio / matplotlib are not in this image: imageio.v2.imread is stood in by PIL (same uint8 array), pyplot by
    # an empty module (only validate_dataset, a plotting helper, touches it).
    from PIL import Image
    scene = os.path.join(OUT, "blender_tiny")
    make_tiny_scene(scene, np.random.default_rng(5))
    im_mod, im_v2, mpl, plt = (types.ModuleType(n) for n in ("imageio", "imageio.v2", "matplotlib", "matplotlib.pyplot"))
    im_v2.imread = lambda f: np.asarray(Image.open(f))
    im_mod.v2 = im_v2
    mpl.pyplot = plt
    sys.modules.update({"imageio": im_mod, "imageio.v2": im_v2, "matplotlib": mpl, "matplotlib.pyplot": plt})
    from mlx_nerf.dataset import dataloader
    imgs, poses, render_poses, hwf, i_split = dataloader.load_blender_data(scene, half_res=False, testskip=2)
    i_tr, i_va, i_te, nr, fr, im_w = dataloader.post_load_blender_data(i_split, imgs, True)
    _, _, _, _, _, im_b = dataloader.post_load_blender_data(i_split, imgs, False)
    np.savez_compressed(os.path.join(OUT, "blender_loader.npz"), imgs=imgs, poses=poses, render_poses=np.asarray(render_poses),
                        hwf=np.asarray(hwf, dtype=np.float64), i_train=i_tr, i_val=i_va, i_test=i_te, near=nr, far=fr,
                        images_white=im_w, images_black=im_b)
    print("golden vectors written to", OUT)


def embedder_bands_golden():
    """Embedder built directly with max_freq_log2 != num_freqs - 1 (models/embedding.py:23-71): bands
    linspace(0, max_freq_log2, N) ** 2.  Own seed and own file, so the vectors above do not move:
        python oracle/make_golden.py --embedder-bands"""
    sys.path.insert(0, os.path.join(HERE, "mlx_shim"))
    sys.path.insert(0, REF)
    import mlx.core as mx
    from mlx_nerf.models import embedding
    rng = np.random.default_rng(20261019)
    x3 = rng.uniform(-2.0, 2.0, size=(37, 3)).astype(np.float32)
    x2 = rng.uniform(-1.0, 1.0, size=(19, 2)).astype(np.float32)
    out = {"x3": x3, "x2": x2}
    for tag, x, kw in (("a", x3, dict(include_input=True, input_dims=3, max_freq_log2=5.0, num_freqs=4)),
                       ("b", x3, dict(include_input=False, input_dims=3, max_freq_log2=2.5, num_freqs=7)),
                       ("c", x2, dict(include_input=True, input_dims=2, max_freq_log2=9, num_freqs=6)),
                       ("d", x3, dict(include_input=True, input_dims=3, max_freq_log2=3.0, num_freqs=1))):
        e = embedding.Embedder(log_sampling=True, periodic_funcs=[mx.sin, mx.cos], **kw)
        out[f"{tag}_out"] = np.asarray(e.embed(x), dtype=np.float32)
        out[f"{tag}_dim"] = e.out_dim
        out[f"{tag}_max"] = float(kw["max_freq_log2"])
        out[f"{tag}_n"] = kw["num_freqs"]
        out[f"{tag}_inc"] = int(bool(kw["include_input"]))
    np.savez_compressed(os.path.join(OUT, "pe_embedder_bands.npz"), **out)
    print("written", os.path.join(OUT, "pe_embedder_bands.npz"))


if __name__ == "__main__":
    if "--embedder-bands" in sys.argv:
        embedder_bands_golden()
    else:
        main()
        embedder_bands_golden()
