"""Parity at the BASELINE.json sizes against the ORACLE (not against the CUDA path itself): one C3 iteration at 8192 rays
x (64 + 128) samples through the graph-replayed trainer, one 32768-ray render_rays_eval chunk, the C4 hash grid at
262144 points, and the 2048-ray resampling golden of the real reference function.  The measured errors are written to
gpurun_out/r2_fullsize_parity.json (copied to profiles/); the gradient bounds below are 2x the values measured on B200.
The oracle needs ~40 s of host time and ~30 GB of host memory for the 8192-ray iteration."""
import json
import os
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoding as oenc, models as omodels, rendering as orend, training as otrain, sampling as osamp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)
# normwise gradient error of the bf16 CUDA path vs the fp32 oracle at full size: 2x the largest value measured on B200
# (profiles/r2_fullsize_parity.json); the loose 1.5e-1 of round 1 is gone
GRAD_BOUND = {"coarse": 1.3e-2, "fine": 2.4e-2}  # measured maxima 6.4e-3 (coarse), 1.17e-2 (fine, alpha_linear.bias)


def _record(key, value):
    path = os.path.join(ROOT, "gpurun_out", "r2_fullsize_parity.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    d = {}
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
    d[key] = value
    with open(path, "w") as f:
        json.dump(d, f, indent=1)


def _scene_rays(B, seed):
    rng = np.random.default_rng(seed)
    o = (rng.uniform(-0.5, 0.5, size=(B, 3)) + np.array([0, 0, 4.0])).astype(np.float32)
    d = rng.standard_normal(size=(B, 3)).astype(np.float32)
    d[:, 2] = -np.abs(d[:, 2]) - 1.0
    d /= np.linalg.norm(d, axis=-1, keepdims=True) * 0.9
    return o, d.astype(np.float32), rng.random(size=(B, 3)).astype(np.float32)


def _nets(seed):
    oc, of = omodels.NeRF(seed=seed, **KW), omodels.NeRF(seed=seed + 1, **KW)
    for m in (oc, of):  # keep sigma away from 0: the reference's 1e10 last-bin delta makes alpha_last = step(sigma_last)
        m.params["alpha_linear.bias"] = m.params["alpha_linear.bias"] * 0 + 0.5
    return oc, of


def test_c3_iteration_8192_rays_graph_replayed_vs_oracle():
    """Iteration 1 runs eagerly (the trainer's warm-up), iteration 2 is captured and REPLAYED from the CUDA graph: both
    must reproduce the oracle's coarse and fine losses to 1e-2 at the full 8192 x 64 / 8192 x 192 size (all 148 CTAs x
    80+ tiles, > 2^31-byte workspace offsets), and iteration 2's gradients stay within the measured bf16 bound."""
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer
    B, n, N = 8192, 64, 128
    args = default_args(N_importance=N, n_depth_samples=n)
    tr = NeRFTrainer(args, max_rays=B, use_cuda_graph=True)
    oc, of = _nets(1)
    tr.coarse.load_reference_parameters(oc.params)
    tr.fine.load_reference_parameters(of.params)
    opt = otrain.AdamMLX(args.lrate, shared_state=True)
    qf = orend.make_query_fn(10, 4)
    rec = {"rays": B, "samples": [n, n + N], "iterations": []}
    t_oracle = 0.0
    for it in range(2):
        o, d, target = _scene_rays(B, seed=10 + it)
        u = np.random.default_rng(50 + it).random(size=(B, N), dtype=np.float32)
        opt.learning_rate = otrain.lr_schedule(it)
        t0 = time.perf_counter()
        ref = otrain.train_iteration(oc, of, opt, o, d, target, u, qf, n_samples=n, near=2.0, far=6.0, white_bkgd=True)
        t_oracle += time.perf_counter() - t0
        out = tr.train_iteration(*(torch.from_numpy(a).cuda() for a in (o, d, target)), u_vals=torch.from_numpy(u).cuda())
        lc, lf = out["loss_coarse"].item(), out["loss_fine"].item()
        e_c, e_f = abs(lc - ref["loss_coarse"]) / ref["loss_coarse"], abs(lf - ref["loss_fine"]) / ref["loss_fine"]
        ge = {}
        for tag, model, gbuf, gref in (("coarse", tr.coarse, tr._g_coarse, ref["grads_coarse"]),
                                       ("fine", tr.fine, tr._g_fine, ref["grads_fine"])):
            got = model.split_flat(gbuf)
            ge[tag] = {k: float((got[k].cpu() - g).norm() / (g.norm() + 1e-20)) for k, g in gref.items()}
        zf = out["z_fine"].cpu().numpy()
        z_close = float(np.mean(np.abs(zf - ref["z_fine"]) < 2e-2))
        rec["iterations"].append({"mode": "eager" if it == 0 else "cuda graph replay", "loss_coarse": lc, "loss_fine": lf,
                                  "oracle_loss_coarse": ref["loss_coarse"], "oracle_loss_fine": ref["loss_fine"],
                                  "rel_err_loss_coarse": e_c, "rel_err_loss_fine": e_f, "grad_rel_err": ge,
                                  "z_fine_within_2e-2": z_close})
        _record("c3_iteration_8192", dict(rec, oracle_seconds=t_oracle))
        assert e_c < 1e-2 and e_f < 1e-2, (it, lc, ref["loss_coarse"], lf, ref["loss_fine"])
        assert np.all(np.diff(zf, axis=-1) >= 0) and z_close > 0.97, z_close
        for tag in ("coarse", "fine"):
            worst = max(ge[tag].items(), key=lambda kv: kv[1])
            assert worst[1] < GRAD_BOUND[tag], (it, tag, worst)
        del ref
    assert tr._graph is not None  # iteration 2 really came from the graph


def test_render_chunk_32768_rays_vs_oracle():
    """One full render chunk (32768 rays, render.py:245) of render_rays_eval: rgb / acc / coarse rgb within 1e-2 of the
    fp32 oracle (north_star's bound; round 1 asserted 2e-2 on a 12 x 16 image only)."""
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer, assemble_rays
    B, n, N = 32768, 64, 128
    tr = NeRFTrainer(default_args(N_importance=N, n_depth_samples=n), max_rays=128)
    oc, of = _nets(11)
    tr.coarse.load_reference_parameters(oc.params)
    tr.fine.load_reference_parameters(of.params)
    o, d, _ = _scene_rays(B, seed=3)
    u = np.random.default_rng(4).random(size=(B, N), dtype=np.float32)
    rays = assemble_rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), 2.0, 6.0)
    got = tr.render_rays_eval(rays, torch.from_numpy(u).cuda())
    qf = orend.make_query_fn(10, 4)
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = orend.render_rays_eval(otrain.assemble_rays(o, d, 2.0, 6.0), oc, qf, n, white_bkgd=True, N_importance=N,
                                     network_fine=of, u_vals=u)
    rec = {"rays": B, "oracle_seconds": time.perf_counter() - t0}
    for k in ("rgb_map", "acc_map", "rgb_coarse"):
        want = ref[k].numpy()
        rec[k] = float(np.abs(got[k].cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-6))
    _record("render_chunk_32768", rec)
    for k in ("rgb_map", "acc_map", "rgb_coarse"):
        assert rec[k] < 1e-2, (k, rec[k])


def test_c4_hashgrid_262144_points_vs_oracle():
    """C4 at its BASELINE size: corner indices bit-exact, encoded features bit-exact, table gradient vs the fp64 oracle."""
    from nerf_meets_mlx_b200 import ops
    L, F, T, P = 16, 2, 19, 262144
    rng = np.random.default_rng(7)
    x = rng.random(size=(P, 3), dtype=np.float32)
    tables = rng.uniform(-1e-4, 1e-4, size=(L, 1 << T, F)).astype(np.float32)
    res = oenc.hashgrid_scaled_res(L, 16, 2048)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out, idx = ops.hashgrid_fwd(dev(x), dev(tables), dev(res), T, return_idx=True)
    ref_idx, _ = oenc.hashgrid_corner_indices(x, res, T)
    np.testing.assert_array_equal(idx.cpu().numpy().astype(np.int64), ref_idx)
    np.testing.assert_array_equal(out.cpu().numpy(), oenc.hashgrid_encode(x, tables, res, T))
    d_out = rng.standard_normal(size=(P, L * F)).astype(np.float32)
    g = ops.hashgrid_bwd(dev(x), dev(res), dev(d_out), L, F, T).cpu().numpy()
    gref = oenc.hashgrid_backward(x, d_out, L, F, res, T)
    err = float(np.linalg.norm(g - gref) / np.linalg.norm(gref))
    _record("c4_hashgrid_262144", {"points": P, "indices_bit_exact": True, "features_bit_exact": True,
                                   "table_grad_rel_err_vs_fp64": err})
    np.testing.assert_allclose(g, gref, rtol=2e-4, atol=2e-5)
    assert err < 1e-5, err


def test_sample_pdf_large_golden_kernel(golden):
    """Kernel (own fp64-sum CDF) vs the REAL reference function on 2048 rays x 128 draws: the end-to-end index mismatch
    count is asserted and printed; the values agree wherever the indices do."""
    from nerf_meets_mlx_b200 import ops
    g = golden("sample_pdf_large")
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    r = ops.sample_pdf(dev(g["z"]), dev(g["w"]), dev(g["u"]), want_inds=True)
    inds = r["inds"].cpu().numpy().astype(np.int64)
    want = g["inds"].astype(np.int64)
    mism = int(np.sum(inds != want))
    print(f"sample_pdf kernel vs reference function: {mism} index mismatches of {inds.size} draws "
          f"(generator: {int(g['big_mismatch'])} of {int(g['big_draws'])} for the canonical CDF)")
    _record("sample_pdf_large", {"draws": int(inds.size), "index_mismatches": mism,
                                 "generator_draws": int(g["big_draws"]), "generator_mismatches": int(g["big_mismatch"])})
    assert mism == int(g["mismatch_2048"]) and mism <= 2
    ok = inds == want
    np.testing.assert_allclose(r["z_imp"].cpu().numpy()[ok], g["out"][ok], rtol=1e-5, atol=1e-5)
    _, o_inds = osamp.sample_pdf(g["z"], g["w"], g["u"], return_inds=True)
    np.testing.assert_array_equal(inds, o_inds)  # kernel == oracle bit for bit
