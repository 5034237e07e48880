"""Achieved HBM GB/s of the memory-bound kernels of the path (compositing, resampling, hash grid, encodings, ray
generation) against the measured B200 peak -- the north_star's "achieved HBM GB/s for encoding, compositing and
resampling".  ALGORITHMIC bytes per unit are SURVEY 8(d)'s figures; every kernel is timed with CUDA events over
rotating input sets whose total size exceeds 2x the 126 MB L2 (so nothing is served from cache between iterations),
calling the C ABI directly on preallocated outputs.

    python scripts/bench_aux.py [--out profiles/r1_aux_kernels.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import _lib_loader as L  # noqa: E402
from nerf_meets_mlx_b200._lib_loader import f32, f64, i32, i64, ptr, stream  # noqa: E402

PEAK = 6545.3
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK))
L2_BYTES = 126e6


def timed(fn, nsets, iters=40, warm=5):
    for i in range(warm):
        fn(i % nsets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i % nsets)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


def nsets_for(bytes_per_set):
    return max(2, int(np.ceil(2 * L2_BYTES / bytes_per_set)) + 1)


def report(rows, name, units, unit_name, alg_bytes_per_unit, sec, note=""):
    gbs = units * alg_bytes_per_unit / sec / 1e9
    rows.append({"kernel": name, "units": units, "unit": unit_name, "algorithmic_bytes_per_unit": alg_bytes_per_unit,
                 "us": sec * 1e6, "achieved_GBps": gbs, "frac_of_hbm_peak": gbs / PEAK, "note": note})
    print(f"{name:34s} {units:>9d} {unit_name:6s} {sec * 1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs / PEAK * 100:5.1f} %  {note}",
          file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    rows = []

    # ---- compositing (raw2outputs) forward / backward, C3 fine (n = 192) and coarse (n = 64)
    for n, B in ((192, 65536), (64, 131072)):
        per = B * (24 * n + 36)
        ns = nsets_for(per)
        raw = [torch.randn(B, n, 4, device=dev, generator=g) for _ in range(ns)]
        z = [torch.sort(rnd(B, n) * 4 + 2, -1).values.contiguous() for _ in range(ns)]
        d = torch.randn(B, 3, device=dev, generator=g)
        rgb, disp, acc, dep = (torch.empty(B, k, device=dev) for k in (3, 1, 1, 1))
        w = torch.empty(B, n, device=dev)
        sec = timed(lambda i: L.call("nmx_composite_fwd", ptr(raw[i]), ptr(z[i]), ptr(d), i32(3), ptr(None), f32(0.0), i32(1),
                                     ptr(rgb), ptr(disp), ptr(acc), ptr(w), ptr(dep), i64(B), i32(n), stream()), ns)
        report(rows, f"composite_fwd n={n}", B, "rays", 24 * n + 36, sec)
        d_rgb = torch.randn(B, 3, device=dev, generator=g)
        d_raw = torch.empty(B, n, 4, device=dev)
        sec = timed(lambda i: L.call("nmx_composite_bwd", ptr(raw[i]), ptr(z[i]), ptr(d), i32(3), ptr(None), f32(0.0), i32(1),
                                     ptr(d_rgb), ptr(None), ptr(None), ptr(None), ptr(None), ptr(d_raw), i64(B), i32(n),
                                     stream()), ns)
        report(rows, f"composite_bwd n={n}", B, "rays", 36 * n + 28, sec)
        del raw, z

    # ---- sample_pdf + merge (n = 64 coarse, N = 128 importance -> 192 merged)
    B, n, N = 262144, 64, 128
    ns = nsets_for(B * 1792)
    z = [torch.sort(rnd(B, n) * 4 + 2, -1).values.contiguous() for _ in range(ns)]
    w = [rnd(B, n) ** 4 for _ in range(ns)]
    u = [rnd(B, N) for _ in range(ns)]
    merged = torch.empty(B, n + N, device=dev)
    sec = timed(lambda i: L.call("nmx_sample_pdf_fwd", ptr(z[i]), ptr(w[i]), ptr(u[i]), ptr(None), f32(1e-5), ptr(None),
                                 ptr(None), ptr(None), ptr(merged), i64(B), i32(n), i32(N), stream()), ns)
    report(rows, "sample_pdf+merge 64+128", B, "rays", 1792, sec)
    del z, w, u, merged

    # ---- hash grid C4: L=16, T=2^19, F=2, 262144 queries (and 16.7 M = 262144 rays x 64)
    Lv, F, T = 16, 2, 19
    from nerf_meets_mlx_b200.encoding import MultiHashEncoding
    enc = MultiHashEncoding(3, Lv, 16, 2048, F, T, device=dev)
    tables = enc.hash_table.detach().contiguous()
    res = enc.scaled_res.contiguous()
    for P in (262144, 262144 * 16):
        ns = max(2, nsets_for(P * 1164))
        x = [rnd(P, 3) for _ in range(ns)]
        out = torch.empty(P, Lv * F, device=dev)
        sec = timed(lambda i: L.call("nmx_hashgrid_fwd", ptr(x[i]), ptr(tables), ptr(res), ptr(out), ptr(None), i64(P), i32(Lv),
                                     i32(F), i32(T), stream()), ns, iters=20)
        report(rows, "hashgrid_fwd L16 T2^19 F2", P, "points", 1164, sec, "64 MiB table is L2-resident: gathers hit L2")
        d_out = torch.randn(P, Lv * F, device=dev, generator=g)
        d_tab = torch.zeros(Lv, 1 << T, F, device=dev)
        sec = timed(lambda i: L.call("nmx_hashgrid_bwd", ptr(x[i]), ptr(res), ptr(d_out), ptr(d_tab), i64(P), i32(Lv), i32(F),
                                     i32(T), stream()), ns, iters=20)
        report(rows, "hashgrid_bwd L16 T2^19 F2", P, "points", 2188, sec, "atomic RMW on the L2-resident gradient table")
        del x, out, d_out, d_tab

    # ---- stand-alone encodings (parity kernels; the product path fuses PE into the MLP chain)
    P = 1 << 21
    ns = nsets_for(P * 372)
    x = [rnd(P, 3) * 8 - 4 for _ in range(ns)]
    out = torch.empty(P, 63, device=dev)
    sec = timed(lambda i: L.call("nmx_pe_embedder_fwd", ptr(x[i]), ptr(out), i64(P), i32(3), i32(10), i32(1), stream()), ns)
    report(rows, "pe_embedder 3->63", P, "points", 12 + 252, sec, "write-dominated; sincosf-bound (write-only HBM peak is 6.3 TB/s, profiles/r2_bw_probe.txt)")
    out = torch.empty(P, 25, device=dev)
    sec = timed(lambda i: L.call("nmx_sh_encode_fwd", ptr(x[i]), i32(3), ptr(out), i64(P), i32(4), stream()), ns)
    report(rows, "sh_encode deg 4", P, "points", 12 + 100, sec, "write-dominated")
    del x, out

    # ---- ray generation: one 800x800 view x 16 rotating poses
    H = W = 800
    K = np.array([[1111.1, 0, 400.0], [0, 1111.1, 400.0], [0, 0, 1]])
    c2w = torch.eye(4, device=dev)[:3].contiguous()
    rays = [torch.empty(H * W, 11, device=dev) for _ in range(12)]
    sec = timed(lambda i: L.call("nmx_gen_rays", ptr(c2w), i32(4), f64(K[0][0]), f64(K[1][1]), f64(K[0][2]), f64(K[1][2]),
                                 i32(H), i32(W), ptr(None), i64(H * W), f32(2.0), f32(6.0), ptr(rays[i]), i32(11), ptr(None),
                                 i32(0), ptr(None), stream()), 12)
    report(rows, "gen_rays 800x800 -> [B,11]", H * W, "rays", 44, sec, "write-only")

    line = {"hbm_peak_GBps": PEAK, "peak_source": "MEASURED_PEAKS.json (copy)", "timing": "CUDA events, rotating inputs > 2x L2",
            "kernels": rows}
    print(json.dumps(line))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(line, f, indent=1)


if __name__ == "__main__":
    main()
