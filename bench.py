#!/usr/bin/env python
"""bench.py -- headline benchmark of the volume-learning hot path (BASELINE.json metric).

Workload (config C3 of BASELINE.json): coarse+fine NeRF TRAINING, 64 stratified + 128 importance samples per ray,
8192-ray batch per GPU (weak scaling: global batch = 8192 * n_gpus, ray-sharded data parallel, one NCCL all-reduce
of the flat fp32 gradient per optimiser step), reference iteration semantics (coarse step -> coarse re-forward ->
detached inverse-CDF resample + merge -> fine step).  Metric: train rays/s (whole job).  Also reported in the same
JSON line: render Msamples/s of a coarse+fine full-frame 800x800 render (config C5).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle restatement)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # multi-GPU (driver launches it this way)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 8192
N_SAMPLES, N_IMPORTANCE = 64, 128
# algorithmic (unpadded) FLOPs, SURVEY 8d / BASELINE.md: fwd 593408 MAC/pt, train 1744512 MAC/pt
FWD_FLOP_PT = 2 * 593408
TRAIN_FLOP_PT = 2 * 1744512
FLOP_PER_RAY = N_SAMPLES * TRAIN_FLOP_PT + N_SAMPLES * FWD_FLOP_PT + (N_SAMPLES + N_IMPORTANCE) * TRAIN_FLOP_PT
METRIC = "train_rays_per_s_coarse+fine_64+128"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(B, seed):
    """Synthetic Blender-style rays (pose_spherical camera, near 2 / far 6) and targets in [0,1]."""
    rng = np.random.default_rng(seed)
    theta = rng.uniform(-180, 180)
    t, ph = np.deg2rad(theta), np.deg2rad(-30.0)
    # camera on a sphere of radius 4 looking at the origin (ops/pose.py convention)
    cam = np.array([4 * np.cos(ph) * np.sin(t), 4 * np.cos(ph) * np.cos(t), -4 * np.sin(ph)], np.float32)
    fwd = -cam / np.linalg.norm(cam)
    up = np.array([0, 0, 1], np.float32)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    upv = np.cross(right, fwd)
    focal = 0.5 * 400 / np.tan(0.5 * 0.6911112)
    px = rng.choice(400 * 400, size=B, replace=False)
    i, j = (px % 400).astype(np.float32), (px // 400).astype(np.float32)
    d = ((i - 200) / focal)[:, None] * right + (-(j - 200) / focal)[:, None] * upv + fwd[None, :]
    o = np.broadcast_to(cam, d.shape).copy()
    target = rng.random(size=(B, 3)).astype(np.float32)
    return o.astype(np.float32), d.astype(np.float32), target


# ====================================================================================== reference arm (CPU oracle)
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import models as omodels, rendering as orend, training as otrain
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = 256  # rays per CPU step (bounded sample of the 8192-ray step)
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=1, **kw), omodels.NeRF(seed=2, **kw)
    opt = otrain.AdamMLX(5e-4)
    qf = orend.make_query_fn(10, 4)
    o, d, tgt = synth_batch(sample, 0)
    u = np.random.default_rng(1).random(size=(sample, N_IMPORTANCE), dtype=np.float32)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        otrain.train_iteration(oc, of, opt, o, d, tgt, u, qf, n_samples=N_SAMPLES)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 coarse+fine NeRF training step, 64+128 samples/ray, reference iteration semantics",
                       "rays_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{sample}-ray slice of the 8192-ray C3 step, torch-CPU fp32 restatement of the "
                                       "reference (MLX unavailable in this image)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ====================================================================================== this repo's arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nerf_meets_mlx_b200 import _lib_loader as L
    from nerf_meets_mlx_b200.models.NeRF import default_args
    from nerf_meets_mlx_b200.training import NeRFTrainer, assemble_rays

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU path in nerf_meets_mlx_b200)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.rays_per_gpu if args.rays_per_gpu > 0 else RAYS_PER_GPU
    # the iteration is captured once into CUDA graphs and replayed (no launch gaps); data parallel: three graphs cut at the
    # two gradient all-reduces, which run eagerly between the replays
    tr = NeRFTrainer(default_args(N_importance=N_IMPORTANCE, n_depth_samples=N_SAMPLES), device=dev, max_rays=B,
                     use_cuda_graph=(not args.no_graph))
    n_batches = 4
    host = []
    for k in range(n_batches):
        o, d, t = synth_batch(B, 1000 * rank + k)
        host.append(tuple(torch.from_numpy(a).pin_memory() for a in (o, d, t)))
    resident = [tuple(a.to(dev) for a in hb) for hb in host]
    u_res = [torch.rand((B, N_IMPORTANCE), device=dev) for _ in range(n_batches)]

    def step_resident(i):
        o, d, t = resident[i % n_batches]
        return tr.train_iteration(o, d, t, u_vals=u_res[i % n_batches])

    def step_e2e(i):
        o, d, t = (a.to(dev, non_blocking=True) for a in host[i % n_batches])
        out = tr.train_iteration(o, d, t)
        return float(out["loss_fine"].item())  # D2H read of the step's result

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.launch_count()
        it0 = tr.iteration
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.launch_count() - launches0
        if tr._graph is not None:  # replayed iterations launch their kernels from the graph, not through the C ABI
            launches += (tr.iteration - it0) * tr.graph_launches
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms / steps, launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    # ---- live per-kernel timing (CUDA events on the launching stream, inside libnmx) over the same steps
    peaks, peak_src = load_peaks()
    lib = L.lib()
    import ctypes
    lib.nmx_profile_enable(1)
    prof_steps = max(1, min(args.steps, 3))
    graph_mode, tr.use_cuda_graph = tr.use_cuda_graph, False  # the library's event timers need eager launches
    for i in range(prof_steps):
        step_resident(i)
    tr.use_cuda_graph = graph_mode
    prof = {}
    for kind in range(5):
        ms_k, fl_k, n_k = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        lib.nmx_profile_report(kind, ctypes.byref(ms_k), ctypes.byref(fl_k), ctypes.byref(n_k))
        prof[kind] = (ms_k.value, fl_k.value, n_k.value)
    lib.nmx_profile_enable(0)

    # ---- secondary metric: coarse+fine full-frame render (C5), ray tiles sharded over ranks
    render = None
    if not args.no_render:
        H = W = 800
        n_rays = H * W
        shard = (n_rays + world - 1) // world
        o, d, _ = synth_batch(min(n_rays, 160000), 7)
        reps = (shard + o.shape[0] - 1) // o.shape[0]
        o_t = torch.from_numpy(np.tile(o, (reps, 1))[:shard]).to(dev)
        d_t = torch.from_numpy(np.tile(d, (reps, 1))[:shard]).to(dev)
        rays = assemble_rays(o_t, d_t, 2.0, 6.0)

        def frame(_):
            outs = []
            for s in range(0, shard, 32768):  # chunk = 32768 rays (render.py:245)
                outs.append(tr.render_rays_eval(rays[s:s + 32768])["rgb_map"])
            return torch.cat(outs)
        ms_frame, _ = timed(frame, 2, 1)
        render = {"metric": "render_Msamples_per_s_coarse+fine_64+192", "value": n_rays * 256 / (ms_frame * 1e-3) / 1e6,
                  "unit": "Msamples/s (MLP point evaluations, 256/ray)", "ms_per_frame": ms_frame,
                  "frame": "800x800", "chunk_rays": 32768,
                  "tensor_frac_of_sustained": n_rays * 256 * FWD_FLOP_PT / (ms_frame * 1e-3) / 1e12
                  / (world * peaks["bf16_tflops_sustained"])}

    if rank == 0:
        rays_total = B * world
        value = rays_total / (ms_step * 1e-3)
        e2e = rays_total / (ms_e2e * 1e-3)
        # dominant kernel: the fused MLP chain (forward inference / forward training / backward data gradients);
        # ALGORITHMIC flops (fwd 593408 MAC/pt, dgrad 557696 MAC/pt) over the measured launch time
        P_c, P_f = B * N_SAMPLES, B * (N_SAMPLES + N_IMPORTANCE)
        DGRAD_FLOP_PT = 2 * 557696
        alg = {2: P_c * FWD_FLOP_PT, 3: (P_c + P_f) * FWD_FLOP_PT, 4: (P_c + P_f) * DGRAD_FLOP_PT}
        chain_ms = sum(prof[k][0] for k in (2, 3, 4))
        chain_n = sum(prof[k][2] for k in (2, 3, 4))
        chain_alg = sum(alg[k] for k in (2, 3, 4)) * prof_steps
        achieved = chain_alg / (chain_ms * 1e-3) / 1e12 if chain_ms > 0 else 0.0
        sub = {}
        for k, name in ((2, "forward_inference"), (3, "forward_training"), (4, "backward_dgrad")):
            if prof[k][0] > 0:
                sub[name] = {"tflops": alg[k] * prof_steps / (prof[k][0] * 1e-3) / 1e12, "launches": int(prof[k][2]),
                             "ms_per_step": prof[k][0] / prof_steps}
        # wgrad: HBM-bound (reads dY[P,M] and X[P,N] bf16 once per launch)
        # bf16 bytes per point: layer 0 (dY 512 + PE(pos) 128), six plain layers (512 + 512), the skip layer as ONE
        # dual-operand launch (dY 512 + h 512 + PE(pos) 128), the dir layer as one dual launch (d_hd 256 + h 512 +
        # PE(dir) 128).  Part of it is served by L2 (dY was written by the chain just before), so the figure can exceed
        # the HBM peak.
        wg_bytes_per_point = 640 + 7 * 1024 + 1152 + 896
        wg_bytes = wg_bytes_per_point * (P_c + P_f) * prof_steps
        wg_ms = prof[1][0]
        chain_traffic = {}
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_v9_chain_traffic.json")
        if world == 1 and os.path.exists(tpath):  # ncu --set full capture of the same six launches per step (N = 1 sizes)
            with open(tpath) as f:
                chain_traffic = json.load(f)
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3 coarse+fine NeRF training step (reference iteration semantics), 64 stratified + "
                                   "128 importance samples/ray, 8x256 MLPs with view-dir head",
                       "rays_per_gpu": B, "global_rays": rays_total, "parallelism": f"dp{world} ray-sharded",
                       "cuda_graph": bool(tr._graph is not None),
                       "l2": "per-step working set (~20 GB of saved activations and data gradients) >> 126 MB L2; "
                             "4 rotating input batches"},
            "clocks": clocks,
            "e2e": {"value": e2e, "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(B * 9 * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "step_tensor_frac": {"algorithmic_tflops": value * FLOP_PER_RAY / 1e12 / world,
                                 "peak_tflops_sustained": peaks["bf16_tflops_sustained"],
                                 "frac": value * FLOP_PER_RAY / 1e12 / world / peaks["bf16_tflops_sustained"],
                                 "flop_per_ray": FLOP_PER_RAY},
            "roofline": {"bound": "tensor", "kernel": "mlp_chain_kernel (fused whole-MLP forward / backward chains, tcgen05 + TMEM)",
                         "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops_sustained"],
                         "peak_source": peak_src + " sustained (timed inside a long step)",
                         "launches": int(chain_n), "avg_launch_ms": chain_ms / max(chain_n, 1),
                         "traffic": chain_traffic.get("dram_bytes_per_launch"), "traffic_source": chain_traffic.get("source"),
                         "modes": sub, "profiled_steps": prof_steps,
                         "wgrad": {"bound": "hbm", "achieved": wg_bytes / (wg_ms * 1e-3) / 1e9 if wg_ms > 0 else 0.0,
                                   "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": (wg_bytes / (wg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if wg_ms > 0 else 0.0,
                                   "tflops": (prof[1][1] / (wg_ms * 1e-3) / 1e12) if wg_ms > 0 else 0.0,
                                   "launches": int(prof[1][2]), "ms_per_step": wg_ms / prof_steps}},
            "render": render,
        }
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def cpu_baseline():
    """The oracle (restated reference) timed on this box's host cores on a bounded sample of the same step."""
    import torch
    from oracle import models as omodels, rendering as orend, training as otrain
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = 256
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=1, **kw), omodels.NeRF(seed=2, **kw)
    opt = otrain.AdamMLX(5e-4)
    qf = orend.make_query_fn(10, 4)
    o, d, tgt = synth_batch(sample, 0)
    u = np.random.default_rng(1).random(size=(sample, N_IMPORTANCE), dtype=np.float32)
    otrain.train_iteration(oc, of, opt, o, d, tgt, u, qf, n_samples=N_SAMPLES)  # warm-up
    ts = []
    t_end = time.perf_counter() + 15.0
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 10):
        t0 = time.perf_counter()
        otrain.train_iteration(oc, of, opt, o, d, tgt, u, qf, n_samples=N_SAMPLES)
        ts.append(time.perf_counter() - t0)
    return {"value": sample / float(np.median(ts)), "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{sample}-ray slice of the C3 step x{len(ts)} (median), torch-CPU fp32 restatement of the reference "
                      "incl. the reference's resampling arithmetic; MLX not installable in this image"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays-per-gpu", type=int, default=0, help="override the 8192 rays per GPU of the headline step")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the training iteration eagerly (no CUDA graph)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
