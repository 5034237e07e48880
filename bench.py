#!/usr/bin/env python
"""bench.py -- headline benchmark of the volume-learning hot path (BASELINE.json metric).

Headline workload (config C3 of BASELINE.json): coarse+fine NeRF TRAINING, 64 stratified + 128 importance samples per
ray, reference iteration semantics (coarse step -> coarse re-forward -> detached inverse-CDF resample + merge -> fine
step), ray-sharded data parallel with the gradient exchange fused into the Adam kernel over NVLink peer memory.
Metric: train rays/s (whole job).  The headline line is WEAK scaling (8192 rays per GPU); the same JSON line carries

  "strong"   : the configuration as BASELINE.json states it -- ONE 8192-ray step sharded 8192/N rays per GPU,
  "render"   : C5, coarse+fine 800x800 frames, ray tiles sharded over the ranks and gathered (Msamples/s),
  "configs"  : C1 (image learning, us/step), C2 (coarse NeRF, 4096 rays), C4 (hash grid + tiny MLP, 262144 points),
  "dp_check" : (N > 1) replicas bit-identical after the timed steps, sharded frame == single-rank frame,
  "roofline" / "cpu_baseline" / "e2e" / "clocks" / "gpu_launches" as the bench contract asks.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --config C1|C2|C4 [--steps K] [--warmup W]      # one other config as the printed line
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path (oracle restatement)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # multi-GPU (driver launches it this way)
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 8192
GLOBAL_RAYS_STRONG = 8192
N_SAMPLES, N_IMPORTANCE = 64, 128
# algorithmic (unpadded) FLOPs, SURVEY 8d / BASELINE.md: fwd 593408 MAC/pt, train 1744512 MAC/pt
FWD_FLOP_PT = 2 * 593408
TRAIN_FLOP_PT = 2 * 1744512
DGRAD_FLOP_PT = 2 * 557696
FLOP_PER_RAY = N_SAMPLES * TRAIN_FLOP_PT + N_SAMPLES * FWD_FLOP_PT + (N_SAMPLES + N_IMPORTANCE) * TRAIN_FLOP_PT
FLOP_PER_RAY_C2 = N_SAMPLES * TRAIN_FLOP_PT                 # 223.3 MFLOP/ray (coarse only)
FLOP_PER_PIXEL_C1 = 2839000                                  # image net, train
HASH_BYTES_PT = 1164 + 2188                                  # hash grid fwd + bwd algorithmic bytes / point
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


def synth_batch(B, seed, side=400):
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
    focal = 0.5 * side / np.tan(0.5 * 0.6911112)
    px = rng.choice(side * side, size=B, replace=False)
    i, j = (px % side).astype(np.float32), (px // side).astype(np.float32)
    d = ((i - side / 2) / focal)[:, None] * right + (-(j - side / 2) / focal)[:, None] * upv + fwd[None, :]
    o = np.broadcast_to(cam, d.shape).copy()
    target = rng.random(size=(B, 3)).astype(np.float32)
    return o.astype(np.float32), d.astype(np.float32), target


def synth_image(side=256):
    """C1 target: analytic 256x256 RGB pattern (2-D sinusoids + checker + gradient) in [0, 1]."""
    yy, xx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    img = np.stack([0.5 + 0.5 * np.sin(xx / 17.0) * np.cos(yy / 11.0), 0.25 + 0.5 * (yy / side) + 0.25 * np.cos(xx / 23.0),
                    ((xx // 32 + yy // 32) % 2).astype(np.float64)], -1).astype(np.float32)
    coords = np.stack([yy, xx], -1).reshape(-1, 2).astype(np.int32)
    return coords, img.reshape(-1, 3)


# ====================================================================================== CPU arm (oracle restatement)
def _oracle_step_fn():
    import torch
    from oracle import models as omodels, rendering as orend, training as otrain
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kw = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
              list_skip_connection_layers=[4], is_use_view_directions=True)
    oc, of = omodels.NeRF(seed=1, **kw), omodels.NeRF(seed=2, **kw)
    opt = otrain.AdamMLX(5e-4)
    qf = orend.make_query_fn(10, 4)

    def step(sample, seed=0):
        o, d, tgt = synth_batch(sample, seed)
        u = np.random.default_rng(seed + 1).random(size=(sample, N_IMPORTANCE), dtype=np.float32)
        t0 = time.perf_counter()
        otrain.train_iteration(oc, of, opt, o, d, tgt, u, qf, n_samples=N_SAMPLES)
        return time.perf_counter() - t0
    return step, cores


def run_reference(args):
    """The reference's CPU implementation of the C3 step (torch-CPU fp32 restatement: MLX is not installable in this
    image), all host threads.  The step is the FULL 8192-ray batch when the run fits a few minutes, else the largest
    power-of-two slice that does (stated in config.rays_per_step / same_config)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    step, cores = _oracle_step_fn()
    step(256)                      # warm-up of the thread pool / allocator
    rate = 1024 / step(1024)       # calibration: rays/s at 1024 rays
    budget = 200.0                 # seconds for the warm-up + timed steps
    sample = GLOBAL_RAYS_STRONG
    while sample > 256 and (args.warmup + args.steps) * sample / rate > budget:
        sample //= 2
    times = []
    for it in range(args.warmup + args.steps):
        dt = step(sample, seed=it)
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3 coarse+fine NeRF training step, 64+128 samples/ray, reference iteration semantics",
                       "rays_per_step": sample, "same_config": sample == GLOBAL_RAYS_STRONG,
                       "calibration_rays_per_s_at_1024": rate},
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"{sample}-ray step of the 8192-ray C3 iteration x{len(times)}, torch-CPU fp32 "
                                       "restatement of the reference (MLX unavailable in this image)"},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def cpu_baseline():
    """The oracle (restated reference) timed on this box's host cores on a bounded sample of the same step:
    full 8192-ray steps when they fit ~25 s, else the largest power-of-two slice that does."""
    step, cores = _oracle_step_fn()
    step(256)
    rate = 1024 / step(1024)
    sample = GLOBAL_RAYS_STRONG
    while sample > 256 and 2 * sample / rate > 25.0:
        sample //= 2
    ts = [step(sample, seed=10 + k) for k in range(2)]
    return {"value": sample / float(np.median(ts)), "unit": "rays/s", "cores": cores, "kind": "port",
            "same_config": sample == GLOBAL_RAYS_STRONG,
            "sample": f"{sample}-ray C3 step x{len(ts)} (median; 8192 = the full step), torch-CPU fp32 restatement of the "
                      "reference incl. its resampling arithmetic; MLX not installable in this image"}


# ====================================================================================== this repo's arm
class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from nerf_meets_mlx_b200 import _lib_loader as L
        self.torch, self.dist, self.L, self.args = torch, dist, L, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU path in nerf_meets_mlx_b200)")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks, self.peak_src = load_peaks()
        self.group = dist.group.WORLD if self.world > 1 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, trainer=None, collective=True):
        """W untimed + K timed calls bracketed by barrier + synchronize; CUDA events; max over ranks."""
        torch, L = self.torch, self.L
        for i in range(warmup):
            fn(i)
        self.barrier() if collective else torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.launch_count()
        it0 = trainer.iteration if trainer is not None else 0
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        self.barrier() if collective else torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = L.launch_count() - launches0
        if trainer is not None and trainer._graph is not None:  # replayed iterations launch from the graph
            launches += (trainer.iteration - it0) * trainer.graph_launches
        if self.world > 1 and collective:
            tms = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(tms, op=self.dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms / steps, launches

    # ---------------------------------------------------------------------------------- C3 (headline + strong)
    def make_batches(self, B, n_batches, global_batch=None):
        """Rotating pinned host batches + resident copies.  global_batch: this rank's slice of ONE global batch."""
        torch = self.torch
        host = []
        for k in range(n_batches):
            if global_batch is None:
                o, d, t = synth_batch(B, 1000 * self.rank + k)
            else:
                o, d, t = (a[self.rank * B:(self.rank + 1) * B] for a in synth_batch(global_batch, 77 + k))
            host.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (o, d, t)))
        resident = [tuple(a.to(self.dev) for a in hb) for hb in host]
        u_res = [torch.rand((B, N_IMPORTANCE), device=self.dev) for _ in range(n_batches)]
        return host, resident, u_res

    def run_c3(self):
        args = self.args
        from nerf_meets_mlx_b200.models.NeRF import default_args
        from nerf_meets_mlx_b200.training import NeRFTrainer
        B = args.rays_per_gpu if args.rays_per_gpu > 0 else RAYS_PER_GPU
        tr = NeRFTrainer(default_args(N_importance=N_IMPORTANCE, n_depth_samples=N_SAMPLES), device=self.dev, max_rays=B,
                         use_cuda_graph=(not args.no_graph), process_group=None)
        self.tr = tr
        n_batches = 4
        host, resident, u_res = self.make_batches(B, n_batches)

        def step_resident(i):
            o, d, t = resident[i % n_batches]
            return tr.train_iteration(o, d, t, u_vals=u_res[i % n_batches])

        def step_e2e(i):
            o, d, t = (a.to(self.dev, non_blocking=True) for a in host[i % n_batches])
            out = tr.train_iteration(o, d, t)
            return float(out["loss_fine"].item())  # D2H read of the step's result

        sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        if sampler:
            sampler.start()
        ms_step, launches = self.timed(step_resident, args.steps, args.warmup, tr)
        clocks = sampler.stop() if sampler else None
        ms_e2e, _ = self.timed(step_e2e, args.steps, max(3, args.warmup // 2), tr)

        # ---- live per-kernel timing (CUDA events on the launching stream, inside libnmx) over eager steps
        import ctypes
        lib = self.L.lib()
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

        # ---- strong scaling: ONE 8192-ray step sharded over the ranks (the configuration BASELINE.json states)
        strong = None
        if self.world > 1 and not args.no_strong and B == RAYS_PER_GPU:
            Bs = GLOBAL_RAYS_STRONG // self.world
            _, res_s, u_s = self.make_batches(Bs, n_batches, global_batch=GLOBAL_RAYS_STRONG)

            def step_strong(i):
                o, d, t = res_s[i % n_batches]
                return tr.train_iteration(o, d, t, u_vals=u_s[i % n_batches])
            ms_strong, _ = self.timed(step_strong, args.steps, args.warmup + 2, tr)
            # the same step on ONE GPU, measured in this run on rank 0 alone (the other ranks wait at the barrier)
            ms_n1 = None
            if self.rank == 0:
                tr1 = NeRFTrainer(default_args(N_importance=N_IMPORTANCE, n_depth_samples=N_SAMPLES), device=self.dev,
                                  max_rays=GLOBAL_RAYS_STRONG, use_cuda_graph=(not args.no_graph), data_parallel=False)
                _, res_1, u_1 = self.make_batches(GLOBAL_RAYS_STRONG, 2)
                ms_n1, _ = self.timed(lambda i: tr1.train_iteration(*res_1[i % 2], u_vals=u_1[i % 2]), args.steps,
                                      args.warmup + 2, tr1, collective=False)
                tr1.close()
                del tr1
            self.barrier()
            strong = {"scaling": "strong", "global_rays": GLOBAL_RAYS_STRONG, "rays_per_gpu": Bs,
                      "value": GLOBAL_RAYS_STRONG / (ms_strong * 1e-3), "unit": "rays/s", "ms_per_step": ms_strong,
                      "n1_ms_per_step_same_run": ms_n1,
                      "speedup_vs_n1_same_run": (ms_n1 / ms_strong) if ms_n1 else None,
                      "tensor_frac_of_sustained": GLOBAL_RAYS_STRONG / (ms_strong * 1e-3) * FLOP_PER_RAY / 1e12
                      / (self.world * self.peaks["bf16_tflops_sustained"])}
        elif self.world == 1 and B == RAYS_PER_GPU:
            strong = {"scaling": "strong", "global_rays": GLOBAL_RAYS_STRONG, "rays_per_gpu": B,
                      "value": B / (ms_step * 1e-3), "unit": "rays/s", "ms_per_step": ms_step,
                      "n1_ms_per_step_same_run": ms_step, "speedup_vs_n1_same_run": 1.0,
                      "note": "at N = 1 the strong configuration is the headline step"}
        return dict(B=B, ms_step=ms_step, ms_e2e=ms_e2e, launches=launches, clocks=clocks, prof=prof,
                    prof_steps=prof_steps, strong=strong)

    # ---------------------------------------------------------------------------------- C5 render
    def run_render(self, tr):
        """Coarse+fine 800x800 frames through NeRFTrainer.render_frame: ray tiles sharded over the ranks, chunks of
        32768 rays (render.py:245), rgb/disp/acc tiles all_gathered into the frame inside the timed region."""
        torch = self.torch
        from nerf_meets_mlx_b200 import ops
        H = W = 800
        n_rays = H * W
        o, d, _ = synth_batch(n_rays, 7, side=800)
        rays = ops.assemble_rays(torch.from_numpy(o).to(self.dev), torch.from_numpy(d).to(self.dev), 2.0, 6.0)
        u = torch.rand((n_rays, N_IMPORTANCE), device=self.dev, generator=torch.Generator(device=self.dev).manual_seed(3))

        def frame(_i):
            return tr.render_frame(rays, chunk=32768, u_vals=u, process_group=self.group)
        ms_est, _ = self.timed(frame, 2, 2)
        n_frames = int(min(64, max(8, math.ceil(1200.0 / ms_est))))
        ms_frame, _ = self.timed(frame, n_frames, 1)
        tf = n_rays * 256 * FWD_FLOP_PT / (ms_frame * 1e-3) / 1e12 / self.world
        return {"metric": "render_Msamples_per_s_coarse+fine_64+192", "value": n_rays * 256 / (ms_frame * 1e-3) / 1e6,
                "unit": "Msamples/s (MLP point evaluations, 256/ray)", "ms_per_frame": ms_frame, "frames_timed": n_frames,
                "timed_s": ms_frame * n_frames * 1e-3, "frame": "800x800", "chunk_rays": 32768,
                "sharding": f"{self.world} contiguous ray tile(s), rgb/disp/acc all_gathered per frame",
                "tensor_frac_of_sustained": tf / self.peaks["bf16_tflops_sustained"],
                "tensor_frac_of_burst": tf / self.peaks["bf16_tflops"]}

    # ---------------------------------------------------------------------------------- dp_check
    def run_dp_check(self, tr):
        torch, dist = self.torch, self.dist
        from nerf_meets_mlx_b200 import ops
        cs = torch.tensor(tr.parameter_checksums(), dtype=torch.float64, device=self.dev)
        allcs = [torch.empty_like(cs) for _ in range(self.world)]
        dist.all_gather(allcs, cs)
        same = all(torch.equal(allcs[0], c) for c in allcs)
        err = tr.xchg.error() if tr.xchg is not None else 0
        # sharded frame vs the same frame rendered by rank 0 alone
        n = 200 * 200
        o, d, _ = synth_batch(n, 9)
        rays = ops.assemble_rays(torch.from_numpy(o).to(self.dev), torch.from_numpy(d).to(self.dev), 2.0, 6.0)
        u = torch.rand((n, N_IMPORTANCE), device=self.dev, generator=torch.Generator(device=self.dev).manual_seed(4))
        sharded = tr.render_frame(rays, chunk=4096, u_vals=u, process_group=self.group)
        ok = torch.ones(1, device=self.dev)
        frame_sum = None
        if self.rank == 0:
            alone = tr.render_frame(rays, chunk=4096, u_vals=u)
            ok[0] = float(all(torch.equal(alone[k], sharded[k]) for k in alone))
            frame_sum = float(sharded["rgb_map"].double().sum().item())
        dist.broadcast(ok, src=0)
        return {"param_checksums_identical_on_all_ranks": bool(same),
                "param_checksum_rank0": [float(x) for x in allcs[0].tolist()],
                "peer_exchange_error": int(err), "peer_exchange": tr.xchg is not None,
                "sharded_frame_equals_single_rank_frame": bool(ok.item() == 1.0), "frame_rgb_checksum": frame_sum}

    # ---------------------------------------------------------------------------------- C1 / C2 / C4
    def run_c1(self, steps, warmup):
        torch = self.torch
        from nerf_meets_mlx_b200.learners import ImageLearner
        coords, img = synth_image(256)
        lrn = ImageLearner(device=self.dev, use_cuda_graph=not self.args.no_graph, max_points=1024)
        perm = np.random.default_rng(0).permutation(256 * 256)
        Xs = [torch.from_numpy(coords[perm[k * 1024:(k + 1) * 1024]]).to(self.dev) for k in range(64)]
        ys = [torch.from_numpy(img[perm[k * 1024:(k + 1) * 1024]]).to(self.dev) for k in range(64)]
        loss0 = float(lrn.step(Xs[0], ys[0]).item())
        ms, _ = self.timed(lambda i: lrn.step(Xs[i % 64], ys[i % 64]), steps, warmup, collective=False)
        loss1 = float(lrn.step(Xs[0], ys[0]).item())
        tfl = 1024 * FLOP_PER_PIXEL_C1 / (ms * 1e-3) / 1e12
        return {"workload": "C1 image learning: SinusoidalEncoding(2,10,0,8) of integer pixel coords -> NeRF(40->8x256->3) -> "
                            "MSE -> Adam, 1024 pixels/step of a synthetic 256x256 image, CUDA-graph replayed",
                "metric": "us_per_step", "value": ms * 1e3, "unit": "us/step", "higher_is_better": False, "steps": steps,
                "ms_per_step": ms, "pixels_per_s": 1024 / (ms * 1e-3), "algorithmic_tflops": tfl,
                "tensor_frac_of_sustained": tfl / self.peaks["bf16_tflops_sustained"],
                "bound": "launch/latency (8 tiles of 128 pixels on 148 SMs)", "loss_first": loss0, "loss_last": loss1}

    def run_c2(self, steps, warmup):
        torch = self.torch
        from nerf_meets_mlx_b200.models.NeRF import default_args
        from nerf_meets_mlx_b200.training import NeRFTrainer
        B = 4096
        tr = NeRFTrainer(default_args(N_importance=0, n_depth_samples=N_SAMPLES), device=self.dev, max_rays=B,
                         use_cuda_graph=not self.args.no_graph, data_parallel=False)
        res = []
        for k in range(4):
            o, d, t = synth_batch(B, 500 + k)
            res.append(tuple(torch.from_numpy(a).to(self.dev) for a in (o, d, t)))
        ms, _ = self.timed(lambda i: tr.train_iteration(*res[i % 4]), steps, warmup, tr, collective=False)
        tfl = B * FLOP_PER_RAY_C2 / (ms * 1e-3) / 1e12
        tr.close()
        return {"workload": "C2 coarse NeRF training step: 64 stratified samples/ray, 8x256 MLP with view-dir head, 4096-ray "
                            "batch, one optimiser step (N_importance = 0), CUDA-graph replayed",
                "metric": "train_rays_per_s_coarse_64", "value": B / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
                "steps": steps, "flop_per_ray": FLOP_PER_RAY_C2, "algorithmic_tflops": tfl,
                "tensor_frac_of_sustained": tfl / self.peaks["bf16_tflops_sustained"]}

    def run_c4(self, steps, warmup):
        torch = self.torch
        from nerf_meets_mlx_b200 import ops
        from nerf_meets_mlx_b200.learners import HashGridLearner
        P = 262144
        lrn = HashGridLearner(device=self.dev, use_cuda_graph=not self.args.no_graph, max_points=P)
        g = torch.Generator(device=self.dev).manual_seed(0)
        xs = [torch.rand((P, 3), device=self.dev, generator=g) for _ in range(4)]
        ys = [torch.rand((P, 4), device=self.dev, generator=g) for _ in range(4)]
        ms, _ = self.timed(lambda i: lrn.step(xs[i % 4], ys[i % 4]), steps, warmup, collective=False)
        # phase times (eager, CUDA events): hash fwd | tiny MLP fwd + loss + bwd | table scatter | Adam (MLP + 64 MiB tables)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        x, y = xs[0], ys[0]
        tables = lrn.enc.hash_table.data
        acc = np.zeros(4)
        for rep in range(12):
            marks[0].record()
            feat = ops.hashgrid_fwd(x, tables, lrn.enc.scaled_res, lrn.log2_T)
            marks[1].record()
            pred = lrn.model._fwd_raw(0, feat, None, None, P, 1, save=True)
            loss, d_pred = ops.mse_fwd_bwd(pred, y)
            _, d_feat = lrn.model._bwd_raw(d_pred, P, out=lrn._g, want_input_grad=True)
            marks[2].record()
            d_tab = ops.hashgrid_bwd(x, lrn.enc.scaled_res, d_feat, lrn.L, lrn.F, lrn.log2_T)
            marks[3].record()
            lrn.opt_mlp.update(lrn.model, lrn._g)
            ops.adam_step(tables.view(-1), d_tab.view(-1), lrn.tab_m.view(-1), lrn.tab_v.view(-1), lrn.lr, 0.9, 0.99, 1e-8)
            marks[4].record()
            torch.cuda.synchronize()
            if rep > 1:
                acc += np.array([marks[k].elapsed_time(marks[k + 1]) for k in range(4)])
        ph = acc / 10
        hbm = self.peaks["hbm_gbs"]
        adam_bytes = tables.numel() * 4 * 7  # p, g, m, v read + p, m, v written
        return {"workload": "C4 hash grid (L=16, T=2^19, F=2, 16..2048) + tiny MLP (2x64 -> 4) training step: encode, MLP fwd, "
                            "MSE, MLP bwd incl. input gradient, table scatter (vector atomics), Adam on MLP and 64 MiB tables; "
                            "262144 points, CUDA-graph replayed",
                "metric": "train_points_per_s_hashgrid", "value": P / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms,
                "steps": steps,
                "phases_ms_eager": {"hash_fwd": ph[0], "mlp_fwd_loss_bwd": ph[1], "hash_bwd_scatter": ph[2],
                                    "adam_mlp_and_tables": ph[3]},
                "hash_fwd_frac_of_hbm": P * 1164 / (ph[0] * 1e-3) / 1e9 / hbm,
                "hash_bwd_frac_of_hbm": P * 2188 / (ph[2] * 1e-3) / 1e9 / hbm,
                "adam_tables_frac_of_hbm": adam_bytes / (ph[3] * 1e-3) / 1e9 / hbm,
                "step_hash_bytes_frac_of_hbm": P * HASH_BYTES_PT / (ms * 1e-3) / 1e9 / hbm,
                "note": "the width-64 tiny MLP runs fully fused (csrc/nmx_tiny.cu: one forward and one backward launch, weight "
                        "gradients in-kernel); NMX_DISABLE_TINY=1 selects the per-layer tcgen05 GEMM path"}

    def run_configs(self):
        out = {}
        for name, fn, st, wu in (("C1", self.run_c1, 200, 20), ("C2", self.run_c2, 20, 5), ("C4", self.run_c4, 20, 5)):
            try:
                out[name] = fn(st, wu)
            except Exception as e:  # a secondary config must not take the headline line down
                out[name] = {"error": f"{type(e).__name__}: {e}"}
        return out


def chain_traffic():
    """dram__bytes_{read,write}.sum per chain launch from this round's ncu --set full capture of the CURRENT build."""
    tpath = os.path.join(ROOT, "profiles", "r2_chain_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            return json.load(f)
    return {}


def run_ours(args):
    b = Bench(args)
    dist = b.dist
    peaks, peak_src = b.peaks, b.peak_src
    if args.config in ("C1", "C2", "C4"):
        if b.rank == 0:
            wu = max(args.warmup, 3)
            r = {"C1": b.run_c1, "C2": b.run_c2, "C4": b.run_c4}[args.config](args.steps, wu)
            line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": 1, "steps": args.steps,
                    "warmup": wu, "ms_per_step": r["ms_per_step"], "higher_is_better": r.get("higher_is_better", True),
                    "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                    "config": {"workload": r["workload"]}, "detail": r}
            print(json.dumps(line))
        if b.world > 1:
            dist.destroy_process_group()
        return 0

    c3 = b.run_c3()
    tr, B, world = b.tr, c3["B"], b.world
    render = None if args.no_render else b.run_render(tr)
    dp = b.run_dp_check(tr) if world > 1 else None
    configs = b.run_configs() if (b.rank == 0 and world == 1 and not args.no_configs) else None

    if b.rank == 0:
        ms_step, ms_e2e, prof, prof_steps = c3["ms_step"], c3["ms_e2e"], c3["prof"], c3["prof_steps"]
        rays_total = B * world
        value = rays_total / (ms_step * 1e-3)
        e2e = rays_total / (ms_e2e * 1e-3)
        # dominant kernel: the fused MLP chain (forward inference / forward training / backward data gradients);
        # ALGORITHMIC flops (fwd 593408 MAC/pt, dgrad 557696 MAC/pt) over the measured launch time
        P_c, P_f = B * N_SAMPLES, B * (N_SAMPLES + N_IMPORTANCE)
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
        # wgrad: HBM-bound (reads dY[P,M] and X[P,N] bf16 once per pass).  bf16 bytes per point: layer 0 (dY 512 + PE(pos)
        # 128), six plain layers (512 + 512), the skip layer (dY 512 + h 512 + PE(pos) 128), the dir layer (d_hd 256 +
        # h 512 + PE(dir) 128).  Part of it is served by L2 (dY was written by the chain just before).
        wg_bytes = (640 + 7 * 1024 + 1152 + 896) * (P_c + P_f) * prof_steps
        wg_ms = prof[1][0]
        traffic = chain_traffic() if (world == 1 and B == RAYS_PER_GPU) else {}
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "C3 coarse+fine NeRF training step (reference iteration semantics), 64 stratified + "
                                   "128 importance samples/ray, 8x256 MLPs with view-dir head",
                       "rays_per_gpu": B, "global_rays": rays_total, "parallelism": f"dp{world} ray-sharded",
                       "gradient_exchange": ("nmx_allreduce_adam over NVLink peer memory, inside the CUDA graph"
                                             if tr.xchg is not None else ("nccl all-reduce" if world > 1 else "none")),
                       "cuda_graph": bool(tr._graph is not None),
                       "l2": "per-step working set (~20 GB of saved activations and data gradients) >> 126 MB L2; "
                             "4 rotating input batches"},
            "clocks": c3["clocks"],
            "e2e": {"value": e2e, "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(B * 9 * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(c3["launches"]),
            "step_tensor_frac": {"algorithmic_tflops": value * FLOP_PER_RAY / 1e12 / world,
                                 "peak_tflops_sustained": peaks["bf16_tflops_sustained"],
                                 "frac": value * FLOP_PER_RAY / 1e12 / world / peaks["bf16_tflops_sustained"],
                                 "frac_of_burst": value * FLOP_PER_RAY / 1e12 / world / peaks["bf16_tflops"],
                                 "flop_per_ray": FLOP_PER_RAY},
            "roofline": {"bound": "tensor",
                         "kernel": "mlp_chain_kernel (fused whole-MLP forward / backward chains, tcgen05 + TMEM)",
                         "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops_sustained"],
                         "peak_source": peak_src + " sustained (timed inside a long step)",
                         "launches": int(chain_n), "avg_launch_ms": chain_ms / max(chain_n, 1),
                         "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
                         "modes": sub, "profiled_steps": prof_steps,
                         "wgrad": {"bound": "hbm", "achieved": wg_bytes / (wg_ms * 1e-3) / 1e9 if wg_ms > 0 else 0.0,
                                   "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": (wg_bytes / (wg_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if wg_ms > 0 else 0.0,
                                   "tflops": (prof[1][1] / (wg_ms * 1e-3) / 1e12) if wg_ms > 0 else 0.0,
                                   "launches": int(prof[1][2]), "ms_per_step": wg_ms / prof_steps}},
            "strong": c3["strong"],
            "render": render,
        }
        if dp is not None:
            line["dp_check"] = dp
        if configs is not None:
            line["configs"] = configs
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4"],
                    help="C3 (default): the headline line with strong / render / configs blocks; C1, C2, C4: that config alone")
    ap.add_argument("--rays-per-gpu", type=int, default=0, help="override the 8192 rays per GPU of the headline step")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the training iteration eagerly (no CUDA graph)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
