"""The reference's training iteration (mlx_nerf/entrypoints/__test_nerf.py:47-145, 200-305) driven straight through
the C ABI ops -- no autograd graph, no host round trip for the resampling, optional ray-sharded data parallelism
 with the gradient exchange fused into the optimiser kernel over NVLink peer memory (`nmx_allreduce_adam`),
or one NCCL all-reduce of the flat fp32 gradient per optimiser step as the fallback."""
import math
import os

import torch
import torch.distributed as dist

from . import ops
from .models.NeRF import AdamMLX, create_NeRF, default_args


def assemble_rays(rays_o, rays_d, near, far):
    """[o, d, near, far, viewdirs] with viewdirs = d/||d|| (__test_nerf.py:57-82), one kernel (nmx_assemble_rays)."""
    return ops.assemble_rays(rays_o, rays_d, near, far)


class NeRFTrainer:
    """Coarse(+fine) NeRF trainer with the reference's semantics:
      * coarse loss on render_rays (white_bkgd from args), fine loss on raw2outputs with white_bkgd=False (quirk);
      * importance samples are a DETACHED input of the fine step and come from the coarse net AFTER its update;
      * one Adam instance without bias correction updates both nets and (faithful default) shares its moments;
      * lr = lrate * 0.1 ** (i / (lrate_decay * 1000)).
    `reuse_coarse_forward=True` is a declared deviation that resamples from the in-step coarse forward (pre-update
    weights) and skips the re-forward.

    Data parallel (`process_group` / an initialised default group): every rank runs the iteration on its own shard of
    the step's rays; the two gradient exchanges + Adam updates are single kernels over peer memory (`use_p2p`, default
    on CUDA) so that the WHOLE iteration, collectives included, is one CUDA graph.  `use_p2p=False` (or the
    environment variable NMX_DP_NCCL=1) selects the NCCL path: three graphs cut at the two eager all-reduces.
    `data_parallel=False` keeps the trainer single-process even inside an initialised process group.

    With `use_cuda_graph=True` the returned dict holds STATIC device buffers that the next iteration overwrites; clone
    what must outlive the call."""

    def __init__(self, args=None, device="cuda", near=2.0, far=6.0, shared_adam_state=True,
                 reuse_coarse_forward=False, process_group=None, max_rays=8192, use_cuda_graph=False, use_p2p=None,
                 data_parallel=True):
        self.args = args if args is not None else default_args(N_importance=128)
        self.device = torch.device(device)
        self.near, self.far = float(near), float(far)
        kw_train, _, _, opt = create_NeRF(self.args, device=device)
        self.kw = kw_train
        self.coarse = kw_train["network_coarse"]
        self.fine = kw_train["network_fine"]
        self.optimizer = AdamMLX(self.args.lrate, betas=(0.9, 0.999), shared_state=shared_adam_state)
        self.n_samples = int(self.args.n_depth_samples)
        self.n_importance = int(self.args.N_importance)
        self.white_bkgd = bool(self.args.white_bkgd)
        self.lindisp = bool(getattr(self.args, "lindisp", False))
        self.reuse_coarse_forward = reuse_coarse_forward
        self.iteration = 0
        # CUDA-graph replay of the iteration: the ~45 launches of a step are captured once per batch shape and replayed,
        # removing the launch gaps; the learning rate lives in a device scalar so replays follow the decay schedule.
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graphs = {}      # key -> (graph or [graphs], io)
        self._graph = None     # the entry used by the last graphed iteration (bench.py reads it)
        self._lr_dev = None
        self.graph_launches = 0  # libnmx kernel launches inside one captured iteration
        self.pg = process_group
        in_group = data_parallel and dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if in_group else 1
        self.xchg = None
        if use_p2p is None:
            use_p2p = self.device.type == "cuda" and os.environ.get("NMX_DP_NCCL", "0") != "1"
        if self.world > 1 and use_p2p:
            if self.fine is not None and self.fine.flat.numel() != self.coarse.flat.numel():
                raise NotImplementedError("peer exchange needs coarse and fine nets of the same size")
            from .parallel import PeerExchange
            self.xchg = PeerExchange(self.coarse.flat.numel(), n_bufs=2, device=self.device, group=process_group)
            self._g_coarse = self.xchg.buffer(0)
            self._g_fine = self.xchg.buffer(1) if self.fine is not None else None
        else:
            self._g_coarse = torch.empty_like(self.coarse.flat.data)
            self._g_fine = torch.empty_like(self.fine.flat.data) if self.fine is not None else None
        self.coarse.reserve(max_rays * self.n_samples, training=True)
        if self.fine is not None:
            self.fine.reserve(max_rays * (self.n_samples + self.n_importance), training=True)
        if self.world > 1:
            self.broadcast_parameters()

    # ------------------------------------------------------------------ data parallel plumbing
    def broadcast_parameters(self):
        for m in (self.coarse, self.fine):
            if m is not None:
                dist.broadcast(m.flat.data, src=0, group=self.pg)
                m.mark_params_updated()

    def _allreduce_mean(self, g):
        if self.world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
            g.mul_(1.0 / self.world)

    def parameter_checksums(self):
        """(sum, sum of squares) of each net's fp32 parameters in float64 -- identical on every rank of a healthy
        data-parallel run (bench.py's dp_check)."""
        out = []
        for m in (self.coarse, self.fine):
            if m is not None:
                p = m.flat.data.double()
                out += [float(p.sum().item()), float((p * p).sum().item())]
        return out

    def close(self):
        """Release CUDA graphs and the peer mappings (call before destroy_process_group)."""
        self._graphs.clear()
        self._graph = None
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None

    # ------------------------------------------------------------------ one optimiser step on one net
    def _grad(self, model, rays, z, target, white_bkgd, g_buf):
        """Forward + backward of one net on this rank's rays: local gradient into g_buf, loss and compositing weights."""
        B, n = z.shape
        rays_d = rays[:, 3:6]  # strided view: the compositing kernels take the row stride
        raw = model._fwd_raw(1, rays, z, None, B, n, save=True)
        # compositing forward + MSE + compositing backward in one kernel; the weights are only needed when the fine
        # pass resamples from THIS forward (reuse_coarse_forward, a declared deviation)
        loss, d_raw, weights, _ = ops.composite_loss_fwd_bwd(raw.view(B, n, 4), z, rays_d, target, white_bkgd=white_bkgd,
                                                             want_weights=self.reuse_coarse_forward)
        model._bwd_raw(d_raw.view(B * n, 4), B * n, out=g_buf)
        return loss, weights

    def _update(self, model, buf_idx, g_buf, lr_dev):
        """Gradient exchange + optimiser step (__test_nerf.py:134,144)."""
        if self.xchg is not None:
            self.optimizer.update_exchange(model, self.xchg, buf_idx, lr_dev=lr_dev)
        else:
            self._allreduce_mean(g_buf)
            self.optimizer.update(model, g_buf, lr_dev=lr_dev)

    def train_iteration(self, rays_o, rays_d, target, u_vals=None):
        """One pass of the reference loop body on this rank's shard of rays.  Returns device scalars."""
        if self.use_cuda_graph and (self.world == 1 or self.xchg is not None):
            return self._train_iteration_graphed(rays_o, rays_d, target, u_vals)
        if self.use_cuda_graph:
            return self._train_iteration_graphed_nccl(rays_o, rays_d, target, u_vals)
        self.iteration += 1
        out = self._iteration_body(rays_o, rays_d, target, u_vals, lr_dev=None)
        self._advance_lr()
        return out

    def train_iteration_pixels(self, H, W, K, c2w, pix, image, u_vals=None):
        """The reference's per-iteration ray selection (__test_nerf.py:208-236) on the device: pixel ids (row*W + col)
        of one training view -> rays and target pixels in one kernel (nmx_gen_rays), then `train_iteration`.
        `image` is the view's [H*W, C>=3] fp32 pixels already resident in HBM; no host round trip."""
        rays, target = ops.gen_rays(H, W, K, c2w, pix, self.near, self.far, 6, image=image)
        return self.train_iteration(rays[:, 0:3], rays[:, 3:6], target, u_vals)

    def _advance_lr(self):
        # learning-rate decay (__test_nerf.py:302-305)
        decay_steps = self.args.lrate_decay * 1000
        self.optimizer.learning_rate = self.args.lrate * (0.1 ** (self.iteration / decay_steps))

    def _refresh_lr_dev(self, device):
        if self._lr_dev is None:
            self._lr_dev = torch.empty((), dtype=torch.float32, device=device)
        self._lr_dev.fill_(float(self.optimizer.learning_rate))

    def _static_io(self, key, rays_o, rays_d, target):
        B = rays_o.shape[0]
        return {"key": key, "o": rays_o.clone(), "d": rays_d.clone(), "t": target.clone(),
                "u": torch.empty((B, max(self.n_importance, 1)), dtype=torch.float32, device=rays_o.device)}

    def _load_io(self, io, rays_o, rays_d, target, u_vals):
        io["o"].copy_(rays_o); io["d"].copy_(rays_d); io["t"].copy_(target)
        if self.fine is not None:
            if u_vals is not None:
                io["u"].copy_(u_vals)
            else:
                io["u"].uniform_()  # the draw sample_from_inverse_cdf_torch makes (sampling/__init__.py:133)

    def _train_iteration_graphed(self, rays_o, rays_d, target, u_vals):
        """Single graph per batch shape: single GPU, or data parallel with the in-kernel peer exchange."""
        from . import _lib_loader as L
        if self.iteration == 0:  # the very first iteration runs eagerly: warm-up of every kernel (and a real step)
            self.iteration += 1
            out = self._iteration_body(rays_o, rays_d, target, u_vals, lr_dev=None)
            self._advance_lr()
            return out
        self._refresh_lr_dev(rays_o.device)
        key = ("g", tuple(rays_o.shape))
        if key not in self._graphs:
            io = self._static_io(key, rays_o, rays_d, target)
            self._load_io(io, rays_o, rays_d, target, u_vals)
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.pg)  # every rank captures (and later replays) the same number of exchanges
            g = torch.cuda.CUDAGraph()
            n0, steps0 = L.launch_count(), self.optimizer.step_count
            with torch.cuda.graph(g):
                io["out"] = self._iteration_body(io["o"], io["d"], io["t"], io["u"], lr_dev=self._lr_dev)
            self.optimizer.step_count = steps0  # capture launches nothing: the replays count the steps
            io["launches"] = L.launch_count() - n0
            self._graphs[key] = (g, io)
        g, io = self._graphs[key]
        self._graph, self.graph_launches = g, io["launches"]
        self._load_io(io, rays_o, rays_d, target, u_vals)
        self.iteration += 1
        self.optimizer.step_count += 2 if self.fine is not None else 1
        g.replay()
        self._advance_lr()
        return io["out"]

    def _train_iteration_graphed_nccl(self, rays_o, rays_d, target, u_vals):
        """Data-parallel graph replay on the NCCL path: the iteration is cut at its two gradient all-reduces into three
        CUDA graphs that share one memory pool (coarse forward/backward | Adam + coarse re-forward + resampling + fine
        forward/backward | Adam); the NCCL all-reduces run eagerly between the replays, so no collective is captured."""
        from . import _lib_loader as L
        if self.iteration == 0:  # first iteration eagerly: warm-up of the kernels and of the NCCL communicator
            self.iteration += 1
            out = self._iteration_body(rays_o, rays_d, target, u_vals, lr_dev=None)
            self._advance_lr()
            return out
        self._refresh_lr_dev(rays_o.device)
        key = ("nccl", tuple(rays_o.shape))
        if key not in self._graphs:
            io = self._static_io(key, rays_o, rays_d, target)
            inv_world = 1.0 / self.world
            st = {}

            def seg_a():
                rays = assemble_rays(io["o"], io["d"], self.near, self.far)
                z = ops.sample_z_rays(rays, self.n_samples, lindisp=self.lindisp)
                loss_c, weights = self._grad(self.coarse, rays, z, io["t"], self.white_bkgd, self._g_coarse)
                st.update(rays=rays, z=z, weights=weights, out={"loss_coarse": loss_c})

            def seg_b():
                self._g_coarse.mul_(inv_world)
                self.optimizer.update(self.coarse, self._g_coarse, lr_dev=self._lr_dev)
                if self.fine is None:
                    return
                rays, z, weights = st["rays"], st["z"], st["weights"]
                z_fine = self._resample(rays, z, weights, io["u"])
                loss_f, _ = self._grad(self.fine, rays, z_fine, io["t"], False, self._g_fine)
                st["out"]["loss_fine"] = loss_f
                st["out"]["z_fine"] = z_fine

            def seg_c():
                self._g_fine.mul_(inv_world)
                self.optimizer.update(self.fine, self._g_fine, lr_dev=self._lr_dev)

            segs = [seg_a, seg_b] + ([seg_c] if self.fine is not None else [])
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            graphs = []
            n0, steps0 = L.launch_count(), self.optimizer.step_count
            for seg in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    seg()
                graphs.append(g)
            self.optimizer.step_count = steps0
            io["launches"] = L.launch_count() - n0
            io["out"], io["state"] = st["out"], st  # keeps the tensors handed from one graph to the next alive
            self._graphs[key] = (graphs, io)
        graphs, io = self._graphs[key]
        self._graph, self.graph_launches = graphs, io["launches"]
        self._load_io(io, rays_o, rays_d, target, u_vals)
        self.iteration += 1
        self.optimizer.step_count += 2 if self.fine is not None else 1
        graphs[0].replay()
        dist.all_reduce(self._g_coarse, op=dist.ReduceOp.SUM, group=self.pg)
        graphs[1].replay()
        if self.fine is not None:
            dist.all_reduce(self._g_fine, op=dist.ReduceOp.SUM, group=self.pg)
            graphs[2].replay()
        self._advance_lr()
        return io["out"]

    def _resample(self, rays, z, weights, u_vals):
        """Importance depths for the fine step from the UPDATED coarse net (__test_nerf.py:270-288): render_rays again,
        inverse-CDF sampling, sort-merge with the coarse depths."""
        B = rays.shape[0]
        if not self.reuse_coarse_forward:
            raw = self.coarse._fwd_raw(1, rays, z, None, B, self.n_samples, save=False)
            _, _, _, weights, _ = ops.composite_fwd(raw.view(B, self.n_samples, 4), z, rays[:, 3:6],
                                                    white_bkgd=self.white_bkgd)
        if u_vals is None:
            u_vals = torch.rand((B, self.n_importance), device=rays.device)
        return ops.sample_pdf(z, weights, u_vals, want_imp=False)["z_merged"]

    def _iteration_body(self, rays_o, rays_d, target, u_vals=None, lr_dev=None):
        rays = assemble_rays(rays_o, rays_d, self.near, self.far)
        z = ops.sample_z_rays(rays, self.n_samples, lindisp=self.lindisp)
        loss_c, weights = self._grad(self.coarse, rays, z, target, self.white_bkgd, self._g_coarse)
        self._update(self.coarse, 0, self._g_coarse, lr_dev)
        out = {"loss_coarse": loss_c}
        if self.fine is not None:
            z_fine = self._resample(rays, z, weights, u_vals)
            loss_f, _ = self._grad(self.fine, rays, z_fine, target, False, self._g_fine)  # white_bkgd=False (:106)
            self._update(self.fine, 1, self._g_fine, lr_dev)
            out["loss_fine"] = loss_f
            out["z_fine"] = z_fine
        return out

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def render_rays_eval(self, rays, u_vals=None):
        """Coarse + fine evaluation of a ray batch (render_rays_eval, rendering/render.py:164-241) through the raw ops."""
        B = rays.shape[0]
        rays_d = rays[:, 3:6]
        z = ops.sample_z_rays(rays, self.n_samples)
        raw = self.coarse._fwd_raw(1, rays, z, None, B, self.n_samples, save=False)
        rgb_c, _, _, weights, _ = ops.composite_fwd(raw.view(B, self.n_samples, 4), z, rays_d, white_bkgd=self.white_bkgd)
        if u_vals is None:
            u_vals = torch.rand((B, self.n_importance), device=rays.device)
        z_fine = ops.sample_pdf(z, weights, u_vals, want_imp=False)["z_merged"]
        net = self.fine if self.fine is not None else self.coarse
        n2 = z_fine.shape[1]
        raw_f = net._fwd_raw(1, rays, z_fine, None, B, n2, save=False)
        rgb, disp, acc, _, _ = ops.composite_fwd(raw_f.view(B, n2, 4), z_fine, rays_d, white_bkgd=self.white_bkgd)
        return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "rgb_coarse": rgb_c}

    @torch.no_grad()
    def render_frame(self, rays, chunk=1024 * 32, u_vals=None, process_group=None):
        """Full-frame coarse+fine render of a ray buffer [n, 11] in chunks of `chunk` rays (batchify_rays,
        rendering/render.py:243-266).  With `process_group` the buffer is sharded into one contiguous ray tile per
        rank and rgb / disp / acc are all_gathered into the assembled frame."""
        n = rays.shape[0]
        start, stop = 0, n
        if process_group is not None:
            from .parallel import gather_tiles, ray_tile
            start, stop = ray_tile(n, dist.get_rank(process_group), dist.get_world_size(process_group))
        outs = {"rgb_map": [], "disp_map": [], "acc_map": []}
        for s in range(start, stop, chunk):
            e = min(s + chunk, stop)
            r = self.render_rays_eval(rays[s:e], None if u_vals is None else u_vals[s:e])
            for k in outs:
                outs[k].append(r[k])
        res = {k: (torch.cat(v) if v else torch.empty((0, 3 if k == "rgb_map" else 1), device=rays.device))
               for k, v in outs.items()}
        if process_group is not None:
            res = {k: gather_tiles(v, n, process_group) for k, v in res.items()}
        return res


def psnr(mse):
    """ops/metric.py:16-18."""
    return 10.0 * math.log10(1.0 / float(mse))
