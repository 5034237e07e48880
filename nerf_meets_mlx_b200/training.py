"""The reference's training iteration (mlx_nerf/entrypoints/__test_nerf.py:47-145, 200-305) driven straight through
the C ABI ops -- no autograd graph, no host round trip for the resampling, optional ray-sharded data parallelism
with one NCCL all-reduce of the flat fp32 gradient per optimiser step."""
import math

import torch
import torch.distributed as dist

from . import ops
from .models.NeRF import AdamMLX, create_NeRF, default_args


def assemble_rays(rays_o, rays_d, near, far):
    """[o, d, near, far, viewdirs] with viewdirs = d/||d|| (__test_nerf.py:57-82)."""
    viewdirs = rays_d / torch.linalg.norm(rays_d, dim=-1, keepdim=True)
    ones = torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, near * ones, far * ones, viewdirs], dim=-1).contiguous()


class NeRFTrainer:
    """Coarse(+fine) NeRF trainer with the reference's semantics:
      * coarse loss on render_rays (white_bkgd from args), fine loss on raw2outputs with white_bkgd=False (quirk);
      * importance samples are a DETACHED input of the fine step and come from the coarse net AFTER its update;
      * one Adam instance without bias correction updates both nets and (faithful default) shares its moments;
      * lr = lrate * 0.1 ** (i / (lrate_decay * 1000)).
    `reuse_coarse_forward=True` is a declared deviation that resamples from the in-step coarse forward (pre-update
    weights) and skips the re-forward."""

    def __init__(self, args=None, device="cuda", near=2.0, far=6.0, shared_adam_state=True,
                 reuse_coarse_forward=False, process_group=None, max_rays=8192, use_cuda_graph=False):
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
        self.reuse_coarse_forward = reuse_coarse_forward
        self.iteration = 0
        # CUDA-graph replay of the iteration (single GPU): the ~70 launches of a step are captured once and replayed,
        # removing the launch gaps; the learning rate lives in a device scalar so replays follow the decay schedule.
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph = None
        self._graph_io = None
        self._lr_dev = None
        self.graph_launches = 0  # libnmx kernel launches inside one captured iteration
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
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

    # ------------------------------------------------------------------ one optimiser step on one net
    def _grad(self, model, rays, z, target, white_bkgd, g_buf):
        """Forward + backward of one net on this rank's rays: local gradient into g_buf, loss and compositing weights."""
        B, n = z.shape
        raw = model._fwd_raw(1, rays, z, None, B, n, save=True)
        rgb, _, _, weights, _ = ops.composite_fwd(raw.view(B, n, 4), z, rays[:, 3:6].contiguous(), white_bkgd=white_bkgd)
        loss, d_rgb = ops.mse_fwd_bwd(rgb, target)
        d_raw = ops.composite_bwd(raw.view(B, n, 4), z, rays[:, 3:6].contiguous(), d_rgb, white_bkgd=white_bkgd)
        model._bwd_raw(d_raw.view(B * n, 4), B * n, out=g_buf)
        return loss, weights

    def _step(self, model, rays, z, target, white_bkgd, g_buf):
        loss, weights = self._grad(model, rays, z, target, white_bkgd, g_buf)
        self._allreduce_mean(g_buf)
        self.optimizer.update(model, g_buf, lr_dev=self._lr_dev)
        return loss, weights

    def train_iteration(self, rays_o, rays_d, target, u_vals=None):
        """One pass of the reference loop body on this rank's shard of rays.  Returns device scalars."""
        if self.use_cuda_graph and self.world == 1:
            return self._train_iteration_graphed(rays_o, rays_d, target, u_vals)
        if self.use_cuda_graph and u_vals is not None:
            return self._train_iteration_graphed_dp(rays_o, rays_d, target, u_vals)
        self.iteration += 1
        out = self._iteration_body(rays_o, rays_d, target, u_vals)
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

    def _train_iteration_graphed(self, rays_o, rays_d, target, u_vals):
        from . import _lib_loader as L
        if self._lr_dev is None:
            self._lr_dev = torch.empty((), dtype=torch.float32, device=rays_o.device)
        self._lr_dev.fill_(float(self.optimizer.learning_rate))
        key = (tuple(rays_o.shape), u_vals is not None)
        if self._graph is None or self._graph_io["key"] != key:
            if self.iteration == 0:  # the very first iteration runs eagerly: it is the warm-up (and a real step)
                self.iteration += 1
                out = self._iteration_body(rays_o, rays_d, target, u_vals)
                self._advance_lr()
                return out
            io = {"key": key, "o": rays_o.clone(), "d": rays_d.clone(), "t": target.clone(),
                  "u": u_vals.clone() if u_vals is not None else None}
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g):
                io["out"] = self._iteration_body(io["o"], io["d"], io["t"], io["u"])
            self.graph_launches = L.launch_count() - n0
            self._graph, self._graph_io = g, io
        io = self._graph_io
        io["o"].copy_(rays_o); io["d"].copy_(rays_d); io["t"].copy_(target)
        if u_vals is not None:
            io["u"].copy_(u_vals)
        self.iteration += 1
        self.optimizer.step_count += 2 if self.fine is not None else 1
        self._graph.replay()
        self._advance_lr()
        return io["out"]

    def _train_iteration_graphed_dp(self, rays_o, rays_d, target, u_vals):
        """Data-parallel graph replay: the iteration is cut at its two gradient all-reduces into three CUDA graphs that
        share one memory pool (coarse forward/backward | Adam + coarse re-forward + resampling + fine forward/backward |
        Adam); the NCCL all-reduces run eagerly between the replays, so no collective is ever captured."""
        from . import _lib_loader as L
        if self._lr_dev is None:
            self._lr_dev = torch.empty((), dtype=torch.float32, device=rays_o.device)
        self._lr_dev.fill_(float(self.optimizer.learning_rate))
        key = ("dp", tuple(rays_o.shape))
        if self._graph is None or self._graph_io["key"] != key:
            if self.iteration == 0:  # first iteration eagerly: warm-up of the kernels and of the NCCL communicator
                self.iteration += 1
                out = self._iteration_body(rays_o, rays_d, target, u_vals)
                self._advance_lr()
                return out
            io = {"key": key, "o": rays_o.clone(), "d": rays_d.clone(), "t": target.clone(), "u": u_vals.clone()}
            inv_world = 1.0 / self.world
            st = {}

            def seg_a():
                rays = assemble_rays(io["o"], io["d"], self.near, self.far)
                z = ops.sample_z(rays[:, 6], rays[:, 7], self.n_samples, lindisp=bool(getattr(self.args, "lindisp", False)))
                loss_c, weights = self._grad(self.coarse, rays, z, io["t"], self.white_bkgd, self._g_coarse)
                st.update(rays=rays, z=z, weights=weights, out={"loss_coarse": loss_c})

            def seg_b():
                self._g_coarse.mul_(inv_world)
                self.optimizer.update(self.coarse, self._g_coarse, lr_dev=self._lr_dev)
                if self.fine is None:
                    return
                rays, z, weights = st["rays"], st["z"], st["weights"]
                B = rays.shape[0]
                if not self.reuse_coarse_forward:
                    raw = self.coarse._fwd_raw(1, rays, z, None, B, self.n_samples, save=False)
                    _, _, _, weights, _ = ops.composite_fwd(raw.view(B, self.n_samples, 4), z, rays[:, 3:6].contiguous(),
                                                            white_bkgd=self.white_bkgd)
                z_fine = ops.sample_pdf(z, weights, io["u"], want_imp=False)["z_merged"]
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
            n0 = L.launch_count()
            for seg in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    seg()
                graphs.append(g)
            self.graph_launches = L.launch_count() - n0
            io["out"], io["state"] = st["out"], st  # keeps the tensors handed from one graph to the next alive
            self._graph, self._graph_io = graphs, io
        io = self._graph_io
        io["o"].copy_(rays_o); io["d"].copy_(rays_d); io["t"].copy_(target); io["u"].copy_(u_vals)
        self.iteration += 1
        self.optimizer.step_count += 2 if self.fine is not None else 1
        self._graph[0].replay()
        dist.all_reduce(self._g_coarse, op=dist.ReduceOp.SUM, group=self.pg)
        self._graph[1].replay()
        if self.fine is not None:
            dist.all_reduce(self._g_fine, op=dist.ReduceOp.SUM, group=self.pg)
            self._graph[2].replay()
        self._advance_lr()
        return io["out"]

    def _iteration_body(self, rays_o, rays_d, target, u_vals=None):
        rays = assemble_rays(rays_o, rays_d, self.near, self.far)
        B = rays.shape[0]
        rays_d_c = rays[:, 3:6].contiguous()
        z = ops.sample_z(rays[:, 6], rays[:, 7], self.n_samples, lindisp=bool(getattr(self.args, "lindisp", False)))
        loss_c, weights = self._step(self.coarse, rays, z, target, self.white_bkgd, self._g_coarse)
        out = {"loss_coarse": loss_c}
        if self.fine is not None:
            if not self.reuse_coarse_forward:
                # render_rays again with the UPDATED coarse net (__test_nerf.py:270)
                raw = self.coarse._fwd_raw(1, rays, z, None, B, self.n_samples, save=False)
                _, _, _, weights, _ = ops.composite_fwd(raw.view(B, self.n_samples, 4), z, rays_d_c, white_bkgd=self.white_bkgd)
            if u_vals is None:
                u_vals = torch.rand((B, self.n_importance), device=rays.device)
            z_fine = ops.sample_pdf(z, weights, u_vals, want_imp=False)["z_merged"]
            loss_f, _ = self._step(self.fine, rays, z_fine, target, False, self._g_fine)  # white_bkgd=False (:106)
            out["loss_fine"] = loss_f
            out["z_fine"] = z_fine
        return out

    # ------------------------------------------------------------------ inference
    @torch.no_grad()
    def render_rays_eval(self, rays, u_vals=None):
        """Coarse + fine evaluation of a ray batch (render_rays_eval, rendering/render.py:164-241) through the raw ops."""
        B = rays.shape[0]
        rays_d_c = rays[:, 3:6].contiguous()
        z = ops.sample_z(rays[:, 6], rays[:, 7], self.n_samples)
        raw = self.coarse._fwd_raw(1, rays, z, None, B, self.n_samples, save=False)
        rgb_c, _, _, weights, _ = ops.composite_fwd(raw.view(B, self.n_samples, 4), z, rays_d_c, white_bkgd=self.white_bkgd)
        if u_vals is None:
            u_vals = torch.rand((B, self.n_importance), device=rays.device)
        z_fine = ops.sample_pdf(z, weights, u_vals, want_imp=False)["z_merged"]
        net = self.fine if self.fine is not None else self.coarse
        n2 = z_fine.shape[1]
        raw_f = net._fwd_raw(1, rays, z_fine, None, B, n2, save=False)
        rgb, disp, acc, _, _ = ops.composite_fwd(raw_f.view(B, n2, 4), z_fine, rays_d_c, white_bkgd=self.white_bkgd)
        return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "rgb_coarse": rgb_c}


def psnr(mse):
    """ops/metric.py:16-18."""
    return 10.0 * math.log10(1.0 / float(mse))
