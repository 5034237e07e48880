"""Mirror of mlx_nerf/models/NeRF.py on CUDA: the NeRF MLP runs as bf16 tcgen05 GEMMs with fp32 accumulation
(libnmx, include/nmx.h `nmx_mlp_*`); parameters stay fp32 in the reference's tree layout."""
import ctypes
import math
from types import SimpleNamespace

import torch

from .. import _lib_loader as L
from . import embedding


class _Cfg(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int), ("width", ctypes.c_int), ("in_pos", ctypes.c_int),
                ("in_dir", ctypes.c_int), ("out_ch", ctypes.c_int), ("skip_layer", ctypes.c_int),
                ("use_viewdirs", ctypes.c_int), ("n_freqs_pos", ctypes.c_int), ("n_freqs_dir", ctypes.c_int)]


class _LinearView:
    """`.weight [out,in]` / `.bias [out]` views into the flat parameter buffer (reference tree leaf)."""

    def __init__(self, weight, bias):
        self.weight = weight
        self.bias = bias


class _MLPFunction(torch.autograd.Function):
    """Autograd node of one MLP pass.  The activations a backward needs live in the model's ONE workspace; every saving
    forward gets a generation id.  If a later saving forward on the same model has overwritten them by the time this
    node's backward runs (chunked callers: run_model's netchunk loop, batchify_rays, render_rays_eval with
    network_fine=None), the forward is RECOMPUTED from the node's own (small) inputs first -- never a silent gradient
    from the wrong activations."""

    @staticmethod
    def forward(ctx, flat, model, enc_kind, x_or_rays, z, bands, B, n):
        # grad mode is off inside Function.forward; an already-encoded input (enc_kind 0) may itself need a gradient:
        # a learnable encoder in front of the MLP (the hash grid) gets it as the reference's autograd would give it
        ctx.input_grad = bool(enc_kind == 0 and ctx.needs_input_grad[3])
        need_grad = bool(ctx.needs_input_grad[0]) or ctx.input_grad
        out = model._fwd_raw(enc_kind, x_or_rays, z, bands, B, n, save=need_grad)
        ctx.model = model
        ctx.P = B * n
        if need_grad:
            ctx.gen = model._save_gen
            ctx.save_for_backward(x_or_rays, z, bands)
            ctx.args = (enc_kind, B, n)
        return out

    @staticmethod
    def backward(ctx, d_out):
        model = ctx.model
        if ctx.gen != model._save_gen:  # a later forward reused the workspace: rebuild this pass's activations
            model.recomputed_backwards += 1
            x_or_rays, z, bands = ctx.saved_tensors
            enc_kind, B, n = ctx.args
            model._fwd_raw(enc_kind, x_or_rays.detach(), z, bands, B, n, save=True)
        if ctx.input_grad:
            g, d_x = model._bwd_raw(d_out.contiguous(), ctx.P, want_input_grad=True)
            return g, None, None, d_x, None, None, None, None
        g = model._bwd_raw(d_out.contiguous(), ctx.P)
        return g, None, None, None, None, None, None, None


class NeRF(torch.nn.Module):
    """NeRF (models/NeRF.py:160-243): n_layers x width ReLU trunk, skip concat [input_pos, h] after the listed layer,
    view-dir head (alpha, feature -> dir layer -> rgb; no sigmoid, no activation on feature) or `output_linear`.

    Parameters live in ONE flat fp32 buffer (`self.flat`, what the optimiser and the NCCL all-reduce see), exposed
    through the reference's names: list_linears_pos[i].weight/.bias, list_linears_dir[0], feature_linear,
    alpha_linear, rgb_linear / output_linear.  Init U(+-1/sqrt(in)) like MLX's nn.Linear.
    """

    def __init__(self, n_layers=8, width_layers=256, channel_input=3, channel_input_views=3, channel_output=4,
                 list_skip_connection_layers=[4], is_use_view_directions=False, device="cuda", seed=None,
                 n_freqs_pos=0, n_freqs_dir=0, max_points=1 << 16, max_save_points=1 << 21):
        super().__init__()
        self.D = n_layers
        self.W = width_layers
        self.channel_input_pos = channel_input
        self.channel_input_dir = channel_input_views
        self.channel_output = channel_output
        self.list_skip_connection_layers = list(list_skip_connection_layers)
        self.is_use_view_directions = bool(is_use_view_directions)
        skips = [s for s in self.list_skip_connection_layers if 0 <= s < n_layers - 1]
        if len(skips) > 1:
            raise NotImplementedError("at most one skip connection (the reference hard-codes [4], NeRF.py:67-68)")
        self._cfg = _Cfg(n_layers, width_layers, channel_input, channel_input_views if is_use_view_directions else 0,
                         channel_output, skips[0] if skips else -1, 1 if is_use_view_directions else 0,
                         n_freqs_pos, n_freqs_dir)
        n_params = int(L.lib().nmx_mlp_param_count(ctypes.byref(self._cfg)))
        dev = torch.device(device)
        self.flat = torch.nn.Parameter(torch.zeros(n_params, dtype=torch.float32, device=dev))
        self._build_views()
        self.reset_parameters(seed)
        self._plan = None
        self._ws = None
        self._ws_train = False
        self._max_points = int(max_points)
        self.max_save_points = int(max_save_points)  # largest differentiable pass kept in one workspace (~10 KB/point)
        self._packed_version = -1
        self._save_gen = 0              # generation id of the activations currently held in the workspace
        self.recomputed_backwards = 0   # backward passes that had to rebuild overwritten activations
        self.out_cols = 4 if self.is_use_view_directions else channel_output

    # ---------------------------------------------------------------- parameter tree
    def _layer_shapes(self):
        W, cin = self.W, self.channel_input_pos
        skip = self._cfg.skip_layer
        shapes = []
        for l in range(self.D):
            i = cin if l == 0 else (W + cin if (skip >= 0 and l == skip + 1) else W)
            shapes.append((f"list_linears_pos.{l}", W, i))
        if self.is_use_view_directions:
            shapes += [("feature_linear", W, W), ("alpha_linear", 1, W),
                       ("list_linears_dir.0", W // 2, W + self.channel_input_dir), ("rgb_linear", 3, W // 2)]
        else:
            shapes.append(("output_linear", self.channel_output, W))
        return shapes

    def _build_views(self):
        off = 0
        self._views = {}
        flat = self.flat.detach()  # detach() shares the version counter: in-place edits of a view mark the weights dirty
        for name, o, i in self._layer_shapes():
            w = flat[off:off + o * i].view(o, i)
            off += o * i
            b = flat[off:off + o]
            off += o
            self._views[name] = _LinearView(w, b)
        assert off == flat.numel()
        self.list_linears_pos = [self._views[f"list_linears_pos.{l}"] for l in range(self.D)]
        if self.is_use_view_directions:
            self.list_linears_dir = [self._views["list_linears_dir.0"]]
            self.feature_linear = self._views["feature_linear"]
            self.alpha_linear = self._views["alpha_linear"]
            self.rgb_linear = self._views["rgb_linear"]
        else:
            self.output_linear = self._views["output_linear"]

    def reset_parameters(self, seed=None):
        g = torch.Generator().manual_seed(0 if seed is None else int(seed))
        with torch.no_grad():
            for name, o, i in self._layer_shapes():
                s = 1.0 / math.sqrt(i)
                v = self._views[name]
                v.weight.copy_(((torch.rand(o, i, generator=g) * 2 - 1) * s).to(v.weight.device))
                v.bias.copy_(((torch.rand(o, generator=g) * 2 - 1) * s).to(v.bias.device))
        self._packed_version = -1

    def named_reference_parameters(self):
        out = {}
        for name, v in self._views.items():
            out[name + ".weight"] = v.weight
            out[name + ".bias"] = v.bias
        return out

    def load_reference_parameters(self, params):
        """params: dict name -> array/tensor with the reference's tree names (e.g. 'list_linears_pos.0.weight')."""
        with torch.no_grad():
            for k, dst in self.named_reference_parameters().items():
                src = torch.as_tensor(params[k]).to(dst.device, torch.float32)
                if tuple(src.shape) != tuple(dst.shape):
                    raise ValueError(f"{k}: shape {tuple(src.shape)} != {tuple(dst.shape)}")
                dst.copy_(src)
        self.mark_params_updated()

    def split_flat(self, flat):
        """Views of a flat vector (e.g. a gradient) by reference parameter name."""
        out, off = {}, 0
        for name, o, i in self._layer_shapes():
            out[name + ".weight"] = flat[off:off + o * i].view(o, i)
            off += o * i
            out[name + ".bias"] = flat[off:off + o]
            off += o
        return out

    # ---------------------------------------------------------------- plan / workspace
    def _ensure(self, points, training):
        lib = L.lib()
        if self._plan is None or (training and points > self._max_points):
            if self._plan is not None:
                lib.nmx_mlp_plan_destroy(self._plan)
                self._plan = None
            if training:
                self._max_points = max(self._max_points, int(points))
            plan = ctypes.c_void_p()
            L.call("nmx_mlp_plan_create", ctypes.byref(self._cfg), L.i64(self._max_points), ctypes.byref(plan))
            self._plan = plan
            self._ws = None
        if self._ws is None or (training and not self._ws_train):
            nbytes = int(lib.nmx_mlp_workspace_bytes(self._plan, L.i32(1 if training else 0)))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.flat.device)
            self._ws_train = bool(training)
            self._packed_version = -1
        if self._packed_version != self.flat._version:
            L.call("nmx_mlp_load_params", self._plan, L.ptr(self.flat), L.ptr(self._ws), L.stream())
            self._packed_version = self.flat._version

    def reserve(self, max_points, training=False):
        """Size the plan/workspace up front (avoids reallocation inside a timed or graph-captured region)."""
        with torch.cuda.device(self.flat.device):
            self._ensure(int(max_points), training)

    def mark_params_updated(self):
        self._packed_version = -1

    def _fwd_raw(self, enc_kind, x_or_rays, z, bands, B, n, save):
        L.require_cuda(x_or_rays, z, bands)
        P = int(B) * int(n)
        self._ensure(max(P, 1), bool(save))
        out = torch.empty((P, self.out_cols), dtype=torch.float32, device=self.flat.device)
        stride = x_or_rays.shape[-1] if enc_kind == 1 else 0
        L.call("nmx_mlp_fwd", self._plan, L.ptr(self._ws), L.ptr(self.flat), L.i32(enc_kind), L.ptr(x_or_rays),
               L.i32(stride), L.ptr(z), L.ptr(bands), L.ptr(out), L.i64(B), L.i32(n), L.i32(1 if save else 0), L.stream())
        if save:
            self._save_gen += 1
        return out

    def _bwd_raw(self, d_out, P, out=None, want_input_grad=False):
        if not self._ws_train:
            raise L.NmxError("nmx_mlp_bwd needs a forward with saved activations on this model first")
        g = out if out is not None else torch.empty_like(self.flat.data)
        if not want_input_grad:
            L.call("nmx_mlp_bwd", self._plan, L.ptr(self._ws), L.ptr(self.flat), L.ptr(d_out), L.ptr(g), L.i64(P), L.stream())
            return g
        # gradient w.r.t. the encoded position inputs [P, pos_pad] fp32; view-direction inputs get zeros (no learnable
        # encoder feeds them in the reference)
        cols = int(L.lib().nmx_mlp_input_grad_cols(self._plan))  # in_pos (fused width-64 kernel) or in_pos padded to 64
        d_pad = torch.empty((P, cols), dtype=torch.float32, device=g.device)
        L.call("nmx_mlp_bwd_input", self._plan, L.ptr(self._ws), L.ptr(self.flat), L.ptr(d_out), L.ptr(g), L.ptr(d_pad),
               L.i64(P), L.stream())
        d_x = d_pad[:, :self.channel_input_pos]
        if self.is_use_view_directions and self.channel_input_dir > 0:
            d_x = torch.cat([d_x, torch.zeros((P, self.channel_input_dir), dtype=torch.float32, device=g.device)], dim=-1)
        return g, d_x.contiguous()

    # ---------------------------------------------------------------- reference API
    def forward(self, x):
        """NeRF.forward (models/NeRF.py:201-243) on an already-encoded input x [P, channel_input (+ views)]."""
        x = x.float().contiguous()
        expect = self.channel_input_pos + (self.channel_input_dir if self.is_use_view_directions else 0)
        if x.shape[-1] != expect:
            raise ValueError(f"expected last dim {expect}, got {x.shape[-1]}")
        P = x.numel() // expect
        x2 = x.reshape(P, expect)
        if not torch.is_grad_enabled():  # inference: nothing is saved (ctx.needs_input_grad is not reliable under no_grad)
            return self._fwd_raw(0, x2, None, None, P, 1, save=False).reshape(*x.shape[:-1], self.out_cols)
        rows = self.max_save_points
        if torch.is_grad_enabled() and (self.flat.requires_grad or x2.requires_grad) and P > rows:
            out = torch.cat([_MLPFunction.apply(self.flat, self, 0, x2[i:i + rows], None, None, min(rows, P - i), 1)
                             for i in range(0, P, rows)], dim=0)
        else:
            out = _MLPFunction.apply(self.flat, self, 0, x2, None, None, P, 1)
        return out.reshape(*x.shape[:-1], self.out_cols)

    def forward_rays(self, rays, z_vals):
        """Fused path: Embedder PE generated inside the operand producer from rays [B, 11] and z [B, n] -> raw [B, n, C]."""
        rays = rays.float().contiguous()
        z_vals = z_vals.float().contiguous()
        B, n = z_vals.shape
        if not torch.is_grad_enabled():
            return self._fwd_raw(1, rays, z_vals, None, B, n, save=False).reshape(B, n, self.out_cols)
        rows = max(1, self.max_save_points // n)
        if torch.is_grad_enabled() and self.flat.requires_grad and B > rows:
            # a differentiable pass keeps ~10 KB of activations per point: bound the workspace by running ray chunks as
            # separate autograd nodes (their backward recomputes what a later chunk overwrote)
            outs = [_MLPFunction.apply(self.flat, self, 1, rays[i:i + rows], z_vals[i:i + rows], None, min(rows, B - i), n)
                    for i in range(0, B, rows)]
            return torch.cat(outs, dim=0).reshape(B, n, self.out_cols)
        out = _MLPFunction.apply(self.flat, self, 1, rays, z_vals, None, B, n)
        return out.reshape(B, n, self.out_cols)

    def forward_sinusoidal(self, x, bands):
        """Fused path for the image demo: SinusoidalEncoding generated in the operand producer from raw coords."""
        x = x.float().contiguous()
        P = x.shape[0]
        bands = bands.float().contiguous()
        if not torch.is_grad_enabled():
            return self._fwd_raw(2, x, None, bands, P, 1, save=False)
        return _MLPFunction.apply(self.flat, self, 2, x, None, bands, P, 1)

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None:
                L.lib().nmx_mlp_plan_destroy(self._plan)
        except Exception:
            pass


def inference_wrapper_batch(model, chunk):
    """inference_wrapper_batch (models/NeRF.py:10-22)."""
    if chunk is None:
        return model

    def __batched_model_inference(inputs_embedded):
        return torch.cat([model.forward(inputs_embedded[i:i + chunk]) for i in range(0, inputs_embedded.shape[0], chunk)], dim=0)

    return __batched_model_inference


def run_model(pos, embed_pos, dir, embed_dir, model, netchunk=64 * 1024):
    """run_model (models/NeRF.py:25-48): flatten -> embed -> chunked forward -> reshape [B, n, C]."""
    assert len(pos.shape) == 3, f"[ERROR] {pos.shape=} should have dimensions as: [n_rays, n_depth_samples, 3d position]!"
    B, n = pos.shape[0], pos.shape[1]
    inputs_embedded = embedding.embed(pos, embed_pos, dir, embed_dir)
    outputs_flat = inference_wrapper_batch(model, netchunk)(inputs_embedded)
    return outputs_flat.reshape(B, n, outputs_flat.shape[-1])


class AdamMLX:
    """optim.Adam of MLX 0.7.0 as the reference uses it (NeRF.py:120): no bias correction; ONE instance updates both
    nets and its state is keyed by the parameter tree, so coarse and fine SHARE moments (reference quirk,
    __test_nerf.py:134,144).  `shared_state=False` keeps separate moments per model."""

    def __init__(self, learning_rate, betas=(0.9, 0.999), eps=1e-8, shared_state=True, bias_correction=False):
        self.learning_rate = learning_rate
        self.betas = betas
        self.eps = eps
        self.shared_state = shared_state
        self.bias_correction = bias_correction
        self.state = {}
        self.step_count = 0

    def moments(self, model, like):
        """(m, v) of `model` -- ONE pair per parameter-tree shape when `shared_state` (the reference quirk)."""
        key = ("shared", like.numel()) if self.shared_state else id(model)
        if key not in self.state:
            self.state[key] = (torch.zeros_like(like), torch.zeros_like(like))
        return self.state[key]

    def update_exchange(self, model, xchg, buf_idx, lr_dev=None):
        """Data-parallel step: average gradient buffer `buf_idx` of the peer exchange over the ranks and update this
        replica, one kernel (nmx_allreduce_adam).  No bias correction (MLX 0.7)."""
        if self.bias_correction:
            raise NotImplementedError("the fused exchange + Adam kernel implements the MLX-style update only")
        m, v = self.moments(model, model.flat.data)
        self.step_count += 1
        xchg.allreduce_adam(buf_idx, model.flat.data, m, v, self.learning_rate, self.betas[0], self.betas[1], self.eps,
                            lr_dev=lr_dev)
        model.mark_params_updated()

    def update(self, model, grads=None, lr_dev=None):
        """`lr_dev`: optional device scalar holding the learning rate (used when the iteration is replayed from a CUDA
        graph, where a by-value learning rate would be frozen at capture time)."""
        from .. import ops
        g = grads if grads is not None else model.flat.grad
        m, v = self.moments(model, g)
        self.step_count += 1
        ops.adam_step(model.flat.data, g, m, v, self.learning_rate, self.betas[0], self.betas[1], self.eps,
                      self.bias_correction, self.step_count, lr_dev=lr_dev)
        model.mark_params_updated()


def create_NeRF(args, device="cuda"):
    """create_NeRF (models/NeRF.py:51-158) -> (render_kwargs_train, render_kwargs_test, idx_iter, optimizer).
    Quirk kept: render_kwargs_test IS render_kwargs_train (same dict), so perturb=False / raw_noise_std=0 apply to
    training too (NeRF.py:152-156)."""
    from ..rendering.render import render_rays, render_rays_eval

    octave_pos = args.multires
    octave_dir = args.multires_views
    is_use_dir = args.use_viewdirs
    n_samples = args.n_depth_samples
    n_importance_samples = args.N_importance
    output_ch = 5 if n_importance_samples else 4
    skips = [4]
    embedder_pos, channel_emb_pos = embedding.get_embedder(octave_pos)
    embedder_dir, channel_emb_dir = embedding.get_embedder(octave_dir) if is_use_dir else (None, None)

    def network_query_fn(inputs, viewdirs, model):
        return run_model(inputs, embedder_pos, viewdirs, embedder_dir, model, netchunk=args.netchunk)

    # lets render_rays recognise the closure and take the fused rays->raw path (same arithmetic, PE never materialised)
    network_query_fn.fused_embedders = (octave_pos, octave_dir if is_use_dir else None)

    def make(n_layers, width, seed):
        return NeRF(n_layers=n_layers, width_layers=width, channel_input=channel_emb_pos, channel_output=output_ch,
                    list_skip_connection_layers=skips, channel_input_views=channel_emb_dir if is_use_dir else 0,
                    is_use_view_directions=is_use_dir, device=device, seed=seed,
                    n_freqs_pos=max(octave_pos, 0), n_freqs_dir=max(octave_dir, 0) if is_use_dir else 0)

    model_coarse = make(args.netdepth, args.netwidth, getattr(args, "seed", 0))
    model_fine = make(args.netdepth_fine, args.netwidth_fine, getattr(args, "seed", 0) + 1) if n_importance_samples > 0 else None
    optimizer = AdamMLX(learning_rate=args.lrate, betas=(0.9, 0.999))
    idx_iter = 0
    render_kwargs_train = {
        "use_viewdirs": is_use_dir,
        "white_bkgd": args.white_bkgd,
        "network_query_fn": network_query_fn,
        "is_test": True,
        "render_rays_func": render_rays,
        "network_coarse": model_coarse,
        "n_depth_samples": n_samples,
        "network_fine": model_fine,
        "perturb": args.perturb,
        "raw_noise_std": args.raw_noise_std,
        "N_importance": n_importance_samples,
    }
    if args.dataset_type != "llff" or args.no_ndc:
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = render_kwargs_train
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0
    render_kwargs_test["is_test"] = False
    render_kwargs_test["render_rays_func"] = render_rays_eval
    return render_kwargs_train, render_kwargs_test, idx_iter, optimizer


def default_args(**over):
    """The hot-path-relevant defaults of the reference's config_parser (config_parser.py:3-80)."""
    d = dict(multires=10, multires_views=4, use_viewdirs=True, n_depth_samples=64, N_importance=0, lrate=5e-4,
             lrate_decay=250, perturb=1.0, raw_noise_std=0.0, netdepth=8, netwidth=256, netdepth_fine=8,
             netwidth_fine=256, netchunk=1024 * 64, chunk=1024 * 32, white_bkgd=True, dataset_type="blender",
             no_ndc=False, lindisp=False, N_rand=4096, seed=0)
    d.update(over)
    return SimpleNamespace(**d)
