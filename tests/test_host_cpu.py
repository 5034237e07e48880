"""CPU-side checks: the C-ABI library loads and exports every symbol include/nmx.h declares, argument validation
fails loudly without a GPU, host-side logic of the mirror packages, and the data-parallel plumbing on gloo (world 2)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nerf_meets_mlx_b200 import build as B
    B.build()
    from nerf_meets_mlx_b200 import _lib_loader as L
    return L


def test_library_exports_every_declared_symbol(lib):
    syms = lib.declared_symbols()
    assert len(syms) >= 25 and "nmx_mlp_fwd" in syms and "nmx_composite_bwd" in syms
    l = lib.lib()
    missing = [s for s in syms if not hasattr(l, s)]
    assert not missing, missing
    assert l.nmx_version() == 100


def test_bad_arguments_fail_loudly(lib):
    with pytest.raises(lib.NmxError, match="bad argument"):
        lib.call("nmx_sample_z_fwd", lib.ptr(None), lib.ptr(None), lib.ptr(None), lib.i64(4), lib.i32(1), lib.i32(0), lib.ptr(None))
    with pytest.raises(lib.NmxError, match="n <= 256"):
        lib.call("nmx_composite_bwd", lib.ptr(8), lib.ptr(8), lib.ptr(8), lib.i32(3), lib.ptr(None), lib.f32(0), lib.i32(0),
                 lib.ptr(8), lib.ptr(None), lib.ptr(None), lib.ptr(None), lib.ptr(None), lib.ptr(8), lib.i64(4), lib.i32(512), lib.ptr(None))


def test_no_cpu_fallback(lib):
    from nerf_meets_mlx_b200 import ops
    with pytest.raises(lib.NmxError, match="CUDA"):
        ops.composite_fwd(torch.zeros(2, 4, 4), torch.zeros(2, 4), torch.zeros(2, 3))
    with pytest.raises(lib.NmxError, match="CUDA"):
        ops.sample_pdf(torch.zeros(2, 4), torch.zeros(2, 4, 1), torch.zeros(2, 8))


def test_param_count_matches_reference_geometry(lib):
    from nerf_meets_mlx_b200.models.NeRF import _Cfg
    c = _Cfg(8, 256, 63, 27, 5, 4, 1, 10, 4)
    assert lib.lib().nmx_mlp_param_count(ctypes.byref(c)) == 595844
    c = _Cfg(8, 256, 40, 0, 3, 4, 0, 10, 0)
    assert lib.lib().nmx_mlp_param_count(ctypes.byref(c)) == 482051  # image net


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "nerf_meets_mlx_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_encoder_host_logic():
    from nerf_meets_mlx_b200.encoding.sinusoidal import SinusoidalEncoding, mlx_linspace
    from nerf_meets_mlx_b200.models.embedding import get_embedder
    from oracle import encoding as oenc, sampling as osamp
    enc = SinusoidalEncoding(2, 10, min_freq_exp=0.0, max_freq_exp=8.0)
    assert enc.get_out_dim() == 40
    np.testing.assert_array_equal(mlx_linspace(0.0, 8.0, 10).numpy(), osamp.linspace_mlx(0.0, 8.0, 10))
    np.testing.assert_allclose(enc.freq_bands("cpu").numpy(), oenc.sinusoidal_freq_bands(10, 0.0, 8.0), rtol=2e-7)
    assert SinusoidalEncoding(3, 4, is_include_input=True).get_out_dim() == 27
    assert SinusoidalEncoding(3, 4, min_freq_exp=0.0).min_freq_exp == 0.0 and SinusoidalEncoding(3, 4).max_freq_exp == 3.0
    _, d = get_embedder(10)
    assert d == 63
    _, d = get_embedder(4)
    assert d == 27
    _, d = get_embedder(6, n_input_dims=2)
    assert d == 24
    f, d = get_embedder(-1)
    assert d == 3


def test_lr_schedule_and_flop_accounting():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.FWD_FLOP_PT == 1186816 and bench.TRAIN_FLOP_PT == 3489024
    assert bench.FLOP_PER_RAY == 969146368  # 969.1 MFLOP/ray
    from oracle.training import lr_schedule
    assert abs(lr_schedule(250000) - 5e-5) < 1e-12


_DP_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# the trainer's gradient averaging and parameter broadcast, exercised on CPU tensors over gloo
from nerf_meets_mlx_b200.training import NeRFTrainer
t = NeRFTrainer.__new__(NeRFTrainer)
t.pg = None; t.world = world
g = torch.full((1000,), float(rank + 1))
t._allreduce_mean(g)
assert torch.allclose(g, torch.full((1000,), (1 + world) / 2.0)), g[:3]
class M:
    def __init__(self): self.flat = torch.nn.Parameter(torch.full((10,), float(rank))); self.dirty = False
    def mark_params_updated(self): self.dirty = True
t.coarse, t.fine = M(), None
t.broadcast_parameters()
assert torch.equal(t.coarse.flat.data, torch.zeros(10)) and t.coarse.dirty
# ray-tile sharding of a frame (rendering.render(process_group=), NeRFTrainer.render_frame): contiguous tiles cover the
# frame exactly once, and the gathered tiles reassemble to the single-process result, incl. ragged / empty last tiles
from nerf_meets_mlx_b200.parallel import ray_tile, gather_tiles
for n in (640000, 7, 1, 2 * 33 + 1):
    lo, hi = ray_tile(n, rank, world)
    tot = torch.tensor([hi - lo]); dist.all_reduce(tot); assert int(tot) == n
    frame = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) * 0.5 + 1.0   # what one process would render
    got = gather_tiles(frame[lo:hi].clone(), n)
    assert got.shape == frame.shape and torch.equal(got, frame), (n, rank)
    got1 = gather_tiles(frame[lo:hi, :1].clone(), n)
    assert torch.equal(got1, frame[:, :1])
dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "dp.py"
    script.write_text(_DP_SCRIPT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_custom_ops_registered_with_fake_kernels():
    """torch.ops.nmx.* exist (north_star: "PyTorch custom ops over a thin C-ABI layer") and their fake kernels propagate
    shapes / dtypes without running CUDA code (FakeTensorMode builds fake cuda tensors on this GPU-less host)."""
    import torch
    from torch._subclasses import FakeTensorMode
    from nerf_meets_mlx_b200.ops import library as lib
    for name in lib.ALL_OPS:
        assert hasattr(torch.ops.nmx, name), name
    B, n, N = 7, 64, 128
    with FakeTensorMode():
        c = lambda *s: torch.empty(*s, device="cuda")
        z = torch.ops.nmx.sample_z(c(B), c(B), n)
        assert z.shape == (B, n) and z.device.type == "cuda"
        rgb, disp, acc, w, depth = torch.ops.nmx.composite_fwd(c(B, n, 4), z, c(B, 3))
        assert rgb.shape == (B, 3) and w.shape == (B, n, 1) and depth.shape == (B, 1)
        assert torch.ops.nmx.composite_bwd(c(B, n, 4), z, c(B, 3), rgb).shape == (B, n, 4)
        zi, zm = torch.ops.nmx.sample_pdf(z, w, c(B, N))
        assert zi.shape == (B, N) and zm.shape == (B, n + N)
        assert torch.ops.nmx.pe_embedder(c(5, 3), 10).shape == (5, 63)
        assert torch.ops.nmx.pe_sinusoidal(c(5, 2), c(10)).shape == (5, 40)
        assert torch.ops.nmx.sh_encode(c(5, 3), 4).shape == (5, 25)
        feat = torch.ops.nmx.hashgrid_fwd(c(9, 3), c(16, 1 << 10, 2), c(16), 10)
        assert feat.shape == (9, 32)
        assert torch.ops.nmx.hashgrid_bwd(c(9, 3), c(16), feat, 16, 2, 10).shape == (16, 1 << 10, 2)
        assert torch.ops.nmx.assemble_rays(c(B, 3), c(B, 3), 2.0, 6.0).shape == (B, 11)
        loss, d = torch.ops.nmx.mse_fwd_bwd(rgb, c(B, 3))
        assert loss.shape == (1,) and d.shape == (B, 3)
        out = torch.ops.nmx.mlp_fwd(0, c(1024, dtype=torch.uint8) if False else torch.empty(1024, dtype=torch.uint8, device="cuda"),
                                    c(100), 1, c(B, 11), z, None, B, n, 4, False)
        assert out.shape == (B * n, 4)
