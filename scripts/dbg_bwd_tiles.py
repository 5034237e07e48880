"""Which (layer, tile) of the fused backward chain disagrees with a torch recomputation from the saved tensors."""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_mlp_gpu as T
from nerf_meets_mlx_b200 import _lib_loader as L
cfg = T.CFGS[0]
P = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 2 + 128 * 5 + 77
torch.manual_seed(1)
ref, net = T.make_pair(**cfg)
x = torch.randn(P, 90).clamp(-1, 1)
net.flat.requires_grad_(True)
y = net.forward(x.cuda())
g_out = torch.randn(P, 4)
(y * g_out.cuda()).sum().backward()
torch.cuda.synchronize()
out = (ctypes.c_int64 * 12)()
L.call("nmx_mlp_debug_layout", net._plan, out, L.i32(12))
base, x0, h0, hs, feat, hd, g0, gs, ghd, bits, cap, x0c = list(out)
ws = net._ws
def reg(off, rows, cols):
    return ws[base + off: base + off + rows * cols * 2].view(torch.bfloat16).view(rows, cols).float()
prm = {k: v.detach().float() for k, v in net.named_reference_parameters().items()}
def rb(t): return t.to(torch.bfloat16).float()
go = g_out.cuda()
# step A
hdv = reg(hd, P, 128)
exp = rb((go[:, :3] @ prm["rgb_linear.weight"]) * (hdv > 0))
got = reg(ghd, P, 128)
def report(name, got, exp):
    err = (got - exp).abs().amax(dim=1) / exp.abs().max().clamp_min(1e-20)
    bad_rows = torch.nonzero(err > 0.05).flatten().cpu().numpy()
    tiles = sorted(set((bad_rows // 128).tolist()))
    print(f"{name}: rel err {float((got - exp).norm() / exp.norm()):.4f}  bad rows {len(bad_rows)}  bad tiles {tiles[:12]}{'...' if len(tiles) > 12 else ''} (n={len(tiles)})")
report("d_hd", got, exp)
Wd = prm["list_linears_dir.0.weight"][:, :256]
dfe = reg(g0 + 8 * gs, P, 256)
report("d_feature", dfe, rb(got @ rb(Wd)))
h7 = reg(h0 + 7 * hs, P, 256)
exp7 = rb((dfe @ rb(prm["feature_linear.weight"]) + go[:, 3:4] * prm["alpha_linear.weight"]) * (h7 > 0))
dy = reg(g0 + 7 * gs, P, 256)
report("dY_7", dy, exp7)
for l in range(7, 0, -1):
    W = prm[f"list_linears_pos.{l}.weight"][:, -256:]
    h = reg(h0 + (l - 1) * hs, P, 256)
    exp = rb((dy @ rb(W)) * (h > 0))
    nxt = reg(g0 + (l - 1) * gs, P, 256)
    report(f"dY_{l-1}", nxt, exp)
    dy = nxt
