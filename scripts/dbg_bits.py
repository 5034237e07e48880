import ctypes, sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_mlp_gpu as T
from nerf_meets_mlx_b200 import _lib_loader as L
cfg = T.CFGS[0]
P = 128
torch.manual_seed(P)
ref, net = T.make_pair(**cfg)
x = torch.randn(P, 90).clamp(-1, 1)
y = net.forward(x.cuda().requires_grad_(False)) if False else None
net.flat.requires_grad_(True)
y = net.forward(x.cuda())
torch.cuda.synchronize()
out = (ctypes.c_int64 * 12)()
L.call("nmx_mlp_debug_layout", net._plan, out, L.i32(12))
base, x0, h0, hs, feat, hd, g0, gs, ghd, bits, cap, x0c = list(out)
ws = net._ws
def bf16_region(off, rows, cols):
    return ws[base + off: base + off + rows * cols * 2].view(torch.bfloat16).view(rows, cols).float().cpu()
b = ws[base + bits: base + bits + 9 * cap * 32].view(torch.int32).view(9, cap, 8).cpu().numpy().astype(np.uint32)
for l in list(range(8)) + [8]:
    if l < 8:
        h = bf16_region(h0 + l * hs, P, 256)
    else:
        h = bf16_region(hd, P, 128)
    exp = (h > 0).numpy()
    got = np.zeros_like(exp)
    ncol = exp.shape[1]
    for w in range(ncol // 32):
        for e in range(16):
            got[:, 32 * w + 2 * e] = (b[l, :P, w] >> e) & 1
            got[:, 32 * w + 2 * e + 1] = (b[l, :P, w] >> (16 + e)) & 1
    bad = (got != exp)
    print("slot", l, "mismatch", int(bad.sum()), "of", exp.size, "rows with mismatch", int(bad.any(1).sum()),
          "cols(words) bad", sorted(set((np.where(bad)[1] // 32).tolist()))[:8])

# ---- backward: check every saved dY against its mask
g_out = torch.randn(P, 4)
(y * g_out.cuda()).sum().backward()
torch.cuda.synchronize()
for l in range(7, -1, -1):
    dy = bf16_region(g0 + l * gs, P, 256)
    h = bf16_region(h0 + l * hs, P, 256)
    nz_where_masked = ((dy != 0) & (h <= 0)).numpy()
    z_where_open = ((dy == 0) & (h > 0)).numpy()
    print("dY slot", l, "nonzero where h<=0:", int(nz_where_masked.sum()), " zero where h>0:", int(z_where_open.sum()),
          "bad words:", sorted(set((np.where(nz_where_masked | z_where_open)[1] // 32).tolist())),
          "bad rows (first 8):", sorted(set(np.where(nz_where_masked | z_where_open)[0].tolist()))[:8])

prm = {k: v.detach().float().cpu() for k, v in net.named_reference_parameters().items()}
def rb(t): return t.to(torch.bfloat16).float()
for l in range(7, 0, -1):
    dy = bf16_region(g0 + l * gs, P, 256)
    W = prm[f"list_linears_pos.{l}.weight"]
    Wh = W[:, -256:]
    h = bf16_region(h0 + (l - 1) * hs, P, 256)
    exp = (dy @ rb(Wh)) * (h > 0)
    got = bf16_region(g0 + (l - 1) * gs, P, 256)
    err = (got - exp).norm() / exp.norm()
    bad = ((got - exp).abs() > 0.05 * exp.abs().max()).numpy()
    print(f"dgrad of layer {l} -> dY_{l-1}: rel err {float(err):.4f}; bad elems {int(bad.sum())}; bad col chunks "
          f"{sorted(set((np.where(bad)[1] // 64).tolist()))}; bad rows {len(set(np.where(bad)[0].tolist()))}")

H = [bf16_region(h0 + k * hs, P, 256) for k in range(8)]
for l in range(7, -1, -1):
    dy = bf16_region(g0 + l * gs, P, 256)
    bad = (((dy != 0) & (H[l] <= 0)) | ((dy == 0) & (H[l] > 0))).numpy()
    if not bad.any():
        continue
    rows, words = np.where(bad)[0], np.where(bad)[1] // 32
    for (r, w) in sorted(set(zip(rows.tolist(), words.tolist())))[:6]:
        seg = slice(32 * w, 32 * w + 32)
        pat = (dy[r, seg] != 0).numpy()
        match = [k for k in range(8) if ((H[k][r, seg] > 0).numpy() | ~pat).all() and ((dy[r, seg] == 0).numpy() | (H[k][r, seg] > 0).numpy()).all()]
        agree = [int(((H[k][r, seg] > 0).numpy() == pat).sum()) for k in range(8)]
        print(f"dY_{l} row {r} word {w}: nonzero pattern agreement with h_k>0 (k=0..7): {agree}")
