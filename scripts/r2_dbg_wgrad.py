import os, sys, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import models as omodels
from nerf_meets_mlx_b200.models import NeRF
KW = dict(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
          list_skip_connection_layers=[4], is_use_view_directions=True)
P = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 2 + 128 * 5 + 77
torch.manual_seed(1)
ref = omodels.NeRF(seed=5, **KW)
net = NeRF(device="cuda", **KW)
net.load_reference_parameters(ref.params)
x = torch.randn(P, 90).clamp(-1, 1)
g_out = torch.randn(P, 4)
y = net.forward(x.cuda())
(y * g_out.cuda()).sum().backward()
got = {k: v.cpu().clone() for k, v in net.split_flat(net.flat.grad).items()}
f = "gpurun_out/dbg_wgrad_ref.pt"
if sys.argv[1] == "ref":
    torch.save(got, f)
else:
    refg = torch.load(f)
    for k in got:
        a, b = got[k], refg[k]
        nn = int(torch.isnan(a).sum())
        d = float((a - b).norm() / b.norm().clamp_min(1e-30)) if nn == 0 else float("nan")
        if "bias" in k and nn: print("   sample", a.flatten()[:4].tolist(), "ref", b.flatten()[:4].tolist())
        print(f"{k:32s} nan {nn:6d}/{a.numel():6d}  rel diff vs per-layer {d:.3e}")
