import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_mlp_gpu as T
cfg = T.CFGS[0]
for P in (128,):
    torch.manual_seed(P)
    ref, net = T.make_pair(**cfg)
    x = torch.randn(P, 90).clamp(-1, 1)
    ref.requires_grad_(True)
    y = net.forward(x.cuda())
    g_out = torch.randn(P, 4)
    names = list(ref.params.keys())
    y_emu = T.emulated_forward(ref, x)
    g_emu = torch.autograd.grad((y_emu * g_out).sum(), [ref.params[n] for n in names])
    (y * g_out.cuda()).sum().backward()
    got = net.split_flat(net.flat.grad)
    for n, ge in zip(names, g_emu):
        print(P, n, float(T.rel_norm(got[n].cpu(), ge)))
