import sys, torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models import NeRF
net = NeRF(n_layers=8, width_layers=256, channel_input=63, channel_input_views=27, channel_output=5,
           list_skip_connection_layers=[4], is_use_view_directions=True, max_points=4096)
x = torch.randn(256, 90, device="cuda").clamp(-1, 1)
y = net.forward(x)
torch.cuda.synchronize()
print("fwd ok", y.abs().mean().item())
(y * torch.randn_like(y)).sum().backward()
torch.cuda.synchronize()
print("bwd ok", net.flat.grad.abs().mean().item())
