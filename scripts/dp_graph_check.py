"""Data-parallel check (torchrun, >= 2 GPUs): the three-graph replay of the training iteration (all-reduces eager between the
graphs) follows the eager iteration step for step, and the ranks' parameters stay identical.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_graph_check.py
"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from nerf_meets_mlx_b200.models.NeRF import default_args
from nerf_meets_mlx_b200.training import NeRFTrainer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
B = 1024
losses = {}
for mode in ("eager", "graph"):
    torch.manual_seed(0)
    tr = NeRFTrainer(default_args(N_importance=128, n_depth_samples=64), device="cuda", max_rays=B,
                     use_cuda_graph=(mode == "graph"))
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    out = []
    for it in range(5):
        o = torch.randn(B, 3, device="cuda", generator=g) * 0.2 + torch.tensor([0.0, 0.0, 4.0], device="cuda")
        d = torch.randn(B, 3, device="cuda", generator=g) - torch.tensor([0.0, 0.0, 3.0], device="cuda")
        t = torch.rand(B, 3, device="cuda", generator=g)
        u = torch.rand(B, 128, device="cuda", generator=g)
        r = tr.train_iteration(o, d, t, u_vals=u)
        out.append((float(r["loss_coarse"]), float(r["loss_fine"])))
    losses[mode] = out
    flat = torch.cat([tr.coarse.flat.data, tr.fine.flat.data])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(ref, flat), f"rank {rank}: parameters diverged from rank 0 in {mode} mode"
    if mode == "graph":
        assert tr._graph is not None and len(tr._graph) == 3
ok = all(abs(a - b) <= 2e-3 * abs(a) for e, gph in zip(losses["eager"], losses["graph"]) for a, b in zip(e, gph))
print(f"rank {rank}: eager {losses['eager'][-1]} graph {losses['graph'][-1]} match={ok}", flush=True)
assert ok
dist.barrier()
dist.destroy_process_group()
