"""Isolated timing of the tcgen05 building blocks at the fine-pass size (P = 8192 rays x 192 samples)."""
import sys
import torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import ops

P = 8192 * 192
torch.manual_seed(0)
A = (torch.randn(P, 256, device="cuda") * 0.5).bfloat16()
B = (torch.randn(256, 256, device="cuda") * 0.1).bfloat16()
bias = torch.randn(256, device="cuda")
dY = (torch.randn(P, 256, device="cuda") * 0.1).bfloat16()


def bench(fn, flops, name, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)


bench(lambda: ops.gemm_bf16(A, B, bias, relu=True), 2.0 * P * 256 * 256, "gemm 256x256 fwd (bias+relu, bf16 out)")
bench(lambda: ops.wgrad_bf16(dY, A, want_db=True), 2.0 * P * 256 * 256, "wgrad 256x256 (+db)")
bench(lambda: torch.matmul(A, B.T), 2.0 * P * 256 * 256, "torch/cuBLAS matmul (no epilogue)")
