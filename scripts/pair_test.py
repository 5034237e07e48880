import sys, torch
sys.path.insert(0, ".")
from nerf_meets_mlx_b200 import ops
torch.manual_seed(0)
for M, K, mp in ((256, 64, 1), (1000, 256, 0), (256 * 74 * 3 + 77, 320, 0)):
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(256, K, device="cuda").bfloat16()
    D = ops.gemm_pair_bf16(A, B, mp)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T
    err = float((D - ref).abs().max() / ref.abs().max())
    print(M, K, "rel max err", err, flush=True)
    if err > 1e-3:
        bad = torch.nonzero((D - ref).abs() > 1e-2 * ref.abs().max())
        print(" bad entries:", bad.shape[0], bad[:5].tolist(), bad[-5:].tolist())
# rate
M, K = 1 << 20, 256
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(256, K, device="cuda").bfloat16()
for _ in range(3): ops.gemm_pair_bf16(A, B)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): ops.gemm_pair_bf16(A, B)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"pair gemm M=2^20 K=256: {ms:.3f} ms, {2 * M * 256 * K / ms / 1e9:.0f} TFLOP/s (fp32 output: write-bound {M * 1024 / ms / 1e6:.0f} GB/s)")
