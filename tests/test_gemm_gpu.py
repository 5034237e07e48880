"""tcgen05 GEMM building blocks vs torch fp32 matmul of the same bf16 operands (floating-point kernel:
torch fp32 reference, tolerance = bf16 output rounding)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from nerf_meets_mlx_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 256), (1000, 256, 320), (128 * 150 + 17, 256, 256),
                                   (4096, 128, 320), (300, 64, 128), (5, 64, 64)])
def test_gemm_kmajor(ops, M, N, K):
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda") * 0.1).bfloat16()
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ B.float().T + bias
    out = ops.gemm_bf16(A, B, bias, relu=False, out_fp32=True)
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3), (out - ref).abs().max().item()
    out = ops.gemm_bf16(A, B, bias, relu=True, out_fp32=False)
    refb = torch.relu(ref)
    assert torch.allclose(out.float(), refb, rtol=1e-2, atol=1e-2), (out.float() - refb).abs().max().item()


@pytest.mark.parametrize("P,M,N", [(64, 128, 64), (128, 128, 256), (1000, 256, 256), (70000, 256, 256),
                                   (4100, 128, 64), (333, 256, 128)])
def test_wgrad_mnmajor(ops, P, M, N):
    torch.manual_seed(P + M + N)
    dY = (torch.randn(P, M, device="cuda") * 0.2).bfloat16()
    X = (torch.randn(P, N, device="cuda") * 0.5).bfloat16()
    ref = dY.float().T @ X.float()
    out, db = ops.wgrad_bf16(dY, X, want_db=True)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-3 * scale), ((out - ref).abs().max().item(), scale)
    ref_b = dY.float().sum(0)
    assert torch.allclose(db, ref_b, rtol=1e-3, atol=1e-3 * ref_b.abs().max().item()), (db - ref_b).abs().max().item()
    out2 = ops.wgrad_bf16(dY, X)
    assert torch.equal(out2 != 0, out != 0)


def test_colsum(ops):
    torch.manual_seed(0)
    Y = torch.randn(10007, 256, device="cuda").bfloat16()
    ref = Y.float().sum(0)
    out = ops.colsum_bf16(Y)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("M,K,max_pairs", [(256, 64, 1), (1000, 256, 0), (256 * 74 * 3 + 77, 320, 0)])
def test_gemm_cta_pair(ops, M, K, max_pairs):
    """tcgen05 cta_group::2 building block: both CTAs of a pair stage their rows of A and their half of B, the leader
    issues M = 256 MMAs, each CTA reads back its own 128 accumulator rows (ragged last tile, several tiles per pair)."""
    torch.manual_seed(M)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(256, K, device="cuda").bfloat16()
    D = ops.gemm_pair_bf16(A, B, max_pairs)
    ref = A.float() @ B.float().T
    assert float((D - ref).abs().max() / ref.abs().max()) < 1e-5
