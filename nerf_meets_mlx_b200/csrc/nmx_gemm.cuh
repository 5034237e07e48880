// Internal C++ API of the tcgen05 GEMM building blocks (nmx_gemm.cu), used by nmx_mlp.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmx {

// D[M,N] = epi(A[M,K] * B[N,K]^T); A = K-concatenation of up to two row-major bf16 tensors.
struct GemmDesc {
  const void* A0; int64_t a0_rows; int a0_cols, a0_ld, a0_col, a0_k;
  const void* A1; int a1_cols, a1_ld, a1_col, a1_k;   // A1 may be null
  const void* B; int b_rows, b_cols, b_ld, b_col;     // B: [N, K] row-major bf16
  int64_t M; int N;
  const float* bias; void* D; int ldd; int out_fp32; int relu;
  int d_cols, d_col;                                  // D tensor width (0 = N) and first output column
  const void* mask; int ldmask; int mask_cols, mask_col;  // output *= (mask > 0); mask tensor [M, mask_cols]
  const float* row_vec; int row_stride; const float* col_vec;  // + row_vec[m] * col_vec[n] before the mask
  int accum;                                          // fp32 output: D += result
};

// dW[M, w_col + (0..n_valid)) += dY[P, dy_col + (0..M))^T * X[P, x_col + (0..N))
struct WgradDesc {
  const void* dY; int dy_cols, dy_ld, dy_col;
  const void* X; int x_cols, x_ld, x_col;
  int64_t P; int M, N;                                // M % 64 == 0, N % 64 == 0 (<= 256)
  float* dW; int ldw, w_col; int n_valid;             // only columns < n_valid are accumulated
  float* db;                                          // optional: db[m] += sum_p dY[p, m] (bias gradient), or null
  int max_ctas;                                       // 0 = all SMs; else cap on the CTA count (concurrent kernels)
  // optional second operand (64 columns of X2 from x2_col) contracted with the same dY in the same pass:
  // dW2[m, w2_col + n] += sum_p dY[p, m] X2[p, x2_col + n] for n < n_valid2 (<= 64)
  const void* X2; int x2_cols, x2_ld, x2_col;
  float* dW2; int ldw2, w2_col, n_valid2;
};

// live profiling hooks (nmx_profile_enable): kind 0 layer GEMM, 1 wgrad, 2 chain forward (inference), 3 chain forward
// (training, saves activations), 4 chain backward (data gradients)
void prof_begin(int kind, double flops, cudaStream_t s);
void prof_end(cudaStream_t s);

int launch_gemm(const GemmDesc& g, cudaStream_t stream);
int launch_wgrad(const WgradDesc& g, cudaStream_t stream);
int launch_colsum(const void* Y, int ld, int col0, int N, int64_t P, float* out, cudaStream_t stream);

}  // namespace nmx

namespace nmx {

// ---- batched weight gradients: every wgrad of one backward pass in ONE launch (nmx_wgrad_batch.cu).
// Operands are row-major bf16 tensors addressed per job by (tensor, first row, first column): either a few whole
// regions (saved activations, saved dY, X0, d_hd: the fused chains write whole 128-row tiles, so rows past P are finite)
// or one tensor per slot with rows = P (layer-by-layer path: its stores are clipped at P, and TMA zero-fills the rest).
constexpr int kMaxWgradJobs = 12;
constexpr int kMaxWgradTensors = 26;
struct WgradBatchTensor { const void* base; int64_t rows; int cols; };  // row-major bf16 [rows, cols], ld = cols
struct WgradBatchJob {
  int dy_t, x_t, x2_t;           // tensor indices (x2_t < 0: no second operand)
  int64_t dy_row0, x_row0, x2_row0;
  int dy_col, x_col, x2_col;
  int M, N;                      // dW rows (multiple of 64, <= 256), first-operand width (multiple of 64, <= 256)
  float* dW; int ldw, w_col, n_valid;
  float* db;                     // optional bias gradient [M]
  float* dW2; int ldw2, w2_col, n_valid2;  // second operand: 64 columns of tensor x2_t
};
struct WgradBatchDesc {
  int n_tensors; WgradBatchTensor t[kMaxWgradTensors];
  int n_jobs; WgradBatchJob job[kMaxWgradJobs];
  int64_t P;                     // points (contraction length), the same for every job
};
int launch_wgrad_batch(const WgradBatchDesc& d, cudaStream_t stream);

}  // namespace nmx
