// K5: inverse-CDF importance resampling (sample_from_inverse_cdf_torch, sampling/__init__.py:101-178)
// fused with the sort-merge of coarse and importance depths (rendering/render.py:225, __test_nerf.py:288).
//
// One warp per ray.  The ray's CDF (n+1 edges), padded bin mid-points (n+1) and the N importance samples live in
// shared memory; every lane owns N/32 samples and runs a branch-free binary search over the shared CDF
// (searchsorted side="right" == number of edges <= u).  The merge ranks every value by counting
// (rank = #smaller + #equal-with-lower-concat-index), i.e. exactly a stable sort of concat([z, z_imp]).
//
// CDF arithmetic: w = weights + 0.01 ; sum and running prefix in fp64 rounded to fp32
// (what torch-CPU cumsum does; both are exact in fp64 for <= 2^20 fp32 addends of this dynamic range, so the
// warp-parallel scan order cannot change the result) ; pdf = w / sum in fp32 ; cdf = min(1, prefix).
// HBM roofline: 1792 B/ray at n=64, N=128 (z 256 + w 256 + u 512 read, z_merged 768 write).
#include "nmx_common.cuh"

using namespace nmx;

namespace {

constexpr int kWarps = 4;

__global__ void __launch_bounds__(kWarps * 32)
sample_pdf_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u,
                  const float* __restrict__ cdf_in, float eps, float* __restrict__ z_imp, int32_t* __restrict__ inds,
                  float* __restrict__ cdf_out, float* __restrict__ z_merged, int64_t B, int n, int N) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int per_warp = 2 * (n + 1) + n + N;
  float* s_cdf = smem + (size_t)wib * per_warp;  // [n+1]
  float* s_mid = s_cdf + (n + 1);                // [n+1] padded mid-points
  float* s_z = s_mid + (n + 1);                  // [n] coarse depths
  float* s_imp = s_z + n;                        // [N] importance samples (unsorted)
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int n1 = n + 1;

  for (int64_t b = warp0; b < B; b += nwarps) {
    const float* zr = z + b * n;
    // ---- coarse depths + padded mid-points (:149-159): mid[0]=mid[1]=m_0 ... mid[n]=m_{n-2}
    for (int i = lane; i < n; i += 32) s_z[i] = zr[i];
    __syncwarp();
    for (int i = lane; i < n1; i += 32) {
      int j = min(max(i - 1, 0), n - 2);
      s_mid[i] = __fdiv_rn(__fadd_rn(s_z[j + 1], s_z[j]), 2.0f);
    }
    // ---- CDF
    if (cdf_in != nullptr) {
      for (int i = lane; i < n1; i += 32) s_cdf[i] = cdf_in[b * n1 + i];
    } else {
      const float* wr = weights + b * n;
      double part = 0.0;
      for (int i = lane; i < n; i += 32) part += (double)__fadd_rn(wr[i], 0.01f);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      float wsum = (float)part;
      float padding = fmaxf(__fsub_rn(eps, wsum), 0.0f);  // relu(eps - sum) (:115)
      float padw = __fdiv_rn(padding, (float)n);
      wsum = __fadd_rn(wsum, padding);
      double carry = 0.0;
      if (lane == 0) s_cdf[0] = 0.0f;
      for (int c = 0; c < n; c += 32) {
        int i = c + lane;
        float pdf = 0.0f;
        if (i < n) pdf = __fdiv_rn(__fadd_rn(__fadd_rn(wr[i], 0.01f), padw), wsum);
        double incl = warp_scan_incl_f64((double)pdf, lane) + carry;
        if (i < n) s_cdf[i + 1] = fminf(1.0f, (float)incl);
        carry = __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncwarp();
    if (cdf_out != nullptr)
      for (int i = lane; i < n1; i += 32) cdf_out[b * n1 + i] = s_cdf[i];

    // ---- inverse-CDF lookup for this lane's samples
    for (int j = lane; j < N; j += 32) {
      float uu = u[b * N + j];
      // searchsorted(side="right"): count of edges <= uu in s_cdf[0..n]
      int lo = 0, hi = n1;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      int ind = lo;
      int below = min(max(ind - 1, 0), n);
      int above = min(max(ind, 0), n);
      float c0 = s_cdf[below], c1 = s_cdf[above];
      float m0 = s_mid[below], m1 = s_mid[above];
      float num = __fsub_rn(uu, c0);
      float den = __fsub_rn(c1, c0);
      if (den < eps) den = 1.0f;
      float t = __fdiv_rn(num, den);
      if (t != t) t = 0.0f;  // nan_to_num
      t = fminf(fmaxf(t, 0.0f), 1.0f);
      float v = __fadd_rn(m0, __fmul_rn(t, __fsub_rn(m1, m0)));
      s_imp[j] = v;
      if (z_imp != nullptr) z_imp[b * N + j] = v;
      if (inds != nullptr) inds[b * N + j] = ind;
    }
    __syncwarp();

    // ---- stable sort of concat([z (n), imp (N)]) by rank counting
    if (z_merged != nullptr) {
      float* outr = z_merged + b * (int64_t)(n + N);
      for (int i = lane; i < n + N; i += 32) {
        float v = (i < n) ? s_z[i] : s_imp[i - n];
        int rank = 0;
        // concat index order: all coarse entries precede all importance entries
        for (int k = 0; k < n; ++k) {
          float o = s_z[k];
          rank += (o < v) || (o == v && k < i);
        }
        for (int k = 0; k < N; ++k) {
          float o = s_imp[k];
          rank += (o < v) || (o == v && (k + n) < i);
        }
        outr[rank] = v;
      }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kWarps * 32)
sort_merge_kernel(const float* __restrict__ a, const float* __restrict__ bb, float* __restrict__ out, int64_t B,
                  int na, int nb) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int tot = na + nb;
  float* s = smem + (size_t)wib * tot;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t r = warp0; r < B; r += nwarps) {
    for (int i = lane; i < tot; i += 32) s[i] = (i < na) ? a[r * na + i] : bb[r * nb + (i - na)];
    __syncwarp();
    for (int i = lane; i < tot; i += 32) {
      float v = s[i];
      int rank = 0;
      for (int k = 0; k < tot; ++k) {
        float o = s[k];
        rank += (o < v) || (o == v && k < i);
      }
      out[r * tot + rank] = v;
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int nmx_sample_pdf_fwd(const float* z, const float* weights, const float* u, const float* cdf_in,
                                  float eps, float* z_imp, int32_t* inds, float* cdf_out, float* z_merged,
                                  int64_t B, int n, int N, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 2 && N >= 1, "B >= 0, n >= 2, N >= 1");
  if (B == 0) return 0;
  NMX_CHECK_ARG(z && u && (weights || cdf_in), "z, u and one of weights / cdf_in must be non-null");
  size_t smem = (size_t)kWarps * (2 * (n + 1) + n + N) * sizeof(float);
  NMX_CHECK_ARG(smem <= 200 * 1024, "n + N too large for shared memory");
  if (smem > 48 * 1024)
    NMX_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = grid_for(B, kWarps, 16);
  sample_pdf_kernel<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(z, weights, u, cdf_in, eps, z_imp, inds,
                                                                         cdf_out, z_merged, B, n, N);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_sort_merge_z(const float* a, const float* b, float* out, int64_t B, int na, int nb, void* stream) {
  NMX_CHECK_ARG(B >= 0 && na >= 0 && nb >= 0 && na + nb >= 1, "B >= 0, na + nb >= 1");
  if (B == 0) return 0;
  size_t smem = (size_t)kWarps * (na + nb) * sizeof(float);
  NMX_CHECK_ARG(smem <= 200 * 1024, "na + nb too large for shared memory");
  if (smem > 48 * 1024)
    NMX_CUDA(cudaFuncSetAttribute(sort_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = grid_for(B, kWarps, 16);
  sort_merge_kernel<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(a, b, out, B, na, nb);
  NMX_LAUNCH_CHECK();
  return 0;
}
