// K5: inverse-CDF importance resampling (sample_from_inverse_cdf_torch, sampling/__init__.py:101-178)
// fused with the sort-merge of coarse and importance depths (rendering/render.py:225, __test_nerf.py:288).
//
// One warp per ray.  The ray's CDF (n+1 edges), padded bin mid-points (n+1) and the N importance samples live in
// shared memory; every lane owns N/32 samples and runs a branch-free binary search over the shared CDF
// (searchsorted side="right" == number of edges <= u).  The merge is a stable sort of concat([z, z_imp]): the
// importance samples are sorted in registers by a warp-wide bitonic network (shuffles for partner distances < 32,
// compile-time register exchanges above), then both sorted runs find their output slots by binary search in the
// other run (ties: coarse entries first, as in the concatenation) and the merged row leaves through shared memory
// as coalesced stores.  Rows whose coarse depths are not ascending, and N > 256, take the O((n+N)^2) rank-counting
// path (any input gives sort(concat) exactly).
//
// CDF arithmetic: w = weights + 0.01 ; sum and running prefix in fp64 rounded to fp32
// (what torch-CPU cumsum does; both are exact in fp64 for <= 2^20 fp32 addends of this dynamic range, so the
// warp-parallel scan order cannot change the result) ; pdf = w / sum in fp32 ; cdf = min(1, prefix).
// HBM roofline: 1792 B/ray at n=64, N=128 (z 256 + w 256 + u 512 read, z_merged 768 write).
// Measured: 683 us per 262 144 rays = 10.5 % of the HBM peak, issue-bound at ~2000 warp instructions per ray.  The sort
// is NOT what costs: a variant that ranks the samples by their draws (uniform u -> 128 buckets, shared-memory counters,
// one warp scan; no sorting network, slots from the CDF bins) was bit-exact and ran at 714 us with the same ~2000
// instructions per ray (ncu: the R = 4 unrolled per-sample work -- IEEE divisions, clamps, gathers, slot arithmetic --
// and the fp64 CDF dominate, not the sorting network).  Kept simple; < 0.5 % of a training step.
#include "nmx_common.cuh"

using namespace nmx;

namespace {

constexpr int kWarps = 4;

// floats of shared memory per warp: cdf[n+1] mid[n+1] z[n] imp[N] out[n+N], every section start a multiple of 4 floats
// is NOT needed except for `out` (float4 stores) -- the total before it is padded to a multiple of 4
__host__ __device__ inline int smem_floats_per_warp(int n, int N) {
  const int head = 2 * (n + 1) + n + N;
  return ((head + 3) & ~3) + ((n + N + 3) & ~3);
}

// Ascending bitonic sort of 32*R values held as v[r] = element r*32 + lane.
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(float (&v)[R], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {  // partner lives in another register of the same lane
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & jr) == 0) {
            const bool asc = (((r * 32) & k) == 0);  // k >= 64 here: decided by the register index alone
            const float a = v[r], b = v[r | jr];
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            v[r] = asc ? lo : hi;
            v[r | jr] = asc ? hi : lo;
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool asc = (k >= 32) ? (((r * 32) & k) == 0) : ((lane & k) == 0);
          const float p = __shfl_xor_sync(0xffffffffu, v[r], j);
          v[r] = (lower == asc) ? fminf(v[r], p) : fmaxf(v[r], p);
        }
      }
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kWarps * 32)
sample_pdf_kernel(const float* __restrict__ z, const float* __restrict__ weights, const float* __restrict__ u,
                  const float* __restrict__ cdf_in, float eps, float* __restrict__ z_imp, int32_t* __restrict__ inds,
                  float* __restrict__ cdf_out, float* __restrict__ z_merged, int64_t B, int n, int N) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int per_warp = smem_floats_per_warp(n, N);
  float* s_cdf = smem + (size_t)wib * per_warp;  // [n+1]
  float* s_mid = s_cdf + (n + 1);                // [n+1] padded mid-points
  float* s_z = s_mid + (n + 1);                  // [n] coarse depths
  float* s_imp = s_z + n;                        // [N] importance samples (unsorted, then sorted)
  float* s_out = s_cdf + ((2 * (n + 1) + n + N + 3) & ~3);  // [n+N] merged row (16 B aligned), staged for coalesced stores
  const int64_t warp0 = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  const int n1 = n + 1;

  for (int64_t b = warp0; b < B; b += nwarps) {
    const float* zr = z + b * n;
    // ---- coarse depths + padded mid-points (:149-159): mid[0]=mid[1]=m_0 ... mid[n]=m_{n-2}
    bool z_sorted = true;
    for (int i = lane; i < n; i += 32) s_z[i] = zr[i];
    __syncwarp();
    for (int i = lane; i < n - 1; i += 32) z_sorted = z_sorted && (s_z[i] <= s_z[i + 1]);
    z_sorted = __all_sync(0xffffffffu, z_sorted);
    for (int i = lane; i < n1; i += 32) {
      int j = min(max(i - 1, 0), n - 2);
      s_mid[i] = __fmul_rn(__fadd_rn(s_z[j + 1], s_z[j]), 0.5f);  // == / 2 exactly (power of two), without the IEEE divide
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

    // ---- stable sort of concat([z (n), imp (N)])
    if (z_merged != nullptr) {
      const int tot = n + N;
      if (R > 0 && z_sorted) {
        float v[R > 0 ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = (r * 32 + lane < N) ? s_imp[r * 32 + lane] : __int_as_float(0x7f800000);
        __syncwarp();
        warp_bitonic_sort<(R > 0 ? R : 1)>(v, lane);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r * 32 + lane < N) s_imp[r * 32 + lane] = v[r];
        __syncwarp();
        // importance sample e lands at e + #(coarse <= v): coarse entries come first among equals
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = r * 32 + lane;
          if (e < N) {
            int lo = 0, hi = n;
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (s_z[mid] <= v[r]) lo = mid + 1; else hi = mid;
            }
            s_out[e + lo] = v[r];
          }
        }
        // coarse entry i lands at i + #(importance < z_i)
        for (int i = lane; i < n; i += 32) {
          const float zv = s_z[i];
          int lo = 0, hi = N;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_imp[mid] < zv) lo = mid + 1; else hi = mid;
          }
          s_out[i + lo] = zv;
        }
      } else {
        // rank counting (rank = #smaller + #equal-with-lower-concat-index): any input order
        for (int i = lane; i < tot; i += 32) {
          float v = (i < n) ? s_z[i] : s_imp[i - n];
          int rank = 0;
          for (int k = 0; k < n; ++k) {
            float o = s_z[k];
            rank += (o < v) || (o == v && k < i);
          }
          for (int k = 0; k < N; ++k) {
            float o = s_imp[k];
            rank += (o < v) || (o == v && (k + n) < i);
          }
          s_out[rank] = v;
        }
      }
      __syncwarp();
      float* outr = z_merged + b * (int64_t)tot;
      if ((tot & 3) == 0) {
        for (int i = lane * 4; i < tot; i += 128)
          *reinterpret_cast<float4*>(outr + i) = *reinterpret_cast<const float4*>(s_out + i);
      } else {
        for (int i = lane; i < tot; i += 32) outr[i] = s_out[i];
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
  size_t smem = (size_t)kWarps * smem_floats_per_warp(n, N) * sizeof(float);
  NMX_CHECK_ARG(smem <= 200 * 1024, "n + N too large for shared memory");
  int blocks = grid_for(B, kWarps, 16);
  auto go = [&](auto kern) -> int {
    if (smem > 48 * 1024) NMX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, kWarps * 32, smem, (cudaStream_t)stream>>>(z, weights, u, cdf_in, eps, z_imp, inds, cdf_out, z_merged, B, n, N);
    NMX_LAUNCH_CHECK();
    return 0;
  };
  // registers per lane for the bitonic network: 32*R >= N (R = 0: rank-counting path only)
  if (N <= 32) return go(sample_pdf_kernel<1>);
  if (N <= 64) return go(sample_pdf_kernel<2>);
  if (N <= 128) return go(sample_pdf_kernel<4>);
  if (N <= 256) return go(sample_pdf_kernel<8>);
  return go(sample_pdf_kernel<0>);
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
