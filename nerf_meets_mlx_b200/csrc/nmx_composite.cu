// K4: alpha compositing (raw2outputs, rendering/render.py:20-96) forward and backward.
// One warp per ray; samples are laid out lane-contiguous in chunks of 32 so that every global access is a
// coalesced 128 B (z, weights) or 512 B (raw as float4) warp transaction; prefix sums are warp shuffles.
//
// Reference semantics replicated:
//   delta_i = z_{i+1}-z_i (last = 1e10), * ||d|| ; tau_i = delta_i*sigma_i
//   alpha_i = 1 - exp(-relu(tau_i)) ; T_i = exp(-sum_{j<i} tau_j)   (RAW tau in T -- not relu'd)
//   w_i = alpha_i*T_i ; rgb = sum w c (no sigmoid) ; depth = sum w z ; acc = sum w
//   disp = 1/max(1e-10, depth/acc) ; rgb += 1-acc if white_bkgd
// HBM roofline (algorithmic bytes / ray): fwd 24n+36, bwd 36n+28.
#include "nmx_common.cuh"

using namespace nmx;

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float ray_norm(const float* d) {
  float x = d[0], y = d[1], zz = d[2];
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(zz, zz)));
}

// loads chunk c of a ray: z_i, delta_i (scaled by norm), raw float4 (sigma with optional noise)
__device__ __forceinline__ void load_chunk(const float* __restrict__ zr, const float4* __restrict__ rawr,
                                           const float* __restrict__ noiser, float noise_std, int n, int c, int lane,
                                           float norm, float& zi, float& delta, float4& rv, bool& valid) {
  int i = c * 32 + lane;
  valid = i < n;
  zi = valid ? zr[i] : 0.0f;
  float znext = __shfl_down_sync(0xffffffffu, zi, 1);
  if (lane == 31) {
    int j = i + 1;
    znext = (j < n) ? zr[j] : 0.0f;
  }
  float d = (i == n - 1) ? 1e10f : __fsub_rn(znext, zi);
  delta = __fmul_rn(d, norm);
  rv = valid ? rawr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  if (noiser != nullptr && valid) rv.w = __fadd_rn(rv.w, __fmul_rn(noiser[i], noise_std));
  if (!valid) delta = 0.0f;
}

// One chunk of 32 samples at a time (32 registers, full occupancy).  Issuing all of a ray's loads up front from
// register arrays was measured slower on B200 (n = 192: 116 us vs 81 us per 65536 rays): occupancy wins here.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int d_stride, const float* __restrict__ noise, float noise_std, int white_bkgd,
                     float* __restrict__ rgb, float* __restrict__ disp, float* __restrict__ acc,
                     float* __restrict__ weights, float* __restrict__ depth, int64_t B, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int nchunks = (n + 31) >> 5;
  for (int64_t b = warp0; b < B; b += nwarps) {
    const float* zr = z + b * n;
    const float4* rawr = reinterpret_cast<const float4*>(raw) + b * n;
    const float* noiser = noise ? noise + b * n : nullptr;
    float norm = ray_norm(rays_d + b * d_stride);
    float carry = 0.0f;
    float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aa = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      float zi, delta;
      float4 rv;
      bool valid;
      load_chunk(zr, rawr, noiser, noise_std, n, c, lane, norm, zi, delta, rv, valid);
      float tau = __fmul_rn(delta, rv.w);
      // exclusive prefix must not see the (huge) last-bin tau of the final sample: mask it out of the scan
      int i = c * 32 + lane;
      float tau_scan = (valid && i < n - 1) ? tau : 0.0f;
      float incl = warp_scan_incl(tau_scan, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.0f;
      float S = carry + excl;
      float T = expf(-S);
      float alpha = 1.0f - expf(-fmaxf(tau, 0.0f));
      float w = valid ? alpha * T : 0.0f;
      if (weights != nullptr && valid) weights[b * n + i] = w;
      ar += w * rv.x;
      ag += w * rv.y;
      ab += w * rv.z;
      ad += w * zi;
      aa += w;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    ar = warp_sum(ar);
    ag = warp_sum(ag);
    ab = warp_sum(ab);
    ad = warp_sum(ad);
    aa = warp_sum(aa);
    if (lane == 0) {
      if (white_bkgd) {
        float bg = 1.0f - aa;
        ar += bg;
        ag += bg;
        ab += bg;
      }
      if (rgb) {
        rgb[b * 3 + 0] = ar;
        rgb[b * 3 + 1] = ag;
        rgb[b * 3 + 2] = ab;
      }
      if (depth) depth[b] = ad;
      if (acc) acc[b] = aa;
      if (disp) disp[b] = 1.0f / fmaxf(1e-10f, ad / aa);
    }
  }
}

// Backward.  NCHUNK = ceil(n/32) register-resident chunks (n <= 32*NCHUNK).
//   g_i   = dL/dw_i = d_rgb.c_i + d_depth' z_i + d_acc' + d_weights_i
//   dL/dtau_i = g_i T_i exp(-relu(tau_i)) [tau_i>0]  -  sum_{k>i} g_k w_k     (T path uses raw tau)
//   dL/dsigma_i = delta_i dL/dtau_i ;  dL/dc_i = w_i d_rgb
// The tau of the last sample never enters any T, so its suffix term is empty.
template <int NCHUNK>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                     int d_stride, const float* __restrict__ noise, float noise_std, int white_bkgd,
                     const float* __restrict__ d_rgb, const float* __restrict__ d_disp,
                     const float* __restrict__ d_acc, const float* __restrict__ d_depth,
                     const float* __restrict__ d_weights, float* __restrict__ d_raw, int64_t B, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t b = warp0; b < B; b += nwarps) {
    const float* zr = z + b * n;
    const float4* rawr = reinterpret_cast<const float4*>(raw) + b * n;
    const float* noiser = noise ? noise + b * n : nullptr;
    float norm = ray_norm(rays_d + b * d_stride);
    float gr = d_rgb[b * 3 + 0], gg = d_rgb[b * 3 + 1], gb = d_rgb[b * 3 + 2];
    float g_acc = d_acc ? d_acc[b] : 0.0f;
    float g_depth = d_depth ? d_depth[b] : 0.0f;
    if (white_bkgd) g_acc -= (gr + gg + gb);

    float w_[NCHUNK], T_[NCHUNK], tau_[NCHUNK], delta_[NCHUNK], zi_[NCHUNK];
    float4 rv_[NCHUNK];
    float carry = 0.0f, aa = 0.0f, ad = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      bool valid;
      load_chunk(zr, rawr, noiser, noise_std, n, c, lane, norm, zi_[c], delta_[c], rv_[c], valid);
      int i = c * 32 + lane;
      float tau = __fmul_rn(delta_[c], rv_[c].w);
      float tau_scan = (valid && i < n - 1) ? tau : 0.0f;
      float incl = warp_scan_incl(tau_scan, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.0f;
      float S = carry + excl;
      float T = expf(-S);
      float alpha = 1.0f - expf(-fmaxf(tau, 0.0f));
      float w = valid ? alpha * T : 0.0f;
      tau_[c] = tau;
      T_[c] = valid ? T : 0.0f;
      w_[c] = w;
      aa += w;
      ad += w * zi_[c];
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (d_disp != nullptr) {  // disp = 1/max(1e-10, depth/acc)
      float acc_t = warp_sum(aa), dep_t = warp_sum(ad);
      float q = dep_t / acc_t;
      if (q > 1e-10f) {
        float dq = -d_disp[b] / (q * q);
        g_depth += dq / acc_t;
        g_acc += -dq * dep_t / (acc_t * acc_t);
      }
    }
    // suffix sums of g_k w_k, chunks in reverse order
    float rcarry = 0.0f;
#pragma unroll
    for (int c = NCHUNK - 1; c >= 0; --c) {
      int i = c * 32 + lane;
      bool valid = i < n;
      float g = gr * rv_[c].x + gg * rv_[c].y + gb * rv_[c].z + g_depth * zi_[c] + g_acc;
      if (d_weights != nullptr && valid) g += d_weights[b * n + i];
      float gw = valid ? g * w_[c] : 0.0f;
      float rincl = warp_rscan_incl(gw, lane);
      float rexcl = __shfl_down_sync(0xffffffffu, rincl, 1);
      if (lane == 31) rexcl = 0.0f;
      float suffix_excl = rcarry + rexcl;
      float tau = tau_[c];
      float dalpha = (tau > 0.0f) ? expf(-tau) : 0.0f;  // d alpha / d tau
      float dtau = g * T_[c] * dalpha;
      if (i < n - 1) dtau -= suffix_excl;  // last sample's tau feeds no transmittance
      float dsigma = delta_[c] * dtau;
      if (valid) {
        float wv = w_[c];
        reinterpret_cast<float4*>(d_raw)[b * n + i] = make_float4(wv * gr, wv * gg, wv * gb, dsigma);
      }
      rcarry += __shfl_sync(0xffffffffu, rincl, 0);
    }
  }
}

// Training step in one pass over a ray (the loss functions of the reference, __test_nerf.py:83-88 and :111-124):
//   rgb = raw2outputs(raw, z, d)[0] ; loss += mean((rgb - target)^2) ; d_raw = d loss / d raw.
// Same forward and backward arithmetic as the two kernels above; the per-ray colour never goes to HBM and the loss
// gradient 2 (rgb - target) / (3 B) is formed in registers.  Optional outputs: rgb [B,3], weights [B,n].
template <int NCHUNK>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_loss_kernel(const float* __restrict__ raw, const float* __restrict__ z, const float* __restrict__ rays_d,
                      int d_stride, int white_bkgd, const float* __restrict__ target, float inv_count,
                      float* __restrict__ loss, float* __restrict__ d_raw, float* __restrict__ rgb_out,
                      float* __restrict__ weights, int64_t B, int n) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float loss_acc = 0.0f;  // this warp's share of sum (rgb - target)^2 (lane 0)
  for (int64_t b = warp0; b < B; b += nwarps) {
    const float* zr = z + b * n;
    const float4* rawr = reinterpret_cast<const float4*>(raw) + b * n;
    float norm = ray_norm(rays_d + b * d_stride);
    float w_[NCHUNK], T_[NCHUNK], tau_[NCHUNK], delta_[NCHUNK];
    float4 rv_[NCHUNK];
    float carry = 0.0f, aa = 0.0f, ar = 0.0f, ag = 0.0f, ab = 0.0f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      bool valid;
      float zi;
      load_chunk(zr, rawr, nullptr, 0.0f, n, c, lane, norm, zi, delta_[c], rv_[c], valid);
      int i = c * 32 + lane;
      float tau = __fmul_rn(delta_[c], rv_[c].w);
      float tau_scan = (valid && i < n - 1) ? tau : 0.0f;
      float incl = warp_scan_incl(tau_scan, lane);
      float excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) excl = 0.0f;
      float S = carry + excl;
      float T = expf(-S);
      float alpha = 1.0f - expf(-fmaxf(tau, 0.0f));
      float w = valid ? alpha * T : 0.0f;
      if (weights != nullptr && valid) weights[b * n + i] = w;
      tau_[c] = tau;
      T_[c] = valid ? T : 0.0f;
      w_[c] = w;
      ar += w * rv_[c].x;
      ag += w * rv_[c].y;
      ab += w * rv_[c].z;
      aa += w;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    ar = warp_sum(ar);
    ag = warp_sum(ag);
    ab = warp_sum(ab);
    aa = warp_sum(aa);
    if (white_bkgd) {
      const float bg = 1.0f - aa;
      ar += bg; ag += bg; ab += bg;
    }
    const float dr = ar - target[b * 3 + 0], dg = ag - target[b * 3 + 1], db = ab - target[b * 3 + 2];
    if (lane == 0) {
      loss_acc += dr * dr + dg * dg + db * db;
      if (rgb_out != nullptr) { rgb_out[b * 3 + 0] = ar; rgb_out[b * 3 + 1] = ag; rgb_out[b * 3 + 2] = ab; }
    }
    const float gr = 2.0f * dr * inv_count, gg = 2.0f * dg * inv_count, gb = 2.0f * db * inv_count;
    const float g_acc = white_bkgd ? -(gr + gg + gb) : 0.0f;
    float rcarry = 0.0f;
#pragma unroll
    for (int c = NCHUNK - 1; c >= 0; --c) {
      int i = c * 32 + lane;
      bool valid = i < n;
      float g = gr * rv_[c].x + gg * rv_[c].y + gb * rv_[c].z + g_acc;
      float gw = valid ? g * w_[c] : 0.0f;
      float rincl = warp_rscan_incl(gw, lane);
      float rexcl = __shfl_down_sync(0xffffffffu, rincl, 1);
      if (lane == 31) rexcl = 0.0f;
      float suffix_excl = rcarry + rexcl;
      float tau = tau_[c];
      float dalpha = (tau > 0.0f) ? expf(-tau) : 0.0f;
      float dtau = g * T_[c] * dalpha;
      if (i < n - 1) dtau -= suffix_excl;
      float dsigma = delta_[c] * dtau;
      if (valid) {
        float wv = w_[c];
        reinterpret_cast<float4*>(d_raw)[b * n + i] = make_float4(wv * gr, wv * gg, wv * gb, dsigma);
      }
      rcarry += __shfl_sync(0xffffffffu, rincl, 0);
    }
  }
  // one atomic per block
  __shared__ float s_loss[kWarpsPerBlock];
  if (lane == 0) s_loss[threadIdx.x >> 5] = loss_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kWarpsPerBlock; ++w) t += s_loss[w];
    atomicAdd(loss, t * inv_count);
  }
}

}  // namespace

extern "C" int nmx_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 const float* noise, float raw_noise_std, int white_bkgd, float* rgb, float* disp,
                                 float* acc, float* weights, float* depth, int64_t B, int n, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 1 && d_stride >= 3, "B >= 0, n >= 1, d_stride >= 3");
  if (B == 0) return 0;
  NMX_CHECK_ARG(raw && z && rays_d, "raw, z, rays_d must be non-null");
  if (raw_noise_std <= 0.0f) noise = nullptr;
  int blocks = grid_for(B, kWarpsPerBlock, 8);
  composite_fwd_kernel<<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      raw, z, rays_d, d_stride, noise, raw_noise_std, white_bkgd, rgb, disp, acc, weights, depth, B, n);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                 const float* noise, float raw_noise_std, int white_bkgd, const float* d_rgb,
                                 const float* d_disp, const float* d_acc, const float* d_depth,
                                 const float* d_weights, float* d_raw, int64_t B, int n, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 1 && d_stride >= 3, "B >= 0, n >= 1, d_stride >= 3");
  if (B == 0) return 0;
  NMX_CHECK_ARG(raw && z && rays_d && d_rgb && d_raw, "raw, z, rays_d, d_rgb, d_raw must be non-null");
  NMX_CHECK_ARG(n <= 256, "n <= 256");
  if (raw_noise_std <= 0.0f) noise = nullptr;
  int blocks = grid_for(B, kWarpsPerBlock, 8);
  cudaStream_t s = (cudaStream_t)stream;
  int nch = (n + 31) / 32;
#define NMX_BWD(NC)                                                                                          \
  composite_bwd_kernel<NC><<<blocks, kWarpsPerBlock * 32, 0, s>>>(raw, z, rays_d, d_stride, noise, raw_noise_std, \
                                                                   white_bkgd, d_rgb, d_disp, d_acc, d_depth,   \
                                                                   d_weights, d_raw, B, n)
  if (nch <= 1) NMX_BWD(1);
  else if (nch <= 2) NMX_BWD(2);
  else if (nch <= 4) NMX_BWD(4);
  else if (nch <= 6) NMX_BWD(6);
  else NMX_BWD(8);
#undef NMX_BWD
  NMX_LAUNCH_CHECK();
  return 0;
}

// Fused training pass of one net: raw2outputs forward, MSE against `target` [B,3] (loss [1] is ACCUMULATED: the caller
// zeroes it; mean over 3 B elements) and the gradient w.r.t. raw [B,n,4], one kernel.  rgb [B,3] and weights [B,n] are
// optional outputs (NULL to skip).  n <= 256.
extern "C" int nmx_composite_loss_fwd_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                                          int white_bkgd, const float* target, float* loss, float* d_raw, float* rgb,
                                          float* weights, int64_t B, int n, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 1 && d_stride >= 3 && n <= 256, "B >= 0, 1 <= n <= 256, d_stride >= 3");
  if (B == 0) return 0;
  NMX_CHECK_ARG(raw && z && rays_d && target && loss && d_raw, "raw, z, rays_d, target, loss, d_raw must be non-null");
  int blocks = grid_for(B, kWarpsPerBlock, 8);
  cudaStream_t s = (cudaStream_t)stream;
  const float inv_count = 1.0f / (float)(3 * B);
  int nch = (n + 31) / 32;
#define NMX_CL(NC)                                                                                                  \
  composite_loss_kernel<NC><<<blocks, kWarpsPerBlock * 32, 0, s>>>(raw, z, rays_d, d_stride, white_bkgd, target, \
                                                                    inv_count, loss, d_raw, rgb, weights, B, n)
  if (nch <= 1) NMX_CL(1);
  else if (nch <= 2) NMX_CL(2);
  else if (nch <= 4) NMX_CL(4);
  else if (nch <= 6) NMX_CL(6);
  else NMX_CL(8);
#undef NMX_CL
  NMX_LAUNCH_CHECK();
  return 0;
}
