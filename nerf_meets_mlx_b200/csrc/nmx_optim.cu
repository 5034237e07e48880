// K6: MSE loss (+ its gradient) and the Adam update of the reference's training step
// (entrypoints/__test_nerf.py:88,124,128-145; models/NeRF.py:120).  Pure streaming kernels, float4 vectorised.
#include "nmx_common.cuh"
#include "nmx_optim.cuh"

using namespace nmx;

namespace {

__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, float* __restrict__ loss,
           float* __restrict__ d_pred, int64_t count, float inv_count, float grad_scale) {
  float acc = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float d = pred[i] - target[i];
    acc += d * d;
    if (d_pred != nullptr) d_pred[i] = 2.0f * d * inv_count * grad_scale;
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s[w] = acc;
  __syncthreads();
  if (w == 0) {
    float v = (lane < (int)(blockDim.x >> 5)) ? s[lane] : 0.0f;
    v = warp_sum(v);
    if (lane == 0) atomicAdd(loss, v * inv_count);
  }
}

// MLX-0.7 Adam (no bias correction) or the standard bias-corrected form.  Seven streams (p, g, m, v in; p, m, v out):
// float4 accesses and two waves of 2048-thread SMs keep ~128 KB per SM in flight (the scalar version moved the hash
// grid's 64 MiB tables at 4.3 TB/s).  VEC = 0: unaligned operands (parameter sub-ranges), one float per access.
template <int VEC>
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t count, float lr, const float* __restrict__ lr_dev, float b1, float b2, float eps, float c1, float c2) {
  if (lr_dev != nullptr) lr = __ldg(lr_dev);  // learning rate from device memory (CUDA-graph replays)
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (VEC) {
    const int64_t n4 = count >> 2;
    for (int64_t i = tid; i < n4; i += nth) {
      float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
      const float4 gi = reinterpret_cast<const float4*>(g)[i];
      adam_update(pi.x, mi.x, vi.x, gi.x, lr, b1, b2, eps, c1, c2);
      adam_update(pi.y, mi.y, vi.y, gi.y, lr, b1, b2, eps, c1, c2);
      adam_update(pi.z, mi.z, vi.z, gi.z, lr, b1, b2, eps, c1, c2);
      adam_update(pi.w, mi.w, vi.w, gi.w, lr, b1, b2, eps, c1, c2);
      reinterpret_cast<float4*>(m)[i] = mi;
      reinterpret_cast<float4*>(v)[i] = vi;
      reinterpret_cast<float4*>(p)[i] = pi;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < count; i += nth) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, g[i], lr, b1, b2, eps, c1, c2);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t count, float lr, const float* lr_dev, float b1,
                float b2, float eps, float c1, float c2, cudaStream_t s) {
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  if (vec) adam_kernel<1><<<grid_for((count + 3) / 4, 256, 16), 256, 0, s>>>(p, g, m, v, count, lr, lr_dev, b1, b2, eps, c1, c2);
  else adam_kernel<0><<<grid_for(count, 256, 16), 256, 0, s>>>(p, g, m, v, count, lr, lr_dev, b1, b2, eps, c1, c2);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" int nmx_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* d_pred, int64_t count,
                               float grad_scale, void* stream) {
  NMX_CHECK_ARG(count >= 1 && pred && target && loss, "count >= 1; pred, target, loss non-null");
  int blocks = grid_for(count, 256, 2);
  mse_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, loss, d_pred, count, 1.0f / (float)count, grad_scale);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float b1,
                             float b2, float eps, int bias_correction, int64_t t, void* stream) {
  NMX_CHECK_ARG(count >= 0 && p && g && m && v, "count >= 0; p, g, m, v non-null");
  if (count == 0) return 0;
  float c1 = 1.0f, c2 = 1.0f;
  if (bias_correction) {
    c1 = 1.0f / (1.0f - powf(b1, (float)t));
    c2 = 1.0f / (1.0f - powf(b2, (float)t));
  }
  return launch_adam(p, g, m, v, count, lr, nullptr, b1, b2, eps, c1, c2, (cudaStream_t)stream);
}

extern "C" int nmx_adam_step_lrdev(float* p, const float* g, float* m, float* v, int64_t count, const float* lr_dev,
                                   float b1, float b2, float eps, void* stream) {
  NMX_CHECK_ARG(count >= 0 && p && g && m && v && lr_dev, "count >= 0; p, g, m, v, lr_dev non-null");
  if (count == 0) return 0;
  return launch_adam(p, g, m, v, count, 0.0f, lr_dev, b1, b2, eps, 1.0f, 1.0f, (cudaStream_t)stream);
}
