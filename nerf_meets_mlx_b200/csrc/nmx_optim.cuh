// The Adam update shared by the single-GPU optimiser kernel (nmx_optim.cu) and the fused all-reduce + Adam kernel
// (nmx_comm.cu), so that both produce bit-identical parameters from the same gradient.
#pragma once

namespace nmx {

// optim.Adam of MLX 0.7.0 (models/NeRF.py:120): m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
// p -= lr * (m c1) / (sqrt(v c2) + eps)   (c1 = c2 = 1: no bias correction, as MLX 0.7)
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, const float g, const float lr, const float b1,
                                            const float b2, const float eps, const float c1, const float c2) {
  const float mi = b1 * m + (1.0f - b1) * g;
  const float vi = b2 * v + (1.0f - b2) * g * g;
  m = mi;
  v = vi;
  p = p - lr * (mi * c1) / (sqrtf(vi * c2) + eps);
}

}  // namespace nmx
