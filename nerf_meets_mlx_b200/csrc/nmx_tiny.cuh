// Internal C++ API of the fully fused width-64 MLP (nmx_tiny.cu), used by nmx_mlp.cu.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmx {

constexpr int kTinyMaxLayers = 4;

// A width-64 MLP without view-dir head or skip connection: in_pos -> 64 -> ... -> 64 -> out_ch, ReLU after every trunk
// layer, parameters read straight from the packed fp32 vector (W[out][in] row-major, then the bias, per Linear).
struct TinyMlpDesc {
  const float* params;
  int64_t w_off[kTinyMaxLayers], b_off[kTinyMaxLayers];  // trunk layers 0 .. D-1
  int64_t wo_off, bo_off;                                // output layer [out_ch][64]
  int D, in_pos, out_ch;                                 // 1 <= D <= 4, in_pos in {32, 64}, 1 <= out_ch <= 8
  int64_t P;
};

// out[P, out_ch] = MLP(x[P, in_pos]); x0_save (optional, training): the bf16 copy of x the backward pass recomputes from,
// row stride ldx0 elements.
int launch_tiny_fwd(const TinyMlpDesc& d, const float* x, float* out, __nv_bfloat16* x0_save, int ldx0, cudaStream_t s);
// d_params += all weight / bias gradients (atomics; the caller zeroes it), d_input[P, in_pos] (optional) = dL/dx.
int launch_tiny_bwd(const TinyMlpDesc& d, const __nv_bfloat16* x0, int ldx0, const float* d_out, float* d_params,
                    float* d_input, cudaStream_t s);

}  // namespace nmx
