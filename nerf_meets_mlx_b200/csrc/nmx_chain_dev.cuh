// Device helpers shared by the fused MLP chain kernels (nmx_chain.cu: one tile per CTA; nmx_chain2.cu: CTA pairs with two
// tiles in ping-pong): packed fp32x2 arithmetic, bf16x2 conversion with folded ReLU, shared-memory vector accesses,
// TMEM loads with a register-naming wait, the branch-free sin/cos and the Embedder positional encoding of one row.
#pragma once
#include <cuda_bf16.h>
#include "nmx_sm100.cuh"

namespace nmx {
namespace chain_dev {
using namespace nmx::sm100;

constexpr int kChunkBytesDev = 128 * 64 * 2;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// packed fp32x2 add (sm_100: one FADD2 for two columns) and fp32x2 -> bf16x2 conversion with the ReLU folded in
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_bf16x2(uint64_t v) {  // low half <- low float, high half <- high float
  uint32_t lo, hi, d;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return d;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 0xFFFF per half whose bf16 value is non-zero (post-ReLU activation > 0)
__device__ __forceinline__ uint32_t nz_mask_bf16x2(uint32_t v) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
  return __hne2_mask(*reinterpret_cast<const __nv_bfloat162*>(&v), z);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// read-only staging data (biases / head weights written once before the first __syncthreads): schedulable loads
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int C>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* r) {
  if constexpr (C == 32) tmem_ld_32x32(taddr, r);
  else tmem_ld_32x16(taddr, r);
}
// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled above it
template <int C>
__device__ __forceinline__ void tmem_ld_wait_regs(uint32_t* r) {
  if constexpr (C == 16) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
  } else {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
  }
}

// Branch-free sin/cos for the fused encoder: three-term Cody-Waite reduction by pi/2 and the usual degree-7/8 minimax
// polynomials on [-pi/4, pi/4] (about 1 ulp for |x| < ~1e4; the encoder's arguments are at most 81 * |pos|).  Unlike
// sincosf it has no slow path, so the compiler can interleave the 30 evaluations of a row; the result is rounded to
// bf16 right away, which hides the (<= 1 ulp fp32) difference from sincosf.
__device__ __forceinline__ void sincos_bf(float x, float* sp, float* cp) {
  const float k = rintf(x * 0.636619747f);
  float r = fmaf(-k, 1.57079601e+00f, x);
  r = fmaf(-k, 3.13916473e-07f, r);
  r = fmaf(-k, 5.39030253e-15f, r);
  const float z = r * r;
  float ps = 2.86567956e-6f;
  ps = fmaf(ps, z, -1.98559923e-4f);
  ps = fmaf(ps, z, 8.33338592e-3f);
  ps = fmaf(ps, z, -1.66666672e-1f);
  float s = fmaf(ps * z, r, r);
  float pc = 2.44677067e-5f;
  pc = fmaf(pc, z, -1.38877297e-3f);
  pc = fmaf(pc, z, 4.16666567e-2f);
  pc = fmaf(pc, z, -5.00000000e-1f);
  float c = fmaf(pc, z, 1.0f);
  const int q = (int)k;
  const float s2 = (q & 1) ? c : s;
  const float c2 = (q & 1) ? s : c;
  *sp = (q & 2) ? -s2 : s2;
  *cp = ((q + 1) & 2) ? -c2 : c2;
}

// Fused input encoding of one point (enc_kind 1, 10 position bands): pos = o + z d with separate mul / add and band
// k^2 exactly as encode_rays_kernel, sin/cos by sincos_bf, written as the 128 B swizzled row of the position chunk.  Channel order (models/embedding.py:35-71): [x y z | per band k: sin(k^2 xyz), cos(k^2 xyz)] + 0 pad.
__device__ __forceinline__ void encode_pos_row(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z,
                                               long long p, int n, bool valid, uint32_t row_addr, uint32_t swz) {
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 0.0f;
  if (valid) {
    const long long b = p / n;
    const float* r = rays + b * ray_stride;
    const float zz = __ldg(z + p);
    const float px = __fadd_rn(__ldg(r + 0), __fmul_rn(zz, __ldg(r + 3)));
    const float py = __fadd_rn(__ldg(r + 1), __fmul_rn(zz, __ldg(r + 4)));
    const float pz = __fadd_rn(__ldg(r + 2), __fmul_rn(zz, __ldg(r + 5)));
    v[0] = px; v[1] = py; v[2] = pz;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const float f = (float)(k * k);
      sincos_bf(__fmul_rn(px, f), &v[3 + 6 * k + 0], &v[3 + 6 * k + 3]);
      sincos_bf(__fmul_rn(py, f), &v[3 + 6 * k + 1], &v[3 + 6 * k + 4]);
      sincos_bf(__fmul_rn(pz, f), &v[3 + 6 * k + 2], &v[3 + 6 * k + 5]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts128(row_addr + ((((uint32_t)j) ^ swz) << 4), pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
           pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
}


// Epilogue of one layer for one warp: for every 64-column chunk, this warp's kEpiCols columns x 32 rows:
// TMEM -> (+bias, ReLU) -> bf16 -> swizzled smem chunk (A operand of the next layer / TMA-store source),
// optional register heads.  HEAD: 0 none, 1 alpha (one output), 2 rgb, 3 output_linear (up to 8 outputs).
// The TMEM load of chunk c+1 is issued before the math of chunk c (two register sets).
// math + store of 32 columns [c0, c0 + 32) of chunk c held in r[] (fp32 accumulators of this thread's row)
// NP4 = number of 8-column pieces (4: 32 columns, 2: 16 columns); piece0 = index of the first 16-byte piece inside the
// chunk's 128-byte row; bit0 = first pair index of these columns inside their 32-column sign-bit word.
// FENCE: end with the generic->async proxy fence (callers that issue their next TMEM load first pass false)
template <bool RELU, int HEAD, int NP4, bool BITS, bool FENCE = (NP4 == 4)>
__device__ __forceinline__ void epi_cols(const uint32_t (&r)[8 * NP4], const int c, const int c0, const int piece0,
                                         const int bit0, const uint32_t bias_addr, const uint32_t act_row_addr,
                                         const uint32_t swz, const uint32_t hw_addr, const int head_n, float (&hp)[8],
                                         float (&rgbp)[3], const uint32_t bits_addr) {
  auto ldb = [&](int col) -> float4 { return lds128(bias_addr + (uint32_t)col * 4u); };
  auto ldw = [&](int idx) -> float4 { return lds128(hw_addr + (uint32_t)idx * 4u); };
  const uint32_t so = act_row_addr + (uint32_t)c * kChunkBytesDev;
  uint32_t bits = 0;  // ReLU sign bits of this thread's columns (bit e / 16+e = columns 2e / 2e+1 of the 32-column word)
#pragma unroll
  for (int p4 = 0; p4 < NP4; ++p4) {
    const float4 b0 = ldb(c0 + p4 * 8);
    const float4 b1 = ldb(c0 + p4 * 8 + 4);
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // bf16(relu(acc + bias)): fp32 add, ReLU folded into the rounding conversion
      const uint64_t x = add_f32x2(pack64(r[p4 * 8 + 2 * e], r[p4 * 8 + 2 * e + 1]),
                                   pack64(__float_as_uint(bv[2 * e]), __float_as_uint(bv[2 * e + 1])));
      pk[e] = cvt_bf16x2<RELU>(x);
      if (RELU && BITS) bits |= nz_mask_bf16x2(pk[e]) & (0x00010001u << (bit0 + p4 * 4 + e));
    }
    sts128(so + ((((uint32_t)(piece0 + p4)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
    if (HEAD != 0) {
      float xr[8];  // the bf16-rounded activations (what a separate head kernel would read back)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xr[2 * e] = __uint_as_float(pk[e] << 16);
        xr[2 * e + 1] = __uint_as_float(pk[e] & 0xffff0000u);
      }
      const int col = c0 + p4 * 8;
      if (HEAD == 1 || HEAD == 3) {  // 1: single output (alpha); 3: up to 8 outputs (output_linear)
#pragma unroll
        for (int o = 0; o < (HEAD == 1 ? 1 : 8); ++o) {
          if (HEAD == 1 || o < head_n) {
            const float4 w0 = ldw(o * 256 + col);
            const float4 w1 = ldw(o * 256 + col + 4);
            hp[o] += xr[0] * w0.x + xr[1] * w0.y + xr[2] * w0.z + xr[3] * w0.w + xr[4] * w1.x + xr[5] * w1.y +
                     xr[6] * w1.z + xr[7] * w1.w;
          }
        }
      } else {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          const float4 w0 = ldw(o * 128 + col);
          const float4 w1 = ldw(o * 128 + col + 4);
          rgbp[o] += xr[0] * w0.x + xr[1] * w0.y + xr[2] * w0.z + xr[3] * w0.w + xr[4] * w1.x + xr[5] * w1.y +
                     xr[6] * w1.z + xr[7] * w1.w;
        }
      }
    }
  }
  if (RELU && BITS && bits_addr != 0) {
    if (NP4 == 4) {
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(bits_addr), "r"(bits) : "memory");
    } else {  // half a word: pairs bit0..bit0+7 live in byte bit0/8 (even columns) and byte 2 + bit0/8 (odd columns)
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(bits_addr + (uint32_t)(bit0 >> 3)), "r"((bits >> bit0) & 0xffu) : "memory");
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(bits_addr + 2u + (uint32_t)(bit0 >> 3)), "r"((bits >> (16 + bit0)) & 0xffu) : "memory");
    }
  }
  if (FENCE) fence_proxy_async_smem();
}


}  // namespace chain_dev
}  // namespace nmx
