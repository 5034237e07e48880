// K1t: the fully fused width-64 MLP -- the "tiny MLP" behind a hash-grid encoding (config C4: 32 -> 64 -> 64 -> 4).
//
// What it mirrors: NeRF.forward (models/NeRF.py:201-243) of a net without view-dir head or skip connection, and its
// autograd backward -- the same arithmetic as the per-layer path of nmx_mlp.cu (bf16 operands, fp32 accumulation, fp32
// bias added before the single bf16 rounding, fp32 output head, ReLU derivative from the stored bf16 activation).
//
// Why a separate kernel: at width 64 the per-layer path is a dozen launches that each stream a [P, 64] tensor through HBM
// (262 144 points: ~220 us of the 620 us C4 step).  The whole net is 24 KB of bf16 weights and 13 KFLOP per point, so
// the op is far from the tcgen05 roofline either way; this version is a register-resident chain of warp-level mma.sync
// m16n8k16 tiles (no TMEM round trip per K = N = 64 layer).  Measured (profiles/r2_tiny_mlp_ncu.txt): forward 20 us,
// backward 57 us per 262 144 points, HMMA pipe 40 / 46 % busy -- the warp-MMA path itself is the next bound, so the
// follow-up is the same chain on tcgen05 tiles of 128 points.  One warp owns 16 points; the fp32 accumulator fragment of layer l IS the bf16 A fragment of layer l+1
// after ReLU + packing, so activations never leave registers.
//
//   forward : x[P, in] fp32 -> out[P, out_ch] fp32; training additionally keeps the bf16 copy of x (64 B/point) -- the
//             ONLY tensor saved.  208 B/point of HBM traffic instead of ~900.
//   backward: recomputes h_0 .. h_{D-1} from the saved bf16 input (bit-identical to the forward: same MMAs), runs the
//             data-gradient chain d_out -> dY_{D-1} -> ... -> dY_0 -> d_x in registers, and stages dY_l / h_l of the
//             CTA's 128 points in shared memory, where all 8 warps contract them into the weight gradients
//             (dW_l = dY_l^T h_{l-1}: ldmatrix.trans operands, K = 128 points per step, bias gradients through a ones
//             operand, the fp32 head gradient through a hi + lo bf16 split of d_out).  The accumulators stay in
//             registers across the CTA's tiles and are flushed once with vector atomics.  272 B/point of HBM traffic.
//
// Column permutations: an MMA contracts over k in any order as long as A and B agree, and the n index of the output is
// whatever B says it is.  The packed weight fragments are built so that a lane's 8 input values of layer 0 (and its 8
// output values of d_x) are CONTIGUOUS in memory: 32-byte loads / stores per lane, 128 B per point row per quad.
#include "nmx_common.cuh"
#include "nmx_tiny.cuh"

using namespace nmx;
typedef __nv_bfloat16 bf16;

namespace {

constexpr int kTileStride = 72;   // bf16 elements per row of a [128][64] shared-memory tile (+8: conflict-free ldmatrix)
constexpr uint32_t kOnes = 0x3f803f80u;  // two bf16 1.0

struct TinyArgs {
  const float* params;
  int64_t w_off[kTinyMaxLayers], b_off[kTinyMaxLayers];
  int64_t wo_off, bo_off;
  int out_ch;
  int64_t P;
};

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- weight fragments in shared memory: entry ((ks * NT + nt) * 32 + lane) * 2 + r is register r of the B fragment of
// k-step ks, n-tile nt for that lane (g = lane / 4 -> n, tq = lane % 4 -> k pair; r = 0: k low half, 1: k + 8).
// forward operand B[k][n] = W[n][k]; PERM: layer 0, lane tq's 8 k's of a 32-column block are columns tq*8 .. tq*8+7
template <int K, bool PERM>
__device__ __forceinline__ void pack_fwd(uint32_t* dst, const float* __restrict__ W) {
#pragma unroll
  for (int it = 0; it < (K / 16) * 2; ++it) {  // 256 threads, all loads of a layer in flight at once
    const int i = it * 256 + threadIdx.x;
    const int r = i & 1, lane = (i >> 1) & 31, nt = (i >> 6) & 7, ks = i >> 9;
    const int g = lane >> 2, tq = lane & 3;
    const int n = nt * 8 + g;
    const int k0 = PERM ? 32 * (ks >> 1) + tq * 8 + (ks & 1) * 4 + r * 2 : ks * 16 + tq * 2 + 8 * r;
    dst[i] = pack_bf16(__ldg(W + n * K + k0), __ldg(W + n * K + k0 + 1));
  }
}
// data-gradient operand B[k = o][n = i] = W[o][i], N input columns; PERM: lane tq of the OUTPUT fragment owns columns
// tq*8 .. tq*8+7 of a 32-column block (layer 0 -> d_x)
template <int N, bool PERM>
__device__ __forceinline__ void pack_bwd(uint32_t* dst, const float* __restrict__ W) {
  constexpr int NT = N / 8;
#pragma unroll
  for (int it = 0; it < NT; ++it) {
    const int i = it * 256 + threadIdx.x;
    const int r = i & 1, lane = (i >> 1) & 31, t = i >> 6;
    const int nt = t % NT, ks = t / NT;
    const int g = lane >> 2, tq = lane & 3;
    const int k0 = ks * 16 + tq * 2 + 8 * r;
    const int n = PERM ? 32 * (nt >> 2) + (g >> 1) * 8 + (nt & 3) * 2 + (g & 1) : nt * 8 + g;
    dst[i] = pack_bf16(__ldg(W + k0 * N + n), __ldg(W + (k0 + 1) * N + n));
  }
}

// Output head W_out [out_ch <= 8][64], kept at fp32 grade on the tensor pipe as bf16 hi + lo (w = hi + lo to 2^-17):
// forward operand B[k = column][n = o]: entry (ks * 32 + lane) * 4 + {0, 1: hi b0 b1; 2, 3: lo b0 b1}
__device__ __forceinline__ void split_pair(float w0, float w1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(w0, w1);
  const float2 h = unpack_bf16(hi);
  lo = pack_bf16(w0 - h.x, w1 - h.y);
}
__device__ __forceinline__ void pack_head_fwd(uint32_t* dst, const float* __restrict__ Wo, int out_ch) {
  for (int i = threadIdx.x; i < 4 * 32 * 2; i += 256) {
    const int r = i & 1, lane = (i >> 1) & 31, ks = i >> 6;
    const int g = lane >> 2, tq = lane & 3;
    const int k0 = ks * 16 + tq * 2 + 8 * r;
    const float w0 = g < out_ch ? __ldg(Wo + g * 64 + k0) : 0.0f, w1 = g < out_ch ? __ldg(Wo + g * 64 + k0 + 1) : 0.0f;
    uint32_t hi, lo;
    split_pair(w0, w1, hi, lo);
    dst[(ks * 32 + lane) * 4 + r] = hi;
    dst[(ks * 32 + lane) * 4 + 2 + r] = lo;
  }
}
// data-gradient operand B[k = o][n = column] (k >= 8 is zero: only b0): entry (nt * 32 + lane) * 2 + {0: hi, 1: lo}
__device__ __forceinline__ void pack_head_bwd(uint32_t* dst, const float* __restrict__ Wo, int out_ch) {
  for (int i = threadIdx.x; i < 8 * 32; i += 256) {
    const int lane = i & 31, nt = i >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int col = nt * 8 + g, o = tq * 2;
    const float w0 = o < out_ch ? __ldg(Wo + o * 64 + col) : 0.0f, w1 = o + 1 < out_ch ? __ldg(Wo + (o + 1) * 64 + col) : 0.0f;
    uint32_t hi, lo;
    split_pair(w0, w1, hi, lo);
    dst[i * 2] = hi;
    dst[i * 2 + 1] = lo;
  }
}

// C[16 x 8*NT] (+)= A[16 x 16*KS] * B, B fragments from shared memory
template <int KS, int NT>
__device__ __forceinline__ void mma_layer(float (&C)[8][4], const uint32_t (&A)[4][4], const uint32_t* __restrict__ wfrag,
                                          int lane) {
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const uint2 b = *reinterpret_cast<const uint2*>(wfrag + ((ks * NT + nt) * 32 + lane) * 2);
      mma16816(C[nt], A[ks], b.x, b.y);
    }
  }
}

__device__ __forceinline__ void init_bias(float (&C)[8][4], const float* __restrict__ bias, int tq) {
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float2 b = *reinterpret_cast<const float2*>(bias + nt * 8 + tq * 2);
    C[nt][0] = b.x; C[nt][1] = b.y; C[nt][2] = b.x; C[nt][3] = b.y;
  }
}

// accumulator fragment -> bf16 A fragment of the next layer, ReLU on the packed pair (round-to-bf16 is monotone and
// keeps 0, so max(round(x), 0) == round(max(x, 0)))
__device__ __forceinline__ uint32_t relu_pack2(float lo, float hi) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
  const __nv_bfloat162 v = __hmax2(__floats2bfloat162_rn(lo, hi), z);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void relu_pack(const float (&C)[8][4], uint32_t (&A)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    A[j][0] = relu_pack2(C[2 * j][0], C[2 * j][1]);
    A[j][1] = relu_pack2(C[2 * j][2], C[2 * j][3]);
    A[j][2] = relu_pack2(C[2 * j + 1][0], C[2 * j + 1][1]);
    A[j][3] = relu_pack2(C[2 * j + 1][2], C[2 * j + 1][3]);
  }
}

// data-gradient fragment: C * [h > 0] (h: bf16 A-fragment layout, post-ReLU) -> bf16 A fragment; the mask is taken on
// the packed pair (0xffff per half where h > 0)
__device__ __forceinline__ uint32_t mask_pack2(float lo, float hi, uint32_t h) {
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
  return pack_bf16(lo, hi) & __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&h), z);
}
__device__ __forceinline__ void mask_pack(const float (&C)[8][4], const uint32_t (&H)[4][4], uint32_t (&Y)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    Y[j][0] = mask_pack2(C[2 * j][0], C[2 * j][1], H[j][0]);
    Y[j][1] = mask_pack2(C[2 * j][2], C[2 * j][3], H[j][1]);
    Y[j][2] = mask_pack2(C[2 * j + 1][0], C[2 * j + 1][1], H[j][2]);
    Y[j][3] = mask_pack2(C[2 * j + 1][2], C[2 * j + 1][3], H[j][3]);
  }
}

// a warp's 16 rows of a [128][kTileStride] tile <-> its A fragment: A[j][0..3] are the four 8 x 8 blocks (rows 0-7 | 8-15)
// x (columns 16j.. | 16j+8..), exactly the register layout of stmatrix / ldmatrix .x4; `frag_addr` is this lane's row
// address for j = 0 (lane -> block lane/8, row lane%8), 144-byte row stride: conflict-free
__device__ __forceinline__ uint32_t frag_addr(const bf16* tile, int wrow, int lane) {
  return smem_addr(tile + (wrow + (lane & 7) + ((lane >> 3) & 1) * 8) * kTileStride + ((lane >> 4) & 1) * 8);
}
__device__ __forceinline__ void store_frag(uint32_t addr, const uint32_t (&A)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};" ::"r"(addr + j * 32), "r"(A[j][0]),
                 "r"(A[j][1]), "r"(A[j][2]), "r"(A[j][3]) : "memory");
}
__device__ __forceinline__ void load_frag(uint32_t addr, uint32_t (&A)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(A[j][0]), "=r"(A[j][1]), "=r"(A[j][2]), "=r"(A[j][3]) : "r"(addr + j * 32) : "memory");
}

template <int D, int KIN>
struct FwdSmem {
  static constexpr int wf0 = 0;                                  // u32 units
  static constexpr int wfh = wf0 + (KIN / 16) * 512;
  static constexpr int bias = wfh + (D - 1) * 2048;              // fp32 [D][64]
  static constexpr int wo = bias + D * 64;                       // head fragments, hi + lo (pack_head_fwd)
  static constexpr int bo = wo + 512;                            // fp32 [8]
  static constexpr int words = bo + 8;
};

template <int D>
__device__ __forceinline__ void load_small_params(const TinyArgs& a, float* bias, uint32_t* wo, float* bo) {
#pragma unroll
  for (int l = 0; l < D; ++l)
    if (threadIdx.x < 64) bias[l * 64 + threadIdx.x] = __ldg(a.params + a.b_off[l] + threadIdx.x);
  pack_head_fwd(wo, a.params + a.wo_off, a.out_ch);
  if (threadIdx.x < 8) bo[threadIdx.x] = (int)threadIdx.x < a.out_ch ? __ldg(a.params + a.bo_off + threadIdx.x) : 0.0f;
}

// ================================================================================================ forward
template <int D, int KIN, bool SAVE>
__global__ void __launch_bounds__(256, 2)
tiny_fwd_kernel(const TinyArgs a, const float* __restrict__ x, float* __restrict__ out, bf16* __restrict__ x0s, int ldx0) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  using S = FwdSmem<D, KIN>;
  uint32_t* wf0 = smem_u32 + S::wf0;
  uint32_t* wfh = smem_u32 + S::wfh;
  float* bias = reinterpret_cast<float*>(smem_u32 + S::bias);
  uint32_t* wo = smem_u32 + S::wo;
  float* bo = reinterpret_cast<float*>(smem_u32 + S::bo);
  pack_fwd<KIN, true>(wf0, a.params + a.w_off[0]);
#pragma unroll
  for (int l = 1; l < D; ++l) pack_fwd<64, false>(wfh + (l - 1) * 2048, a.params + a.w_off[l]);
  load_small_params<D>(a, bias, wo, bo);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
  const int64_t P = a.P, n_tiles = (P + 127) >> 7;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * 128 + warp * 16;
    if (r0 >= P) continue;
    const int64_t ra = r0 + g, rb = r0 + g + 8;
    const bool va = ra < P, vb = rb < P;
    {  // the CTA's next tile: pull this lane's input bytes towards L2 while this tile computes
      const int64_t na = ra + (int64_t)gridDim.x * 128, nb = na + 8;
#pragma unroll
      for (int b = 0; b < KIN / 32; ++b) {
        if (na < P) prefetch_l2(x + na * KIN + b * 32 + tq * 8);
        if (nb < P) prefetch_l2(x + nb * KIN + b * 32 + tq * 8);
      }
    }
    uint32_t A[4][4];
#pragma unroll
    for (int b = 0; b < KIN / 32; ++b) {
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* pa = reinterpret_cast<const float4*>(x + ra * KIN + b * 32 + tq * 8);
      const float4* pb = reinterpret_cast<const float4*>(x + rb * KIN + b * 32 + tq * 8);
      const float4 a0 = va ? __ldg(pa) : z4, a1 = va ? __ldg(pa + 1) : z4;
      const float4 b0 = vb ? __ldg(pb) : z4, b1 = vb ? __ldg(pb + 1) : z4;
      A[2 * b][0] = pack_bf16(a0.x, a0.y); A[2 * b][2] = pack_bf16(a0.z, a0.w);
      A[2 * b + 1][0] = pack_bf16(a1.x, a1.y); A[2 * b + 1][2] = pack_bf16(a1.z, a1.w);
      A[2 * b][1] = pack_bf16(b0.x, b0.y); A[2 * b][3] = pack_bf16(b0.z, b0.w);
      A[2 * b + 1][1] = pack_bf16(b1.x, b1.y); A[2 * b + 1][3] = pack_bf16(b1.z, b1.w);
      if (SAVE) {
        if (va) *reinterpret_cast<uint4*>(x0s + ra * ldx0 + b * 32 + tq * 8) =
            make_uint4(A[2 * b][0], A[2 * b][2], A[2 * b + 1][0], A[2 * b + 1][2]);
        if (vb) *reinterpret_cast<uint4*>(x0s + rb * ldx0 + b * 32 + tq * 8) =
            make_uint4(A[2 * b][1], A[2 * b][3], A[2 * b + 1][1], A[2 * b + 1][3]);
      }
    }
    float C[8][4];
    init_bias(C, bias, tq);
    mma_layer<KIN / 16, 8>(C, A, wf0, lane);
    relu_pack(C, A);
#pragma unroll
    for (int l = 1; l < D; ++l) {
      init_bias(C, bias + l * 64, tq);
      mma_layer<4, 8>(C, A, wfh + (l - 1) * 2048, lane);
      relu_pack(C, A);
    }
    // output head from the bf16-rounded h_{D-1}, fp32-grade weights (hi + lo): C rows g / g+8, columns o = tq*2, tq*2+1
    float O[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint4 w = *reinterpret_cast<const uint4*>(wo + (ks * 32 + lane) * 4);
      mma16816(O, A[ks], w.x, w.y);
      mma16816(O, A[ks], w.z, w.w);
    }
    const int o0 = tq * 2;
    if (o0 < a.out_ch) {
      if (va) out[ra * a.out_ch + o0] = O[0] + bo[o0];
      if (vb) out[rb * a.out_ch + o0] = O[2] + bo[o0];
    }
    if (o0 + 1 < a.out_ch) {
      if (va) out[ra * a.out_ch + o0 + 1] = O[1] + bo[o0 + 1];
      if (vb) out[rb * a.out_ch + o0 + 1] = O[3] + bo[o0 + 1];
    }
  }
}

// ================================================================================================ backward
template <int D, int KIN>
struct BwdSmem {
  static constexpr int x0_stride = KIN + 8;                       // bf16 elements
  static constexpr int wf0 = 0;                                   // u32 units from here on
  static constexpr int wfh = wf0 + (KIN / 16) * 512;
  static constexpr int wb0 = wfh + (D - 1) * 2048;
  static constexpr int wbh = wb0 + 4 * (KIN / 8) * 64;
  static constexpr int bias = wbh + (D - 1) * 2048;
  static constexpr int wo = bias + D * 64;                        // head fragments, hi + lo (pack_head_bwd)
  static constexpr int t_do = wo + 512;                           // d_out as bf16 hi | lo, [2][128][8]
  static constexpr int t_x0 = t_do + 128 * 8;                     // bf16 [128][KIN + 8]
  static constexpr int t_h = t_x0 + 128 * x0_stride / 2;          // bf16 [D][128][72]
  static constexpr int t_dy = t_h + D * 128 * kTileStride / 2;    // bf16 [D][128][72]
  static constexpr int words = t_dy + D * 128 * kTileStride / 2;
  static constexpr int ctas_per_sm = (2 * (words * 4 + 1024) <= 228 * 1024) ? 2 : 1;
};

template <int D, int KIN>
__global__ void __launch_bounds__(256, (BwdSmem<D, KIN>::ctas_per_sm))
tiny_bwd_kernel(const TinyArgs a, const bf16* __restrict__ x0, int ldx0, const float* __restrict__ d_out,
                float* __restrict__ d_params, float* __restrict__ d_in) {
  extern __shared__ __align__(16) uint32_t smem_u32[];
  using S = BwdSmem<D, KIN>;
  uint32_t* wf0 = smem_u32 + S::wf0;
  uint32_t* wfh = smem_u32 + S::wfh;
  uint32_t* wb0 = smem_u32 + S::wb0;
  uint32_t* wbh = smem_u32 + S::wbh;
  float* bias = reinterpret_cast<float*>(smem_u32 + S::bias);
  uint32_t* wo = smem_u32 + S::wo;
  bf16* t_do = reinterpret_cast<bf16*>(smem_u32 + S::t_do);  // [0]: hi, [1024 elements on]: lo
  bf16* t_x0 = reinterpret_cast<bf16*>(smem_u32 + S::t_x0);
  bf16* t_h = reinterpret_cast<bf16*>(smem_u32 + S::t_h);
  bf16* t_dy = reinterpret_cast<bf16*>(smem_u32 + S::t_dy);
  constexpr int kTile = 128 * kTileStride;  // bf16 elements per [128][64] tile
  {
    pack_fwd<KIN, true>(wf0, a.params + a.w_off[0]);
    pack_bwd<KIN, true>(wb0, a.params + a.w_off[0]);
#pragma unroll
    for (int l = 1; l < D; ++l) {
      pack_fwd<64, false>(wfh + (l - 1) * 2048, a.params + a.w_off[l]);
      pack_bwd<64, false>(wbh + (l - 1) * 2048, a.params + a.w_off[l]);
    }
#pragma unroll
    for (int l = 0; l < D; ++l)
      if (threadIdx.x < 64) bias[l * 64 + threadIdx.x] = __ldg(a.params + a.b_off[l] + threadIdx.x);
    pack_head_bwd(wo, a.params + a.wo_off, a.out_ch);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
  const int wrow = warp * 16;
  const int mt = warp & 3, hh = warp >> 2;  // weight-gradient block of this warp: dW rows mt*16.., column half hh
  constexpr int NTW0 = KIN / 16;            // n-tiles per warp, layer 0 (input columns) / hidden layers: 4
  float accW0[NTW0][4], accWh[D > 1 ? D - 1 : 1][4][4], accB[D][4], accO[4], dbo[2] = {0.0f, 0.0f};
#pragma unroll
  for (int q = 0; q < NTW0; ++q)
#pragma unroll
    for (int e = 0; e < 4; ++e) accW0[q][e] = 0.0f;
#pragma unroll
  for (int l = 0; l < (D > 1 ? D - 1 : 1); ++l)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) accWh[l][q][e] = 0.0f;
#pragma unroll
  for (int l = 0; l < D; ++l)
#pragma unroll
    for (int e = 0; e < 4; ++e) accB[l][e] = 0.0f;
#pragma unroll
  for (int e = 0; e < 4; ++e) accO[e] = 0.0f;

  // phase-2 operand addresses (32-bit shared-memory byte addresses of this lane's ldmatrix rows at k-step 0):
  //   A = dY_l^T: matrices (points 0-7 | 8-15) x (o 0-7 | 8-15) of the warp's 16 rows mt*16..;  B = X_l: two n-tiles per x4
  const int a_row = (lane & 7) + ((lane >> 4) & 1) * 8, a_col = mt * 16 + ((lane >> 3) & 1) * 8;
  const int b_row = (lane & 7) + ((lane >> 3) & 1) * 8, b_col = ((lane >> 4) & 1) * 8;
  const uint32_t sm_a = smem_addr(t_dy + a_row * kTileStride + a_col);
  const uint32_t sm_b0 = smem_addr(t_x0 + b_row * S::x0_stride + hh * NTW0 * 8 + b_col);
  const uint32_t sm_bh = smem_addr(t_h + b_row * kTileStride + hh * 32 + b_col);
  const uint32_t sm_hl = smem_addr(t_h + (D - 1) * kTile + a_row * kTileStride + a_col);  // A = h_{D-1}^T, rows mt*16..
  const uint32_t sm_do = smem_addr(t_do + hh * 1024 + b_row * 8);  // B = d_out (hi for warps 0-3, lo for warps 4-7)
  const uint32_t sm_fh = frag_addr(t_h, wrow, lane), sm_fy = frag_addr(t_dy, wrow, lane);  // phase-1 fragment rows

  const int64_t P = a.P, n_tiles = (P + 127) >> 7;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---------------------------------------------------------------- phase 1: a warp's 16 points, in registers
    const int64_t ra = tile * 128 + wrow + g, rb = ra + 8;
    const bool va = ra < P, vb = rb < P;
    {  // the CTA's next tile: saved input and d_out rows towards L2
      const int64_t na = ra + (int64_t)gridDim.x * 128, nb = na + 8;
      if (na < P) { prefetch_l2(x0 + na * ldx0 + tq * 8); if (tq == 0) prefetch_l2(d_out + na * a.out_ch); }
      if (nb < P) { prefetch_l2(x0 + nb * ldx0 + tq * 8); if (tq == 0) prefetch_l2(d_out + nb * a.out_ch); }
      if (KIN == 64) {
        if (na < P) prefetch_l2(x0 + na * ldx0 + 32 + tq * 8);
        if (nb < P) prefetch_l2(x0 + nb * ldx0 + 32 + tq * 8);
      }
    }
    uint32_t A[4][4];
#pragma unroll
    for (int b = 0; b < KIN / 32; ++b) {
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      const uint4 xa = va ? __ldg(reinterpret_cast<const uint4*>(x0 + ra * ldx0 + b * 32 + tq * 8)) : z4;
      const uint4 xb = vb ? __ldg(reinterpret_cast<const uint4*>(x0 + rb * ldx0 + b * 32 + tq * 8)) : z4;
      A[2 * b][0] = xa.x; A[2 * b][2] = xa.y; A[2 * b + 1][0] = xa.z; A[2 * b + 1][2] = xa.w;
      A[2 * b][1] = xb.x; A[2 * b][3] = xb.y; A[2 * b + 1][1] = xb.z; A[2 * b + 1][3] = xb.w;
      *reinterpret_cast<uint4*>(t_x0 + (wrow + g) * S::x0_stride + b * 32 + tq * 8) = xa;
      *reinterpret_cast<uint4*>(t_x0 + (wrow + g + 8) * S::x0_stride + b * 32 + tq * 8) = xb;
    }
    // d_out of rows g / g+8, channels o = tq*2, tq*2+1: the A fragment of the head's data gradient (K = o, padded to 16)
    const int o0 = tq * 2;
    const float da0 = (va && o0 < a.out_ch) ? __ldg(d_out + ra * a.out_ch + o0) : 0.0f;
    const float da1 = (va && o0 + 1 < a.out_ch) ? __ldg(d_out + ra * a.out_ch + o0 + 1) : 0.0f;
    const float db0 = (vb && o0 < a.out_ch) ? __ldg(d_out + rb * a.out_ch + o0) : 0.0f;
    const float db1 = (vb && o0 + 1 < a.out_ch) ? __ldg(d_out + rb * a.out_ch + o0 + 1) : 0.0f;
    dbo[0] += da0 + db0;  // db_out[o] = sum_p d_out[p][o]: this lane's two rows, channels tq*2, tq*2+1
    dbo[1] += da1 + db1;
    uint32_t dhi[4], dlo[4];  // bf16 hi + lo of d_out: A fragment of the head's data gradient, and (through the
    split_pair(da0, da1, dhi[0], dlo[0]);  // shared-memory tiles) the transposed operand of its weight gradient
    split_pair(db0, db1, dhi[1], dlo[1]);
    dhi[2] = dhi[3] = dlo[2] = dlo[3] = 0u;
    *reinterpret_cast<uint32_t*>(t_do + (wrow + g) * 8 + o0) = dhi[0];
    *reinterpret_cast<uint32_t*>(t_do + (wrow + g + 8) * 8 + o0) = dhi[1];
    *reinterpret_cast<uint32_t*>(t_do + 1024 + (wrow + g) * 8 + o0) = dlo[0];
    *reinterpret_cast<uint32_t*>(t_do + 1024 + (wrow + g + 8) * 8 + o0) = dlo[1];
    float C[8][4];
    // recompute h_0 .. h_{D-1} (kept in the shared-memory tiles: ReLU masks and weight-gradient operands)
    init_bias(C, bias, tq);
    mma_layer<KIN / 16, 8>(C, A, wf0, lane);
    relu_pack(C, A);
    store_frag(sm_fh, A);
#pragma unroll
    for (int l = 1; l < D; ++l) {
      init_bias(C, bias + l * 64, tq);
      mma_layer<4, 8>(C, A, wfh + (l - 1) * 2048, lane);
      relu_pack(C, A);
      store_frag(sm_fh + l * kTile * 2, A);
    }
    // dY_{D-1} = (d_out W_out) * [h_{D-1} > 0]: d_out and W_out as bf16 hi + lo, three MMAs per n-tile (the lo * lo term
    // is below fp32 rounding); A still holds h_{D-1}
    {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint2 w = *reinterpret_cast<const uint2*>(wo + (nt * 32 + lane) * 2);
        C[nt][0] = C[nt][1] = C[nt][2] = C[nt][3] = 0.0f;
        mma16816(C[nt], dlo, w.x, 0u);
        mma16816(C[nt], dhi, w.y, 0u);
        mma16816(C[nt], dhi, w.x, 0u);
      }
    }
    uint32_t Y[4][4];
    mask_pack(C, A, Y);
    store_frag(sm_fy + (D - 1) * kTile * 2, Y);
#pragma unroll
    for (int l = D - 1; l >= 1; --l) {  // dY_{l-1} = (dY_l W_l) * [h_{l-1} > 0]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) C[nt][0] = C[nt][1] = C[nt][2] = C[nt][3] = 0.0f;
      mma_layer<4, 8>(C, Y, wbh + (l - 1) * 2048, lane);
      __syncwarp();  // stmatrix rows were written by other lanes of this warp
      load_frag(sm_fh + (l - 1) * kTile * 2, A);
      mask_pack(C, A, Y);
      store_frag(sm_fy + (l - 1) * kTile * 2, Y);
    }
    if (d_in != nullptr) {  // d_x = dY_0 W_0: a lane's 8 output columns of a 32-column block are contiguous
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) C[nt][0] = C[nt][1] = C[nt][2] = C[nt][3] = 0.0f;
      mma_layer<4, KIN / 8>(C, Y, wb0, lane);
#pragma unroll
      for (int b = 0; b < KIN / 32; ++b) {
        float4* pa = reinterpret_cast<float4*>(d_in + ra * KIN + b * 32 + tq * 8);
        float4* pb = reinterpret_cast<float4*>(d_in + rb * KIN + b * 32 + tq * 8);
        if (va) {
          pa[0] = make_float4(C[4 * b][0], C[4 * b][1], C[4 * b + 1][0], C[4 * b + 1][1]);
          pa[1] = make_float4(C[4 * b + 2][0], C[4 * b + 2][1], C[4 * b + 3][0], C[4 * b + 3][1]);
        }
        if (vb) {
          pb[0] = make_float4(C[4 * b][2], C[4 * b][3], C[4 * b + 1][2], C[4 * b + 1][3]);
          pb[1] = make_float4(C[4 * b + 2][2], C[4 * b + 2][3], C[4 * b + 3][2], C[4 * b + 3][3]);
        }
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 2: weight gradients of the 128 points
    // dW_l[o][i] += sum_p dY_l[p][o] X_l[p][i]: A = dY_l^T and B = X_l both come out of the [point][column] tiles
    // through ldmatrix.trans; this warp owns rows mt*16.. and a half of the columns.
    // One fully unrolled loop over the 8 k-steps (16 points each) carries every product -- ten independent accumulator
    // chains per warp, so no MMA waits for its predecessor -- with all shared-memory addresses as base + immediate.
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      {  // output head, transposed: dW_out^T[c][o] += sum_p h_{D-1}[p][c] d_out[p][o] with A = h^T (rows mt*16..) and
         // B = d_out as bf16 hi (warps 0-3) or lo (warps 4-7): the two halves add up to an fp32-grade product
        uint32_t ah[4], bd[2];
        ldmatrix_x4_trans(ah, sm_hl + ks * 16 * kTileStride * 2);
        ldmatrix_x2_trans(bd, sm_do + ks * 16 * 16);
        mma16816(accO, ah, bd[0], bd[1]);
      }
#pragma unroll
      for (int l = 0; l < D; ++l) {
        uint32_t af[4];
        ldmatrix_x4_trans(af, sm_a + (l * kTile + ks * 16 * kTileStride) * 2);
        if (l == 0) {
#pragma unroll
          for (int q = 0; q < NTW0 / 2; ++q) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, sm_b0 + (ks * 16 * S::x0_stride + 2 * q * 8) * 2);
            mma16816(accW0[2 * q], af, bf[0], bf[1]);
            mma16816(accW0[2 * q + 1], af, bf[2], bf[3]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t bf[4];
            ldmatrix_x4_trans(bf, sm_bh + ((l - 1) * kTile + ks * 16 * kTileStride + 2 * q * 8) * 2);
            mma16816(accWh[l > 0 ? l - 1 : 0][2 * q], af, bf[0], bf[1]);
            mma16816(accWh[l > 0 ? l - 1 : 0][2 * q + 1], af, bf[2], bf[3]);
          }
        }
        mma16816(accB[l], af, kOnes, kOnes);  // db_l[o] = sum_p dY_l[p][o] (both column halves compute it: no branch)
      }
    }
    __syncthreads();
  }

  // ------------------------------------------------------------------ flush (once per CTA)
  {
    float* dW0 = d_params + a.w_off[0];
#pragma unroll
    for (int q = 0; q < NTW0; ++q) {
      const int col = (hh * NTW0 + q) * 8 + tq * 2;
      atomicAdd(reinterpret_cast<float2*>(dW0 + (mt * 16 + g) * KIN + col), make_float2(accW0[q][0], accW0[q][1]));
      atomicAdd(reinterpret_cast<float2*>(dW0 + (mt * 16 + g + 8) * KIN + col), make_float2(accW0[q][2], accW0[q][3]));
    }
#pragma unroll
    for (int l = 1; l < D; ++l) {
      float* dW = d_params + a.w_off[l];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = (hh * 4 + q) * 8 + tq * 2;
        atomicAdd(reinterpret_cast<float2*>(dW + (mt * 16 + g) * 64 + col), make_float2(accWh[l - 1][q][0], accWh[l - 1][q][1]));
        atomicAdd(reinterpret_cast<float2*>(dW + (mt * 16 + g + 8) * 64 + col),
                  make_float2(accWh[l - 1][q][2], accWh[l - 1][q][3]));
      }
    }
    if (hh == 0 && tq == 0) {
#pragma unroll
      for (int l = 0; l < D; ++l) {
        atomicAdd(d_params + a.b_off[l] + mt * 16 + g, accB[l][0]);
        atomicAdd(d_params + a.b_off[l] + mt * 16 + g + 8, accB[l][2]);
      }
    }
    {  // head: accO = dW_out^T[c = mt*16 + g (+8)][o = tq*2 (+1)], hi and lo halves both add in
      const int o0 = tq * 2, c0 = mt * 16 + g;
      float* dWo = d_params + a.wo_off;
      if (o0 < a.out_ch) { atomicAdd(dWo + o0 * 64 + c0, accO[0]); atomicAdd(dWo + o0 * 64 + c0 + 8, accO[2]); }
      if (o0 + 1 < a.out_ch) { atomicAdd(dWo + (o0 + 1) * 64 + c0, accO[1]); atomicAdd(dWo + (o0 + 1) * 64 + c0 + 8, accO[3]); }
#pragma unroll
      for (int e = 0; e < 2; ++e) {  // db_out: sum this lane's partial over the 8 row groups of the warp
        float v = dbo[e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0 && o0 + e < a.out_ch) atomicAdd(d_params + a.bo_off + o0 + e, v);
      }
    }
  }
}

TinyArgs make_args(const TinyMlpDesc& d) {
  TinyArgs a;
  a.params = d.params;
  for (int l = 0; l < kTinyMaxLayers; ++l) { a.w_off[l] = d.w_off[l < d.D ? l : 0]; a.b_off[l] = d.b_off[l < d.D ? l : 0]; }
  a.wo_off = d.wo_off; a.bo_off = d.bo_off; a.out_ch = d.out_ch; a.P = d.P;
  return a;
}

int check_desc(const TinyMlpDesc& d) {
  if (d.D < 1 || d.D > kTinyMaxLayers || (d.in_pos != 32 && d.in_pos != 64) || (d.D == 4 && d.in_pos == 64) || d.out_ch < 1 ||
      d.out_ch > 8 || d.P < 0) {
    set_error("tiny MLP: 1 <= D <= %d (D = 4: in_pos 32 only), in_pos in {32, 64}, 1 <= out_ch <= 8", kTinyMaxLayers);
    return NMX_E_BADARG;
  }
  // the weight gradients are flushed as float2 atomics: even weight offsets (true for every packed layout of such a net)
  for (int l = 0; l < d.D; ++l)
    if (d.w_off[l] & 1) { set_error("tiny MLP: odd weight offset"); return NMX_E_BADARG; }
  return 0;
}

template <int D, int KIN, bool SAVE>
int run_fwd(const TinyArgs& a, const float* x, float* out, bf16* x0s, int ldx0, int blocks, cudaStream_t s) {
  constexpr int smem = FwdSmem<D, KIN>::words * 4;
  static bool attr[64] = {};
  if (smem > 48 * 1024 && once_per_device(attr))
    NMX_CUDA(cudaFuncSetAttribute(tiny_fwd_kernel<D, KIN, SAVE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tiny_fwd_kernel<D, KIN, SAVE><<<blocks, 256, smem, s>>>(a, x, out, x0s, ldx0);
  NMX_LAUNCH_CHECK();
  return 0;
}

template <int D, int KIN>
int run_bwd(const TinyArgs& a, const bf16* x0, int ldx0, const float* d_out, float* d_params, float* d_in, int blocks,
            cudaStream_t s) {
  constexpr int smem = BwdSmem<D, KIN>::words * 4;
  static_assert(smem <= 227 * 1024, "tiny backward: shared memory");
  static bool attr[64] = {};
  if (once_per_device(attr))
    NMX_CUDA(cudaFuncSetAttribute(tiny_bwd_kernel<D, KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tiny_bwd_kernel<D, KIN><<<blocks, 256, smem, s>>>(a, x0, ldx0, d_out, d_params, d_in);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace

namespace nmx {

int launch_tiny_fwd(const TinyMlpDesc& d, const float* x, float* out, bf16* x0_save, int ldx0, cudaStream_t s) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (d.P == 0) return 0;
  if (x0_save != nullptr && ((ldx0 & 7) || (reinterpret_cast<uintptr_t>(x0_save) & 15))) {
    set_error("tiny MLP: saved-input rows must be 16-byte aligned");
    return NMX_E_BADARG;
  }
  if (reinterpret_cast<uintptr_t>(x) & 15) { set_error("tiny MLP: input must be 16-byte aligned"); return NMX_E_BADARG; }
  const TinyArgs a = make_args(d);
  const int64_t n_tiles = (d.P + 127) / 128;
  const int blocks = (int)(n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs);
#define NMX_TINY_FWD(DD, KK)                                                                          \
  if (d.D == DD && d.in_pos == KK)                                                                    \
    return x0_save ? run_fwd<DD, KK, true>(a, x, out, x0_save, ldx0, blocks, s)                       \
                   : run_fwd<DD, KK, false>(a, x, out, nullptr, 0, blocks, s)
  NMX_TINY_FWD(1, 32); NMX_TINY_FWD(2, 32); NMX_TINY_FWD(3, 32); NMX_TINY_FWD(4, 32);
  NMX_TINY_FWD(1, 64); NMX_TINY_FWD(2, 64); NMX_TINY_FWD(3, 64);
#undef NMX_TINY_FWD
  return NMX_E_UNSUPPORTED;
}

int launch_tiny_bwd(const TinyMlpDesc& d, const bf16* x0, int ldx0, const float* d_out, float* d_params, float* d_input,
                    cudaStream_t s) {
  int rc = check_desc(d);
  if (rc) return rc;
  if (d.P == 0) return 0;
  if ((ldx0 & 7) || (reinterpret_cast<uintptr_t>(x0) & 15) || (reinterpret_cast<uintptr_t>(d_params) & 7) ||
      (d_input && (reinterpret_cast<uintptr_t>(d_input) & 15))) {
    set_error("tiny MLP backward: operands must be 16-byte aligned");
    return NMX_E_BADARG;
  }
  const TinyArgs a = make_args(d);
  const int64_t n_tiles = (d.P + 127) / 128;
#define NMX_TINY_BWD(DD, KK)                                                                  \
  if (d.D == DD && d.in_pos == KK) {                                                          \
    const int cap = BwdSmem<DD, KK>::ctas_per_sm * kNumSMs;                                   \
    return run_bwd<DD, KK>(a, x0, ldx0, d_out, d_params, d_input, (int)(n_tiles < cap ? n_tiles : cap), s); \
  }
  NMX_TINY_BWD(1, 32); NMX_TINY_BWD(2, 32); NMX_TINY_BWD(3, 32); NMX_TINY_BWD(4, 32);
  NMX_TINY_BWD(1, 64); NMX_TINY_BWD(2, 64); NMX_TINY_BWD(3, 64);
#undef NMX_TINY_BWD
  return NMX_E_UNSUPPORTED;
}

}  // namespace nmx
