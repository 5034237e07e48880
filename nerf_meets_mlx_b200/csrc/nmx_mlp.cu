// K3: the NeRF MLP (models/NeRF.py:160-243) on the tcgen05 GEMM building blocks of nmx_gemm.cu.
//
// Forward   : encode inputs (PE generated on the fly from rays/z, never materialised in fp32) -> bf16 operand tile
//             X0 = [PE(pos) | pad | PE(dir) | pad];  trunk layers h_l = relu(W_l [x|h] + b_l) as bf16 GEMMs with fp32
//             accumulation in TMEM; skip concat [input_pos, h] is a K-concatenation of two TMA sources (no copy);
//             view-dir head: feature (no act), dir layer relu, rgb/alpha heads (tiny N) on CUDA cores -> raw fp32.
// Backward  : heads -> dir layer -> feature -> trunk, each layer = one dgrad GEMM (ReLU mask of the saved
//             activation and the alpha rank-1 term fused in the epilogue) + one wgrad GEMM (contraction over points,
//             MN-major operands straight from the saved row-major activations) + a column-sum for the bias.
// Dead work skipped exactly as autograd would: no dgrad into pure-encoding inputs.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "nmx_common.cuh"
#include "nmx_gemm.cuh"
#include "nmx_tiny.cuh"
#include "nmx_chain.cuh"
#include "nmx_sm100.cuh"


using namespace nmx;
typedef __nv_bfloat16 bf16;

namespace {
inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }
constexpr int64_t kInferChunk = 65536;  // points per inference chunk (activations stay L2-resident)
}  // namespace

struct LinearRef {
  int out, in;
  int64_t w_off, b_off;  // offsets (floats) into the packed fp32 parameter buffer
};

struct nmx_mlp_plan {
  nmx_mlp_config cfg;
  int64_t max_points;
  int D, W, in_pos, in_dir, pos_pad, dir_pad, x0_cols;
  std::vector<LinearRef> trunk;  // D layers
  LinearRef feat, alpha, dir, rgb, outl;
  int64_t n_params;
  // packed bf16 weights (offsets in bytes from workspace base)
  std::vector<int64_t> wf_off, wt_off;  // forward [W, K_l] and transposed-h-part [W, W] per trunk layer
  std::vector<int> wf_k;                // padded K of each trunk layer
  int64_t wf_feat, wt_feat, wf_dir, wt_dir;
  int wf_dir_k;
  int64_t wt_in;  // [pos_pad, 2W]: transposed position-input columns of layer 0 (k < W) and of the skip layer (k >= W)
  int64_t weights_bytes;
  // activation region (offsets relative to its start; depend on capacity)
  int64_t act_points_train, act_points_infer;
  int bits_word_major;  // layout of the sign-bit tiles written by the last saving forward (1 = pair kernel, word-major)
  int last_fwd_tiny;    // the last saving forward ran the fused width-64 kernel (nmx_tiny.cu): only the bf16 input is saved
};

namespace {

bool chain_eligible(const nmx_mlp_plan* p);
// points per inference chunk: the fused chain only needs the encoded-input tile per chunk -> 16 full waves of tiles
int64_t infer_cap(const nmx_mlp_plan* p) { return chain_eligible(p) ? (int64_t)kNumSMs * 128 * 16 : kInferChunk; }

struct ActLayout {
  int64_t x0, h0, feat, hd, g0, ghd, bits, gfold, dsig, total;
  int64_t h_stride;  // bytes between consecutive saved trunk activations (0 = ping-pong of 2 buffers)
  int64_t g_stride;  // bytes between gradient buffers: 2 ping-pong buffers, or D+1 saved dY slots (fused chain)
};

bool chain_bwd_eligible(const nmx_mlp_plan* p);

ActLayout act_layout(const nmx_mlp_plan* p, int64_t cap, bool training) {
  ActLayout a;
  int64_t off = 0;
  a.x0 = off; off += align256(cap * p->x0_cols * 2);
  const bool x0_only = !training && chain_eligible(p);  // inference through the fused chain keeps activations on chip
  int64_t hbytes = x0_only ? 0 : align256(cap * p->W * 2);
  a.h0 = off;
  a.h_stride = hbytes;
  off += hbytes * (training ? p->D : 2);
  a.feat = off; off += hbytes;
  a.hd = off; off += x0_only ? 0 : align256(cap * (p->W / 2) * 2);
  a.g0 = a.ghd = a.bits = a.gfold = a.dsig = 0;
  a.g_stride = hbytes;
  if (training) {
    // fused backward chain: every layer's dY is kept for the wgrad kernels (slots 0..D-1 = dY_l, slot D = d_feature)
    // every layer's dY in its own slot where ALL weight gradients run as one batched launch: the fused backward chain,
    // and the layer-by-layer path of nets without a view-dir head
    a.g0 = off; off += hbytes * ((chain_bwd_eligible(p) || !p->cfg.use_viewdirs) ? p->D + 1 : 2);
    a.ghd = off; off += align256(cap * (p->W / 2) * 2);
    // ReLU sign bits (32 B per point and slot): slots 0..D-1 = h_l, slot D = hd
    a.bits = off; if (chain_bwd_eligible(p)) off += align256(cap * 32) * (p->D + 1);
    // fp32 scratch G = d_hd^T h_{D-1} [W/2, W] of the feature / dir-layer weight-gradient folding
    a.gfold = off; if (chain_bwd_eligible(p)) off += align256((int64_t)(p->W / 2) * p->W * 4);
  }
  a.total = off;
  return a;
}

// ------------------------------------------------------------------------------------------------ weight packing
// One launch packs every bf16 operand copy: segment t copies (optionally transposed) a [rows, cols] window of an
// fp32 weight matrix into its padded bf16 layout.  dst[r, dst_col0 + c] = src[r, src_col0 + c], or transposed
// dst[c, r] = src[r, src_col0 + c].
struct PackSeg {
  int64_t src_off;   // floats, into params
  int64_t dst_off;   // bytes, into the workspace
  int src_ld, src_col0, dst_ld, dst_col0, rows, cols, transpose, pad_;
};
constexpr int kMaxPackSegs = 48;
struct PackTable {
  int n;
  PackSeg seg[kMaxPackSegs];
};

__global__ void __launch_bounds__(256)
pack_weights_kernel(const float* __restrict__ params, uint8_t* __restrict__ ws, const PackTable tab) {
  const PackSeg sg = tab.seg[blockIdx.y];
  const float* src = params + sg.src_off;
  bf16* dst = reinterpret_cast<bf16*>(ws + sg.dst_off);
  const int total = sg.rows * sg.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    if (!sg.transpose) {
      int r = i / sg.cols, c = i - r * sg.cols;
      dst[(size_t)r * sg.dst_ld + sg.dst_col0 + c] = __float2bfloat16_rn(src[(size_t)r * sg.src_ld + sg.src_col0 + c]);
    } else {
      int c = i / sg.rows, r = i - c * sg.rows;
      dst[(size_t)c * sg.dst_ld + sg.dst_col0 + r] = __float2bfloat16_rn(src[(size_t)r * sg.src_ld + sg.src_col0 + c]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ input encoding
// enc_kind 1: X0[p, :] = [pos(3), {sin(k^2 pos), cos(k^2 pos)}_k, 0-pad | dir PE (precomputed per ray), 0-pad]
// One thread per (point, unit); unit u in [0, n_freqs_pos] handles the raw input (u=0) or band u-1; the last
// units copy the per-ray view-dir encoding and write the zero padding.
__global__ void __launch_bounds__(256)
encode_rays_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z,
                   const bf16* __restrict__ dir_pe, int64_t b0, bf16* __restrict__ x0, int64_t p0, int64_t npts, int n,
                   int n_freqs_pos, int in_pos, int pos_pad, int dir_pad) {
  const int units = n_freqs_pos + 2;  // input, bands..., tail (padding + dir copy)
  const int x0_cols = pos_pad + dir_pad;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts * units;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t lp = t / units;
    int u = (int)(t - lp * units);
    int64_t p = p0 + lp;
    int64_t b = p / n;
    bf16* row = x0 + lp * x0_cols;
    if (u == units - 1) {
      for (int c = in_pos; c < pos_pad; ++c) row[c] = __float2bfloat16_rn(0.0f);
      if (dir_pad > 0) {
        const uint4* src = reinterpret_cast<const uint4*>(dir_pe + (b - b0) * dir_pad);
        uint4* dst = reinterpret_cast<uint4*>(row + pos_pad);
        for (int c = 0; c < dir_pad / 8; ++c) dst[c] = __ldg(src + c);
      }
      continue;
    }
    const float* r = rays + b * ray_stride;
    float zz = z[p];
    float px = __fadd_rn(r[0], __fmul_rn(zz, r[3]));  // pos = o + z*d (render.py:142), mul then add
    float py = __fadd_rn(r[1], __fmul_rn(zz, r[4]));
    float pz = __fadd_rn(r[2], __fmul_rn(zz, r[5]));
    if (u == 0) {
      row[0] = __float2bfloat16_rn(px);
      row[1] = __float2bfloat16_rn(py);
      row[2] = __float2bfloat16_rn(pz);
    } else {
      int k = u - 1;
      float f = (float)(k * k);  // embedding.py:47-49: linspace(0, N-1, N) ** 2
      float sx, cx, sy, cy, sz, cz;
      sincosf(__fmul_rn(px, f), &sx, &cx);
      sincosf(__fmul_rn(py, f), &sy, &cy);
      sincosf(__fmul_rn(pz, f), &sz, &cz);
      bf16* o = row + 3 + 6 * k;
      o[0] = __float2bfloat16_rn(sx);
      o[1] = __float2bfloat16_rn(sy);
      o[2] = __float2bfloat16_rn(sz);
      o[3] = __float2bfloat16_rn(cx);
      o[4] = __float2bfloat16_rn(cy);
      o[5] = __float2bfloat16_rn(cz);
    }
  }
}

// per-ray view-dir PE: dir_pe[b, :] = [d(3), {sin(k^2 d), cos(k^2 d)}_k, 0-pad], d = last 3 columns of the ray row
__global__ void encode_dirs_kernel(const float* __restrict__ rays, int ray_stride, bf16* __restrict__ dir_pe, int64_t b0,
                                   int64_t B, int n_freqs_dir, int dir_pad) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B * dir_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = t / dir_pad;
    int c = (int)(t - b * dir_pad);
    const float* d = rays + (b0 + b) * ray_stride + (ray_stride - 3);
    float v = 0.0f;
    int in_dir = 3 + 6 * n_freqs_dir;
    if (c < 3) v = d[c];
    else if (c < in_dir) {
      int q = c - 3, k = q / 6, rr = q - 6 * k, fn = rr / 3, dd = rr - 3 * fn;
      float a = __fmul_rn(d[dd], (float)(k * k));
      v = fn ? cosf(a) : sinf(a);
    }
    dir_pe[t] = __float2bfloat16_rn(v);
  }
}

// enc_kind 0: X0[p, :] = [x[p, 0:in_pos], 0-pad | x[p, in_pos:in_pos+in_dir], 0-pad]  (already-encoded fp32 input).
// One thread per 8 output columns: eight fp32 loads, one 16 B bf16 store.
__global__ void __launch_bounds__(256)
encode_copy_kernel(const float* __restrict__ x, bf16* __restrict__ x0, int64_t p0, int64_t npts, int in_pos, int in_dir,
                   int pos_pad, int dir_pad) {
  const int cols = pos_pad + dir_pad;
  const int groups = cols >> 3;
  const int in_tot = in_pos + in_dir;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts * groups; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lp = t / groups;
    const int c0 = (int)(t - lp * groups) << 3;
    const float* row = x + (p0 + lp) * in_tot;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      v[e] = c < in_pos ? __ldg(row + c) : ((c >= pos_pad && c - pos_pad < in_dir) ? __ldg(row + in_pos + (c - pos_pad)) : 0.0f);
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(v[0], v[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[2], v[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[4], v[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(v[6], v[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    *reinterpret_cast<uint4*>(x0 + lp * cols + c0) = o;
  }
}

// enc_kind 2: SinusoidalEncoding (encoding/sinusoidal.py:49-66) of raw coords x [P, in_dim] with bands[n_freqs]
__global__ void encode_sinusoidal_kernel(const float* __restrict__ x, const float* __restrict__ bands,
                                         bf16* __restrict__ x0, int64_t p0, int64_t npts, int in_dim, int n_freqs,
                                         int pos_pad) {
  const float half_pi = 1.57079637050628662109375f;
  const int half = in_dim * n_freqs;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < npts * pos_pad;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t lp = t / pos_pad;
    int c = (int)(t - lp * pos_pad);
    float v = 0.0f;
    if (c < 2 * half) {
      int q = c < half ? c : c - half;
      int d = q / n_freqs, k = q - d * n_freqs;
      float s = __fmul_rn(x[(p0 + lp) * in_dim + d], bands[k]);
      if (c >= half) s = __fadd_rn(s, half_pi);
      v = sinf(s);
    }
    x0[t] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------ small heads
// out[p, col0 + o] = h[p, :] . Wt[o, :] + b[o],  o < n_out <= 8.  Every lane loads 16 B (8 bf16 columns) of a point's row,
// so a point takes K/8 lanes and a warp iteration covers 256/K points with one coalesced 512 B transaction (K = 64: four
// points per warp instead of one with 24 idle lanes).  K in {64, 128, 256}.
__global__ void __launch_bounds__(256)
head_fwd_kernel(const bf16* __restrict__ h, int ldh, int K, const float* __restrict__ Wt, const float* __restrict__ b,
                int n_out, float* __restrict__ out, int ldo, int col0, int64_t P) {
  extern __shared__ float s_w[];  // [n_out][K]
  for (int i = threadIdx.x; i < n_out * K; i += blockDim.x) s_w[i] = Wt[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int lpp = K >> 3;        // lanes per point
  const int ppw = 32 / lpp;      // points per warp iteration
  const int sub = lane / lpp, cl = lane - sub * lpp;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pb = warp0 * ppw; pb < P; pb += nw * ppw) {
    const int64_t p = pb + sub;
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.0f;
    if (p < P) {
      const int k0 = cl * 8;
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(h + p * ldh + k0));
      const bf16* hv = reinterpret_cast<const bf16*>(&raw);
      float hf[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) hf[e] = __bfloat162float(hv[e]);
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        if (o < n_out) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + o * K + k0);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + o * K + k0 + 4);
          acc[o] = hf[0] * w0.x + hf[1] * w0.y + hf[2] * w0.z + hf[3] * w0.w + hf[4] * w1.x + hf[5] * w1.y + hf[6] * w1.z +
                   hf[7] * w1.w;
        }
      }
    }
    // sum over the lpp lanes of a point (lpp = 8, 16 or 32: xor offsets below lpp stay inside the group)
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < n_out) {
        for (int off = lpp >> 1; off > 0; off >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off);
      }
    }
    if (p < P && cl < n_out) {
      float v = 0.0f;
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o == cl) v = acc[o];
      out[p * ldo + col0 + cl] = v + b[cl];
    }
  }
}

// backward of a small head: d_out[p, col0 + o] given.
//   dW[o, k] += sum_p d_out[p,o] h[p,k] ; db[o] += sum_p d_out[p,o]
//   optionally d_h[p, k] = (sum_o d_out[p,o] W[o,k]) * (h[p,k] > 0)   (bf16; when the head input is post-ReLU)
// Every lane loads 16 B (8 bf16 columns) of a point's row, so a point takes K/8 lanes and a warp iteration covers
// 256/K points with one fully coalesced 512 B transaction; the dW partials stay in registers over the whole point loop
// as packed fp32 pairs (fma.rn.f32x2), and the block reduces them through shared memory into ONE set of atomics.
template <int K, int NOUT, bool HAS_DH>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const bf16* __restrict__ h, int ldh, const float* __restrict__ Wt, const float* __restrict__ d_out,
                int ldo, int col0, int64_t P, float* __restrict__ dW, float* __restrict__ db, bf16* __restrict__ d_h,
                int ldd) {
  constexpr int LPP = K / 8;     // lanes per point
  constexpr int PPW = 32 / LPP;  // points per warp iteration
  __shared__ float s_acc[NOUT * K + NOUT];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int sub = lane / LPP, cl = lane % LPP;
  for (int i = threadIdx.x; i < NOUT * K + NOUT; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
  float w[HAS_DH ? NOUT : 1][8];
  uint64_t acc[NOUT][4];
  float accb[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    accb[o] = 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[o][c] = 0ull;
    if (HAS_DH) {
#pragma unroll
      for (int c = 0; c < 8; ++c) w[o][c] = Wt[o * K + cl * 8 + c];
    }
  }
  const int64_t gw = ((int64_t)blockIdx.x * nwarps + warp) * PPW + sub;
  const int64_t gstride = (int64_t)gridDim.x * nwarps * PPW;
#pragma unroll 4
  for (int64_t p = gw; p < P; p += gstride) {
    const uint4 hv = __ldg(reinterpret_cast<const uint4*>(h + p * ldh + cl * 8));
    float d[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) d[o] = __ldg(d_out + p * ldo + col0 + o);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
    uint32_t ov[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float x0 = __uint_as_float(hw[c] << 16), x1 = __uint_as_float(hw[c] & 0xffff0000u);
      uint64_t xx;
      asm("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x0), "f"(x1));
#pragma unroll
      for (int o = 0; o < NOUT; ++o) {
        uint64_t dd;
        asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(d[o]));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[o][c]) : "l"(dd), "l"(xx));
      }
      if (HAS_DH) {
        float t0 = 0.0f, t1 = 0.0f;
#pragma unroll
        for (int o = 0; o < NOUT; ++o) {
          t0 += d[o] * w[o][2 * c];
          t1 += d[o] * w[o][2 * c + 1];
        }
        const __nv_bfloat162 r = __floats2bfloat162_rn(x0 > 0.0f ? t0 : 0.0f, x1 > 0.0f ? t1 : 0.0f);
        ov[c] = *reinterpret_cast<const uint32_t*>(&r);
      }
    }
#pragma unroll
    for (int o = 0; o < NOUT; ++o) accb[o] += d[o];
    if (HAS_DH) *reinterpret_cast<uint4*>(d_h + p * ldd + cl * 8) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float a0, a1;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc[o][c]));
      atomicAdd(&s_acc[o * K + cl * 8 + 2 * c], a0);
      atomicAdd(&s_acc[o * K + cl * 8 + 2 * c + 1], a1);
    }
    if (cl == 0) atomicAdd(&s_acc[NOUT * K + o], accb[o]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NOUT * K; i += blockDim.x) atomicAdd(dW + i, s_acc[i]);
  if (threadIdx.x < NOUT) atomicAdd(db + threadIdx.x, s_acc[NOUT * K + threadIdx.x]);
}

// Weight gradients of the feature layer and of the dir layer's feature part from G = d_hd^T h_{D-1} (no activation sits
// between feature_linear and list_linears_dir[0], models/NeRF.py:231-236, so the two Linears compose):
//   dW_dir[o, j]  = sum_k G[o,k] Wf[j,k] + db_dir[o] bf[j]      (= d_hd^T feature,   feature = h Wf^T + bf)
//   dW_feat[j, k] = sum_o Wd[o,j] G[o,k]                        (= d_feature^T h,    d_feature = d_hd Wd[:, :W])
//   db_feat[j]    = sum_o db_dir[o] Wd[o,j]
// with the bf16-rounded weights the tensor-core layers used.  Neither `feature` nor `d_feature` has to be kept in HBM.
// 32 x 32 output tiles, K staged through shared memory in chunks of 32 (two tiny fp32 GEMMs of 8.4 MFLOP each: the first
// version -- one block per output row with 256-iteration dependent loops -- took 30 us per pass, a fixed cost that is
// 3 % of a 1024-ray shard's step).  Blocks [0, tilesA): dW_dir tiles; the rest: dW_feat tiles (+ db_feat on k-tile 0).
__global__ void __launch_bounds__(256)
fold_feature_grads_kernel(const float* __restrict__ G, const float* __restrict__ db_dir, const float* __restrict__ Wf,
                          const float* __restrict__ bf, const float* __restrict__ Wd, int ldwd, int W, int Wh,
                          float* __restrict__ dWd, float* __restrict__ dWf, float* __restrict__ dbf) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 x 32 threads, 4 outputs each
  const int tiles_j = W / 32;
  const int tilesA = (Wh / 32) * tiles_j;
  auto bfr = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
  float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if ((int)blockIdx.x < tilesA) {  // dW_dir[o, j] += sum_k G[o,k] bf16(Wf[j,k]) + db_dir[o] bf[j]
    const int o0 = ((int)blockIdx.x / tiles_j) * 32, j0 = ((int)blockIdx.x % tiles_j) * 32;
    float ra[4], rb[4];  // next K chunk, loaded while the current one is multiplied (the loop is latency-bound)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ra[i] = G[(size_t)(o0 + ty * 4 + i) * W + tx];
      rb[i] = Wf[(size_t)(j0 + ty * 4 + i) * W + tx];
    }
    for (int k0 = 0; k0 < W; k0 += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sa[ty * 4 + i][tx] = ra[i];
        sb[ty * 4 + i][tx] = bfr(rb[i]);
      }
      __syncthreads();
      if (k0 + 32 < W) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ra[i] = G[(size_t)(o0 + ty * 4 + i) * W + k0 + 32 + tx];
          rb[i] = Wf[(size_t)(j0 + ty * 4 + i) * W + k0 + 32 + tx];
        }
      }
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const float w = sb[tx][kk];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += sa[ty * 4 + i][kk] * w;
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = o0 + ty * 4 + i, j = j0 + tx;
      dWd[(size_t)o * ldwd + j] += acc[i] + db_dir[o] * bf[j];
    }
  } else {  // dW_feat[j, k] += sum_o bf16(Wd[o,j]) G[o,k];  db_feat[j] += sum_o db_dir[o] bf16(Wd[o,j])
    const int t = (int)blockIdx.x - tilesA;
    const int j0 = (t / tiles_j) * 32, k0 = (t % tiles_j) * 32;
    float accb = 0.0f;
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ra[i] = Wd[(size_t)(ty * 4 + i) * ldwd + j0 + tx];
      rb[i] = G[(size_t)(ty * 4 + i) * W + k0 + tx];
    }
    for (int o0 = 0; o0 < Wh; o0 += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sa[ty * 4 + i][tx] = bfr(ra[i]);
        sb[ty * 4 + i][tx] = rb[i];
      }
      __syncthreads();
      if (o0 + 32 < Wh) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ra[i] = Wd[(size_t)(o0 + 32 + ty * 4 + i) * ldwd + j0 + tx];
          rb[i] = G[(size_t)(o0 + 32 + ty * 4 + i) * W + k0 + tx];
        }
      }
#pragma unroll
      for (int oo = 0; oo < 32; ++oo) {
        const float g = sb[oo][tx];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += sa[oo][ty * 4 + i] * g;
      }
      if (k0 == 0 && ty == 0) {
        for (int oo = 0; oo < 32; ++oo) accb += db_dir[o0 + oo] * sa[oo][tx];
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) dWf[(size_t)(j0 + ty * 4 + i) * W + k0 + tx] += acc[i];
    if (k0 == 0 && ty == 0) dbf[j0 + tx] += accb;
  }
}

template <int K>
int launch_head_bwd_nout(int n_out, const bf16* h, int ldh, const float* Wt, const float* d_out, int ldo, int col0,
                         int64_t P, float* dW, float* db, bf16* d_h, int ldd, cudaStream_t s) {
  const int blocks = kNumSMs * 4;  // streaming kernel: exactly the resident blocks
#define NMX_HB2(NO, DH) head_bwd_kernel<K, NO, DH><<<blocks, 256, 0, s>>>(h, ldh, Wt, d_out, ldo, col0, P, dW, db, d_h, ldd)
#define NMX_HB(NO)                                                         \
  case NO:                                                                 \
    if (d_h != nullptr) NMX_HB2(NO, true); else NMX_HB2(NO, false);        \
    break;
  switch (n_out) {
    NMX_HB(1) NMX_HB(2) NMX_HB(3) NMX_HB(4) NMX_HB(5) NMX_HB(6) NMX_HB(7) NMX_HB(8)
    default:
      set_error("head_bwd: n_out must be in [1,8]");
      return NMX_E_BADARG;
  }
#undef NMX_HB
#undef NMX_HB2
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_head_bwd(int K, int n_out, const bf16* h, int ldh, const float* Wt, const float* d_out, int ldo, int col0,
                    int64_t P, float* dW, float* db, bf16* d_h, int ldd, cudaStream_t s) {
  if ((ldh & 7) || (d_h != nullptr && (ldd & 7))) { set_error("head_bwd: leading dimensions must be multiples of 8"); return NMX_E_BADARG; }
  if (K == 256) return launch_head_bwd_nout<256>(n_out, h, ldh, Wt, d_out, ldo, col0, P, dW, db, d_h, ldd, s);
  if (K == 128) return launch_head_bwd_nout<128>(n_out, h, ldh, Wt, d_out, ldo, col0, P, dW, db, d_h, ldd, s);
  if (K == 64) return launch_head_bwd_nout<64>(n_out, h, ldh, Wt, d_out, ldo, col0, P, dW, db, d_h, ldd, s);
  set_error("head_bwd: K must be 64, 128 or 256 (got %d)", K);
  return NMX_E_UNSUPPORTED;
}

}  // namespace

// ================================================================================================ plan
extern "C" int64_t nmx_mlp_param_count(const nmx_mlp_config* c) {
  if (!c) return -1;
  int64_t n = 0;
  for (int l = 0; l < c->n_layers; ++l) {
    int in = (l == 0) ? c->in_pos : (c->skip_layer >= 0 && l == c->skip_layer + 1 ? c->width + c->in_pos : c->width);
    n += (int64_t)c->width * in + c->width;
  }
  if (c->use_viewdirs) {
    n += (int64_t)c->width * c->width + c->width;                    // feature
    n += c->width + 1;                                               // alpha
    n += (int64_t)(c->width / 2) * (c->width + c->in_dir) + c->width / 2;  // dir
    n += 3 * (c->width / 2) + 3;                                     // rgb
  } else {
    n += (int64_t)c->out_ch * c->width + c->out_ch;
  }
  return n;
}

extern "C" int nmx_mlp_plan_create(const nmx_mlp_config* c, int64_t max_points, nmx_mlp_plan** out) {
  NMX_CHECK_ARG(c && out && max_points > 0, "cfg, out non-null; max_points > 0");
  NMX_CHECK_ARG(c->n_layers >= 1 && c->n_layers <= 12, "1 <= n_layers <= 12");
  NMX_CHECK_ARG(c->width == 64 || c->width == 128 || c->width == 256, "width must be 64, 128 or 256");
  NMX_CHECK_ARG(!c->use_viewdirs || c->width % 128 == 0, "view-dir head needs width 128 or 256");
  NMX_CHECK_ARG(c->in_pos >= 1 && c->in_pos <= 256, "1 <= in_pos <= 256");
  NMX_CHECK_ARG(c->in_dir >= 0 && c->in_dir <= 64, "0 <= in_dir <= 64");
  NMX_CHECK_ARG(c->use_viewdirs ? c->in_dir > 0 : true, "view-dir head needs in_dir > 0");
  NMX_CHECK_ARG(c->use_viewdirs || (c->out_ch >= 1 && c->out_ch <= 8), "1 <= out_ch <= 8");
  NMX_CHECK_ARG(c->skip_layer < c->n_layers - 1, "skip_layer must be < n_layers - 1 (or -1)");
  nmx_mlp_plan* p = new nmx_mlp_plan();
  p->cfg = *c;
  p->max_points = (max_points + 127) / 128 * 128;  // activation regions hold whole 128-point tiles
  p->bits_word_major = 0;
  p->last_fwd_tiny = 0;
  p->D = c->n_layers;
  p->W = c->width;
  p->in_pos = c->in_pos;
  p->in_dir = c->use_viewdirs ? c->in_dir : 0;
  p->pos_pad = round_up(p->in_pos, 64);
  p->dir_pad = p->in_dir > 0 ? round_up(p->in_dir, 64) : 0;
  p->x0_cols = p->pos_pad + p->dir_pad;
  int64_t off = 0;
  auto lin = [&](int o, int i) {
    LinearRef r;
    r.out = o; r.in = i; r.w_off = off; off += (int64_t)o * i; r.b_off = off; off += o;
    return r;
  };
  for (int l = 0; l < p->D; ++l) {
    int in = (l == 0) ? p->in_pos : (c->skip_layer >= 0 && l == c->skip_layer + 1 ? p->W + p->in_pos : p->W);
    p->trunk.push_back(lin(p->W, in));
  }
  if (c->use_viewdirs) {
    p->feat = lin(p->W, p->W);
    p->alpha = lin(1, p->W);
    p->dir = lin(p->W / 2, p->W + p->in_dir);
    p->rgb = lin(3, p->W / 2);
  } else {
    p->outl = lin(c->out_ch, p->W);
  }
  p->n_params = off;
  // packed bf16 weights
  int64_t wb = 0;
  for (int l = 0; l < p->D; ++l) {
    int k = (l == 0) ? p->pos_pad : (c->skip_layer >= 0 && l == c->skip_layer + 1 ? p->pos_pad + p->W : p->W);
    p->wf_k.push_back(k);
    p->wf_off.push_back(wb); wb += align256((int64_t)p->W * k * 2);
    p->wt_off.push_back(wb); wb += align256((int64_t)p->W * p->W * 2);
  }
  if (c->use_viewdirs) {
    p->wf_feat = wb; wb += align256((int64_t)p->W * p->W * 2);
    p->wt_feat = wb; wb += align256((int64_t)p->W * p->W * 2);
    p->wf_dir_k = p->W + p->dir_pad;
    p->wf_dir = wb; wb += align256((int64_t)(p->W / 2) * p->wf_dir_k * 2);
    p->wt_dir = wb; wb += align256((int64_t)p->W * (p->W / 2) * 2);
  }
  p->wt_in = wb; wb += align256((int64_t)p->pos_pad * 2 * p->W * 2);
  // scratch for padded wgrad outputs (fp32 [W, 64]) and per-ray dir PE
  p->weights_bytes = align256(wb);
  *out = p;
  return 0;
}

extern "C" void nmx_mlp_plan_destroy(nmx_mlp_plan* plan) { delete plan; }

namespace {
// per-ray view-dir PE scratch: one row per ray of the current pass / inference chunk (rays <= points)
int64_t dirpe_bytes(const nmx_mlp_plan* p) {
  int64_t rows = p->max_points > infer_cap(p) ? p->max_points : infer_cap(p);
  return align256(rows * (int64_t)(p->dir_pad > 0 ? p->dir_pad : 1) * 2);
}
}

extern "C" int64_t nmx_mlp_workspace_bytes(const nmx_mlp_plan* p, int training) {
  if (!p) return -1;
  int64_t cap = training ? p->max_points : infer_cap(p);
  if (training) {  // a training workspace also serves inference passes (chunked layout must fit)
    int64_t ti = act_layout(p, infer_cap(p), false).total, tt = act_layout(p, cap, true).total;
    return p->weights_bytes + dirpe_bytes(p) + (ti > tt ? ti : tt) + 1024;
  }
  return p->weights_bytes + dirpe_bytes(p) + act_layout(p, cap, training != 0).total + 1024;
}

extern "C" int nmx_mlp_load_params(nmx_mlp_plan* p, const float* params, void* workspace, void* stream_) {
  NMX_CHECK_ARG(p && params && workspace, "plan, params, workspace non-null");
  cudaStream_t s = (cudaStream_t)stream_;
  uint8_t* ws = (uint8_t*)workspace;
  NMX_CUDA(cudaMemsetAsync(ws, 0, p->weights_bytes, s));  // zero padding columns (1.3 MB, cheaper than tracking)
  const int W = p->W;
  PackTable tab;
  tab.n = 0;
  auto add = [&](int64_t src_off, int src_ld, int src_col0, int64_t dst_off, int dst_ld, int dst_col0, int rows, int cols, int tr) {
    PackSeg& g = tab.seg[tab.n++];
    g.src_off = src_off; g.dst_off = dst_off; g.src_ld = src_ld; g.src_col0 = src_col0; g.dst_ld = dst_ld;
    g.dst_col0 = dst_col0; g.rows = rows; g.cols = cols; g.transpose = tr; g.pad_ = 0;
  };
  for (int l = 0; l < p->D; ++l) {
    const LinearRef& r = p->trunk[l];
    bool skip_in = (r.in == W + p->in_pos);
    if (l == 0) {
      add(r.w_off, r.in, 0, p->wf_off[l], p->wf_k[l], 0, W, p->in_pos, 0);
      add(r.w_off, r.in, 0, p->wt_in, 2 * W, 0, W, p->in_pos, 1);
    } else if (skip_in) {
      add(r.w_off, r.in, 0, p->wf_off[l], p->wf_k[l], 0, W, p->in_pos, 0);
      add(r.w_off, r.in, 0, p->wt_in, 2 * W, W, W, p->in_pos, 1);
      add(r.w_off, r.in, p->in_pos, p->wf_off[l], p->wf_k[l], p->pos_pad, W, W, 0);
      add(r.w_off, r.in, p->in_pos, p->wt_off[l], W, 0, W, W, 1);
    } else {
      add(r.w_off, r.in, 0, p->wf_off[l], p->wf_k[l], 0, W, W, 0);
      add(r.w_off, r.in, 0, p->wt_off[l], W, 0, W, W, 1);
    }
  }
  if (p->cfg.use_viewdirs) {
    add(p->feat.w_off, W, 0, p->wf_feat, W, 0, W, W, 0);
    add(p->feat.w_off, W, 0, p->wt_feat, W, 0, W, W, 1);
    add(p->dir.w_off, p->dir.in, 0, p->wf_dir, p->wf_dir_k, 0, W / 2, W, 0);
    add(p->dir.w_off, p->dir.in, W, p->wf_dir, p->wf_dir_k, W, W / 2, p->in_dir, 0);
    add(p->dir.w_off, p->dir.in, 0, p->wt_dir, W / 2, 0, W / 2, W, 1);
  }
  if (tab.n > kMaxPackSegs) { set_error("too many layers for the pack table"); return NMX_E_UNSUPPORTED; }
  pack_weights_kernel<<<dim3(32, tab.n), 256, 0, s>>>(params, ws, tab);
  NMX_LAUNCH_CHECK();
  return 0;
}

// ================================================================================================ forward
namespace {

struct Ctx {
  nmx_mlp_plan* p;
  uint8_t* ws;
  uint8_t* act;
  ActLayout al;
  const float* params;
  cudaStream_t s;
  bool training;
  int enc_kind;
  bf16* X0() const { return (bf16*)(act + al.x0); }
  bf16* H(int l) const { return (bf16*)(act + al.h0 + al.h_stride * (training ? l : (l & 1))); }
  bf16* FEAT() const { return (bf16*)(act + al.feat); }
  bf16* HD() const { return (bf16*)(act + al.hd); }
  bf16* G(int i) const { return (bf16*)(act + al.g0 + al.g_stride * i); }
  bf16* GHD() const { return (bf16*)(act + al.ghd); }
};

// the chain kernel can evaluate the Embedder PE itself (enc_kind 1 with the reference's 10 position bands)
bool enc_fused_ok(const nmx_mlp_plan* p, int enc_kind) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("NMX_DISABLE_FUSED_ENC");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  return !disabled && enc_kind == 1 && chain_eligible(p) && p->cfg.n_freqs_pos == 10 && p->in_pos == 63 &&
         (p->dir_pad == 0 || p->dir_pad == 64);
}

int encode_chunk(const Ctx& c, const float* x_or_rays, int ray_stride, const float* z, const float* bands, int64_t p0,
                 int64_t npts, int n) {
  nmx_mlp_plan* p = c.p;
  int kind = c.enc_kind;
  if (kind == 1 && enc_fused_ok(p, kind)) {  // only the per-ray view-dir table; positions are encoded inside the chain
    const int64_t b0 = p0 / n, b1 = (p0 + npts - 1) / n;
    bf16* dir_pe = (bf16*)(c.ws + p->weights_bytes);
    if (p->dir_pad > 0) {
      encode_dirs_kernel<<<grid_for((b1 - b0 + 1) * p->dir_pad, 256, 8), 256, 0, c.s>>>(
          x_or_rays, ray_stride, dir_pe, b0, b1 - b0 + 1, p->cfg.n_freqs_dir, p->dir_pad);
      NMX_LAUNCH_CHECK();
    }
    return 0;
  }
  if (kind == 1) {
    const int64_t b0 = p0 / n, b1 = (p0 + npts - 1) / n;
    bf16* dir_pe = (bf16*)(c.ws + p->weights_bytes);
    if (p->dir_pad > 0) {
      encode_dirs_kernel<<<grid_for((b1 - b0 + 1) * p->dir_pad, 256, 8), 256, 0, c.s>>>(
          x_or_rays, ray_stride, dir_pe, b0, b1 - b0 + 1, p->cfg.n_freqs_dir, p->dir_pad);
      NMX_LAUNCH_CHECK();
    }
    int units = p->cfg.n_freqs_pos + 2;
    int blocks = grid_for(npts * units, 256, 16);
    encode_rays_kernel<<<blocks, 256, 0, c.s>>>(x_or_rays, ray_stride, z, dir_pe, b0, c.X0(), p0, npts, n,
                                                p->cfg.n_freqs_pos, p->in_pos, p->pos_pad, p->dir_pad);
  } else if (kind == 2) {
    int in_dim = p->in_pos / (2 * p->cfg.n_freqs_pos);
    encode_sinusoidal_kernel<<<grid_for(npts * p->pos_pad, 256, 16), 256, 0, c.s>>>(
        x_or_rays, bands, c.X0(), p0, npts, in_dim, p->cfg.n_freqs_pos, p->pos_pad);
  } else {
    encode_copy_kernel<<<grid_for(npts * (p->x0_cols / 8), 256, 16), 256, 0, c.s>>>(x_or_rays, c.X0(), p0, npts, p->in_pos,
                                                                            p->in_dir, p->pos_pad, p->dir_pad);
  }
  NMX_LAUNCH_CHECK();
  return 0;
}

int forward_chunk(const Ctx& c, int64_t npts, float* out, int out_cols) {
  nmx_mlp_plan* p = c.p;
  const int W = p->W;
  int rc;
  for (int l = 0; l < p->D; ++l) {
    GemmDesc g{};
    bool skip_in = (p->trunk[l].in == W + p->in_pos);
    if (l == 0) {
      g.A0 = c.X0(); g.a0_cols = p->x0_cols; g.a0_ld = p->x0_cols; g.a0_col = 0; g.a0_k = p->pos_pad;
    } else if (skip_in) {
      g.A0 = c.X0(); g.a0_cols = p->x0_cols; g.a0_ld = p->x0_cols; g.a0_col = 0; g.a0_k = p->pos_pad;
      g.A1 = c.H(l - 1); g.a1_cols = W; g.a1_ld = W; g.a1_col = 0; g.a1_k = W;
    } else {
      g.A0 = c.H(l - 1); g.a0_cols = W; g.a0_ld = W; g.a0_col = 0; g.a0_k = W;
    }
    g.a0_rows = npts;
    g.B = c.ws + p->wf_off[l]; g.b_rows = W; g.b_cols = p->wf_k[l]; g.b_ld = p->wf_k[l]; g.b_col = 0;
    g.M = npts; g.N = W; g.bias = c.params + p->trunk[l].b_off; g.D = c.H(l); g.ldd = W; g.relu = 1;
    if ((rc = launch_gemm(g, c.s))) return rc;
  }
  const bf16* hl = c.H(p->D - 1);
  if (p->cfg.use_viewdirs) {
    // alpha head (NeRF.py:230) -> raw[:, 3]
    head_fwd_kernel<<<grid_for(npts, 8, 16), 256, W * sizeof(float), c.s>>>(hl, W, W, c.params + p->alpha.w_off,
                                                                           c.params + p->alpha.b_off, 1, out, out_cols, 3, npts);
    NMX_LAUNCH_CHECK();
    // feature (no activation, NeRF.py:231)
    GemmDesc g{};
    g.A0 = hl; g.a0_rows = npts; g.a0_cols = W; g.a0_ld = W; g.a0_k = W;
    g.B = c.ws + p->wf_feat; g.b_rows = W; g.b_cols = W; g.b_ld = W;
    g.M = npts; g.N = W; g.bias = c.params + p->feat.b_off; g.D = c.FEAT(); g.ldd = W; g.relu = 0;
    if ((rc = launch_gemm(g, c.s))) return rc;
    // dir layer on [feature, input_dir] (NeRF.py:232-236)
    GemmDesc d{};
    d.A0 = c.FEAT(); d.a0_rows = npts; d.a0_cols = W; d.a0_ld = W; d.a0_k = W;
    d.A1 = c.X0(); d.a1_cols = p->x0_cols; d.a1_ld = p->x0_cols; d.a1_col = p->pos_pad; d.a1_k = p->dir_pad;
    d.B = c.ws + p->wf_dir; d.b_rows = W / 2; d.b_cols = p->wf_dir_k; d.b_ld = p->wf_dir_k;
    d.M = npts; d.N = W / 2; d.bias = c.params + p->dir.b_off; d.D = c.HD(); d.ldd = W / 2; d.relu = 1;
    if ((rc = launch_gemm(d, c.s))) return rc;
    // rgb head (NeRF.py:238) -> raw[:, 0:3]
    head_fwd_kernel<<<grid_for(npts, 8, 16), 256, 3 * (W / 2) * sizeof(float), c.s>>>(
        c.HD(), W / 2, W / 2, c.params + p->rgb.w_off, c.params + p->rgb.b_off, 3, out, out_cols, 0, npts);
    NMX_LAUNCH_CHECK();
  } else {
    head_fwd_kernel<<<grid_for(npts, 8, 16), 256, p->cfg.out_ch * W * sizeof(float), c.s>>>(
        hl, W, W, c.params + p->outl.w_off, c.params + p->outl.b_off, p->cfg.out_ch, out, out_cols, 0, npts);
    NMX_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace

// ---- fused chain (nmx_chain.cu): whole MLP per 128-point tile, activations stay on chip
namespace {
bool chain_eligible(const nmx_mlp_plan* p) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("NMX_DISABLE_CHAIN");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (disabled) return false;
  const int layers = p->D + (p->cfg.use_viewdirs ? 2 : 0);
  return p->W == 256 && p->pos_pad == 64 && (p->dir_pad == 0 || p->dir_pad == 64) && layers <= kMaxChainLayers &&
         p->D >= 2;
}

struct EncIn { const float* rays; int ray_stride; const float* z; int64_t p0; int n; };

int forward_chain(const Ctx& c, int64_t npts, int64_t cap, float* out, int out_cols, const EncIn* enc) {
  nmx_mlp_plan* p = c.p;
  const int W = p->W, D = p->D;
  ChainMaps maps;
  ChainParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  int nl = 0;
  prm.pos_last_layer = 0;
  prm.dir_layer = -1;
  for (int l = 0; l < D; ++l, ++nl) {
    ChainLayerDesc& d = prm.L[nl];
    const bool skip_in = (p->trunk[l].in == W + p->in_pos);
    int ns = 0;
    if (l == 0) {
      d.src[ns++] = kSrcPos;
    } else {
      if (skip_in) { d.src[ns++] = kSrcPos; prm.pos_last_layer = nl; }
      for (int k = 0; k < 4; ++k) d.src[ns++] = k;
    }
    d.bits_row0 = (int)(l * cap);
    d.n_slabs = ns; d.N = W; d.relu = 1; d.bias_off = (int)p->trunk[l].b_off;
    d.feeds_next = (l < D - 1) || p->cfg.use_viewdirs;
    d.save_kind = c.training ? 1 : 0; d.save_row0 = (int)(l * cap);
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wf_off[l], W, p->wf_k[l], p->wf_k[l], 128))) return rc;
  }
  prm.head7_layer = D - 1;
  prm.rgb_layer = -1;
  if (p->cfg.use_viewdirs) {
    ChainLayerDesc& f = prm.L[nl];
    for (int k = 0; k < 4; ++k) f.src[k] = k;
    f.bits_row0 = -1;
    f.n_slabs = 4; f.N = W; f.relu = 0; f.bias_off = (int)p->feat.b_off; f.feeds_next = 1;
    f.save_kind = (c.training && !chain_bwd_eligible(p)) ? 1 : 0; f.save_row0 = (int)(D * cap);
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wf_feat, W, W, W, 128))) return rc;
    ++nl;
    ChainLayerDesc& d = prm.L[nl];
    for (int k = 0; k < 4; ++k) d.src[k] = k;
    d.src[4] = kSrcDir;
    d.bits_row0 = (int)(D * cap);
    d.n_slabs = 5; d.N = W / 2; d.relu = 1; d.bias_off = (int)p->dir.b_off; d.feeds_next = 0;
    d.save_kind = c.training ? 2 : 0; d.save_row0 = 0;
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wf_dir, W / 2, p->wf_dir_k, p->wf_dir_k, 128))) return rc;
    prm.dir_layer = nl;
    prm.rgb_layer = nl;
    ++nl;
    prm.head7_n = 1; prm.head7_w_off = (int)p->alpha.w_off; prm.head7_b_off = (int)p->alpha.b_off;
    prm.rgb_w_off = (int)p->rgb.w_off; prm.rgb_b_off = (int)p->rgb.b_off;
    prm.uses_dir = 1; prm.x0_dir_col = p->pos_pad;
  } else {
    prm.head7_n = p->cfg.out_ch; prm.head7_w_off = (int)p->outl.w_off; prm.head7_b_off = (int)p->outl.b_off;
    prm.uses_dir = 0;
  }
  for (int l = nl; l < kMaxChainLayers; ++l) maps.w[l] = maps.w[0];
  prm.n_layers = nl;
  prm.pos_prefetch_layer = prm.pos_last_layer + 2 < nl ? prm.pos_last_layer + 2 : nl - 1;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = experiment_env("NMX_CHAIN_DBG");
    prm.dbg = dbg;
  }
  prm.bits = (c.training && chain_bwd_eligible(p)) ? (uint32_t*)(c.act + c.al.bits) : nullptr;
  if (enc != nullptr && enc_fused_ok(p, c.enc_kind)) {
    prm.enc_fused = 1; prm.rays = enc->rays; prm.ray_stride = enc->ray_stride; prm.z = enc->z;
    prm.dir_pe = c.ws + p->weights_bytes; prm.p0 = enc->p0; prm.b0 = enc->p0 / enc->n; prm.n_per_ray = enc->n;
  }
  prm.P = (int)npts; prm.save = c.training ? 1 : 0; prm.params = c.params; prm.out = out; prm.out_cols = out_cols;
  if ((rc = make_tmap_bf16_2d(&maps.x0, c.X0(), npts, p->x0_cols, p->x0_cols, 128))) return rc;
  if (c.training) {
    prm.cap = (int)cap;
    if (NMX_DBG(prm, 32)) {
      if ((rc = make_tmap_bf16_2d(&maps.save, c.act + c.al.h0, (uint64_t)(D + 1) * cap * 4, 64, 64, 128))) return rc;
    } else if ((rc = make_tmap_bf16_2d(&maps.save, c.act + c.al.h0, (uint64_t)(D + 1) * cap, W, W, 128))) return rc;
    if (p->cfg.use_viewdirs) {
      if ((rc = make_tmap_bf16_2d(&maps.hd, c.HD(), npts, W / 2, W / 2, 128))) return rc;
    } else {
      maps.hd = maps.x0;
    }
  } else {
    maps.save = maps.x0;
    maps.hd = maps.x0;
  }
  return launch_chain_fwd(maps, prm, c.s);
}

bool chain_bwd_eligible(const nmx_mlp_plan* p) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("NMX_DISABLE_CHAIN_BWD");
    disabled = (e && e[0] == '1') ? 1 : 0;
  }
  return !disabled && chain_eligible(p) && p->cfg.use_viewdirs && p->dir_pad == 64;
}

// Fused data-gradient chain (nmx_chain.cu MODE 1): d_hd -> d_feature -> dY_{D-1} -> ... -> dY_0, every one kept in
// its G slot for the wgrad kernels.  Layer order: dir-layer dgrad, feature dgrad (+ alpha rank-1, mask h_{D-1}),
// trunk layers D-1 .. 1 (mask h_{l-1}).
int backward_chain(const Ctx& c, int64_t row0, int64_t P, int64_t cap, const float* d_out, int max_ctas, cudaStream_t st) {
  nmx_mlp_plan* p = c.p;
  const int W = p->W, D = p->D;
  ChainMaps maps;
  ChainParams prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  int nl = 0;
  {  // d_feature = d_hd . W_dir[:, 0:W]   (K = W/2 over the dir layer's outputs)
    ChainLayerDesc& d = prm.L[nl];
    d.src[0] = 0; d.src[1] = 1; d.n_slabs = 2; d.N = W; d.feeds_next = 1; d.epi = 0;
    d.save_kind = 0; d.save_row0 = 0;  // d_feature is consumed on chip only (its weight gradients fold, see below)
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wt_dir, W, W / 2, W / 2, 128))) return rc;
    ++nl;
  }
  {  // dY_{D-1} = (d_feature . W_feat + d_sigma (x) w_alpha) * [h_{D-1} > 0]
    ChainLayerDesc& d = prm.L[nl];
    for (int k = 0; k < 4; ++k) d.src[k] = k;
    d.n_slabs = 4; d.N = W; d.feeds_next = 1; d.epi = 2; d.bits_row0 = (int)((D - 1) * cap);
    d.save_kind = 1; d.save_row0 = (int)((D - 1) * cap);
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wt_feat, W, W, W, 128))) return rc;
    ++nl;
  }
  for (int l = D - 1; l >= 1; --l, ++nl) {  // dY_{l-1} = (dY_l . W_l[:, h part]) * [h_{l-1} > 0]
    ChainLayerDesc& d = prm.L[nl];
    for (int k = 0; k < 4; ++k) d.src[k] = k;
    d.n_slabs = 4; d.N = W; d.feeds_next = l > 1; d.epi = 1; d.bits_row0 = (int)((l - 1) * cap);
    d.save_kind = 1; d.save_row0 = (int)((l - 1) * cap);
    if ((rc = make_tmap_bf16_2d(&maps.w[nl], c.ws + p->wt_off[l], W, W, W, 128))) return rc;
  }
  for (int l = nl; l < kMaxChainLayers; ++l) maps.w[l] = maps.w[0];
  if (nl >= kMaxChainLayers) { set_error("backward chain: too many layers"); return NMX_E_UNSUPPORTED; }
  prm.n_layers = nl;
  prm.P = (int)P; prm.save = 1; prm.params = c.params; prm.out = nullptr; prm.out_cols = 4;
  prm.head7_layer = -1; prm.rgb_layer = -1; prm.head7_n = 1;
  prm.head7_w_off = (int)p->alpha.w_off; prm.rgb_w_off = (int)p->rgb.w_off;
  prm.uses_dir = 0; prm.pos_last_layer = -1; prm.pos_prefetch_layer = -1; prm.dir_layer = -1;
  prm.d_out = d_out + row0 * 4; prm.bits = (uint32_t*)(c.act + c.al.bits) + row0 * 8; prm.hd_bits_row0 = (int)(D * cap);
  prm.max_ctas = max_ctas;
  {
    static int dbg = -1;
    if (dbg < 0) dbg = experiment_env("NMX_CHAIN_DBG");
    prm.dbg = dbg & ~3;  // bits 2 (trace), 3 (no mask reads), 4 (no dY stores) apply to the backward chain
  }
  if ((rc = make_tmap_bf16_2d(&maps.save, c.G(0) + row0 * W, (uint64_t)(D + 1) * cap - row0, W, W, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&maps.hd, c.GHD() + row0 * (W / 2), P, W / 2, W / 2, 128))) return rc;
  maps.x0 = maps.hd;
  return launch_chain_bwd(maps, prm, st);
}

// Inference of the reference's 8 x 256 view-dir net straight from rays: CTA pairs, two tiles in ping-pong (nmx_chain2.cu);
// NMX_DISABLE_CHAIN2=1 falls back to the one-tile chain.
bool chain2_ok(const nmx_mlp_plan* p, int enc_kind, int n) {
  static int off = -1;
  if (off < 0) off = getenv("NMX_DISABLE_CHAIN2") ? 1 : 0;
  return !off && chain_eligible(p) && p->cfg.use_viewdirs && p->W == 256 && p->D == 8 && p->cfg.skip_layer == 4 &&
         p->pos_pad == 64 && p->dir_pad == 64 && enc_fused_ok(p, enc_kind) && n >= 8;
}

int forward_chain2(const Ctx& c, int64_t npts, float* out, const EncIn& enc) {
  const nmx_mlp_plan* p = c.p;
  Chain2Launch a;
  memset(&a, 0, sizeof(a));
  a.P = npts; a.params = c.params; a.out = out;
  for (int l = 0; l < 8; ++l) {
    a.w_ptr[l] = c.ws + p->wf_off[l]; a.w_k[l] = p->wf_k[l]; a.bias_off[l] = (int)p->trunk[l].b_off;
  }
  a.w_ptr[8] = c.ws + p->wf_feat; a.w_k[8] = p->W; a.bias_off[8] = (int)p->feat.b_off;
  a.w_ptr[9] = c.ws + p->wf_dir; a.w_k[9] = p->wf_dir_k; a.bias_off[9] = (int)p->dir.b_off;
  a.alpha_w_off = (int)p->alpha.w_off; a.alpha_b_off = (int)p->alpha.b_off;
  a.rgb_w_off = (int)p->rgb.w_off; a.rgb_b_off = (int)p->rgb.b_off;
  a.rays = enc.rays; a.ray_stride = enc.ray_stride; a.z = enc.z; a.p0 = enc.p0; a.n_per_ray = enc.n;
  a.n_freqs_dir = p->cfg.n_freqs_dir; a.dir_w_off = (int)p->dir.w_off; a.dir_ldw = p->dir.in;
  a.dir_bias = (float*)(c.ws + p->weights_bytes);  // per-ray scratch: 512 B per ray <= 128 B per point for n >= 4
  return launch_chain2(a, c.s);
}

// Training forward on CTA pairs (nmx_chain2t.cu): same net as chain2_ok, saved tensors in the one-tile chain's layouts.
// NMX_DISABLE_CHAIN2T=1 falls back to the one-tile training chain.
bool chain2t_ok(const nmx_mlp_plan* p, int enc_kind, int n, int64_t P) {
  static int off = -1;
  if (off < 0) off = getenv("NMX_DISABLE_CHAIN2T") ? 1 : 0;
  return !off && chain2_ok(p, enc_kind, n) && chain_bwd_eligible(p) && P >= 4096;
}

int forward_chain2_train(const Ctx& c, int64_t P, float* out, const EncIn& enc) {
  const nmx_mlp_plan* p = c.p;
  Chain2TrainLaunch a;
  memset(&a, 0, sizeof(a));
  a.P = P; a.params = c.params; a.out = out;
  for (int l = 0; l < 8; ++l) {
    a.w_ptr[l] = c.ws + p->wf_off[l]; a.w_k[l] = p->wf_k[l]; a.bias_off[l] = (int)p->trunk[l].b_off;
  }
  a.w_ptr[8] = c.ws + p->wf_feat; a.w_k[8] = p->W; a.bias_off[8] = (int)p->feat.b_off;
  a.w_ptr[9] = c.ws + p->wf_dir; a.w_k[9] = p->wf_dir_k; a.bias_off[9] = (int)p->dir.b_off;
  a.alpha_w_off = (int)p->alpha.w_off; a.alpha_b_off = (int)p->alpha.b_off;
  a.rgb_w_off = (int)p->rgb.w_off; a.rgb_b_off = (int)p->rgb.b_off;
  a.rays = enc.rays; a.ray_stride = enc.ray_stride; a.z = enc.z; a.n_per_ray = enc.n; a.in_dir = p->in_dir;
  a.n_freqs_dir = p->cfg.n_freqs_dir;
  a.dir_w_off = (int)p->dir.w_off; a.dir_ldw = p->dir.in;
  // per-ray scratch inside the dirpe_bytes region: [PE(dir) table: rays x 128 B | constants + dir-layer term: 512 B / ray]
  const int64_t n_rays = (P + enc.n - 1) / enc.n;
  uint8_t* base = c.ws + p->weights_bytes;
  a.dir_pe = base;
  const int64_t off = align256(n_rays * 128);
  if (off + chain2_train_scratch_bytes(n_rays) > dirpe_bytes(p)) { set_error("chain2 train: per-ray scratch does not fit"); return NMX_E_BADARG; }
  a.scratch = (float*)(base + off);
  a.save_base = c.act + c.al.h0; a.save_rows = (int64_t)(p->D + 1) * p->max_points; a.cap = p->max_points;
  a.hd = c.HD(); a.x0 = c.X0(); a.bits = (uint32_t*)(c.act + c.al.bits);
  return launch_chain2_train(a, c.s);
}

}  // namespace

// ---- fused width-64 MLP (nmx_tiny.cu): nets without view-dir head or skip connection on an already-encoded input
namespace {
bool tiny_ok(const nmx_mlp_plan* p, int enc_kind) {
  const bool off = getenv("NMX_DISABLE_TINY") != nullptr;  // read per call: the parity test runs both paths in one process
  return !off && enc_kind == 0 && p->W == 64 && !p->cfg.use_viewdirs && p->cfg.skip_layer < 0 && p->D >= 1 &&
         p->D <= kTinyMaxLayers && (p->in_pos == 32 || (p->in_pos == 64 && p->D <= 3)) && p->cfg.out_ch >= 1 &&
         p->cfg.out_ch <= 8;
}
TinyMlpDesc tiny_desc(const nmx_mlp_plan* p, const float* params, int64_t P) {
  TinyMlpDesc d;
  memset(&d, 0, sizeof(d));
  d.params = params;
  for (int l = 0; l < p->D; ++l) { d.w_off[l] = p->trunk[l].w_off; d.b_off[l] = p->trunk[l].b_off; }
  d.wo_off = p->outl.w_off; d.bo_off = p->outl.b_off;
  d.D = p->D; d.in_pos = p->in_pos; d.out_ch = p->cfg.out_ch; d.P = P;
  return d;
}
}  // namespace

extern "C" int nmx_mlp_input_grad_cols(const nmx_mlp_plan* p) {
  if (!p) return 0;
  return p->last_fwd_tiny ? p->in_pos : p->pos_pad;
}

extern "C" int nmx_mlp_fwd(nmx_mlp_plan* p, void* workspace, const float* params, int enc_kind,
                           const float* x_or_rays, int ray_stride, const float* z, const float* bands, float* out,
                           int64_t B, int n, int save_activations, void* stream_) {
  NMX_CHECK_ARG(p && workspace && params && x_or_rays && out, "plan, workspace, params, input, out non-null");
  NMX_CHECK_ARG(B >= 0 && n >= 1, "B >= 0, n >= 1");
  const int64_t P = B * n;
  if (P == 0) return 0;
  NMX_CHECK_ARG(!save_activations || P <= p->max_points, "B*n exceeds the plan's max_points (training capacity)");
  NMX_CHECK_ARG(enc_kind >= 0 && enc_kind <= 2, "enc_kind in {0,1,2}");
  NMX_CHECK_ARG(enc_kind != 1 || (z != nullptr && ray_stride >= 9), "enc_kind 1 needs z and ray_stride >= 9");
  NMX_CHECK_ARG(enc_kind != 1 || (p->in_pos == 3 + 6 * p->cfg.n_freqs_pos && (p->in_dir == 0 || p->in_dir == 3 + 6 * p->cfg.n_freqs_dir)),
                "enc_kind 1 needs in_pos == 3+6*n_freqs_pos and in_dir == 3+6*n_freqs_dir");
  NMX_CHECK_ARG(enc_kind != 2 || (bands != nullptr && p->cfg.n_freqs_pos > 0 && p->in_pos % (2 * p->cfg.n_freqs_pos) == 0),
                "enc_kind 2 needs bands and in_pos == 2*in_dim*n_freqs_pos");
  Ctx c;
  c.p = p; c.ws = (uint8_t*)workspace; c.params = params; c.s = (cudaStream_t)stream_; c.training = save_activations != 0;
  c.act = c.ws + p->weights_bytes + dirpe_bytes(p);
  const int out_cols = p->cfg.use_viewdirs ? 4 : p->cfg.out_ch;
  int rc;
  c.enc_kind = enc_kind;
  if (c.training) p->last_fwd_tiny = 0;
  if (tiny_ok(p, enc_kind)) {  // one launch, nothing but the bf16 input kept for the backward pass
    bf16* x0_save = nullptr;
    if (c.training) {
      c.al = act_layout(p, p->max_points, true);
      x0_save = c.X0();
      p->last_fwd_tiny = 1;
    }
    return launch_tiny_fwd(tiny_desc(p, params, P), x_or_rays, out, x0_save, p->x0_cols, c.s);
  }
  if (c.training) {
    c.al = act_layout(p, p->max_points, true);
    const bool pair_train = chain_eligible(p) && chain2t_ok(p, enc_kind, n, P);
    // (the pair kernel's own preparation launch builds the per-ray view-dir table: no encode_chunk)
    if (!pair_train && (rc = encode_chunk(c, x_or_rays, ray_stride, z, bands, 0, P, n))) return rc;
    if (chain_eligible(p)) {
      EncIn ei{x_or_rays, ray_stride, z, 0, n};
      p->bits_word_major = pair_train ? 1 : 0;
      if (pair_train) return forward_chain2_train(c, P, out, ei);
      return forward_chain(c, P, p->max_points, out, out_cols, &ei);
    }
    return forward_chunk(c, P, out, out_cols);
  }
  int64_t cap = infer_cap(p);
  c.al = act_layout(p, cap, false);
  if (chain2_ok(p, enc_kind, n)) {
    // The CTA-pair chain keeps nothing per point in the workspace: a launch is only limited by the per-ray scratch
    // (13 KB of constants + 512 B per ray in the dirpe_bytes region), so whole ray batches go in one launch.
    const int64_t max_rays = (dirpe_bytes(p) - 16384) / 512 - 2;
    int64_t cap2 = max_rays * n;
    if (cap2 > (int64_t)1 << 30) cap2 = (int64_t)1 << 30;
    cap2 -= cap2 % 256;
    for (int64_t p0 = 0; p0 < P; p0 += cap2) {
      const int64_t npts = P - p0 < cap2 ? P - p0 : cap2;
      EncIn ei{x_or_rays, ray_stride, z, p0, n};
      if ((rc = forward_chain2(c, npts, out + p0 * out_cols, ei))) return rc;
    }
    return 0;
  }
  for (int64_t p0 = 0; p0 < P; p0 += cap) {
    int64_t npts = P - p0 < cap ? P - p0 : cap;
    EncIn ei{x_or_rays, ray_stride, z, p0, n};
    if ((rc = encode_chunk(c, x_or_rays, ray_stride, z, bands, p0, npts, n))) return rc;
    if (chain_eligible(p)) rc = forward_chain(c, npts, cap, out + p0 * out_cols, out_cols, &ei);
    else rc = forward_chunk(c, npts, out + p0 * out_cols, out_cols);
    if (rc) return rc;
  }
  return 0;
}

// diagnostics: byte offsets of the training workspace regions (tests / scripts inspect saved tensors with them)
extern "C" int nmx_mlp_debug_layout(const nmx_mlp_plan* p, int64_t* out, int n) {
  NMX_CHECK_ARG(p && out && n >= 12, "plan, out non-null; n >= 12");
  const ActLayout a = act_layout(p, p->max_points, true);
  const int64_t base = p->weights_bytes + dirpe_bytes(p);
  const int64_t v[12] = {base, a.x0, a.h0, a.h_stride, a.feat, a.hd, a.g0, a.g_stride, a.ghd, a.bits, p->max_points,
                         (int64_t)p->x0_cols};
  for (int i = 0; i < 12; ++i) out[i] = v[i];
  return 0;
}

// ================================================================================================ backward
static int mlp_bwd_impl(nmx_mlp_plan* p, void* workspace, const float* params, const float* d_out, float* d_params,
                        float* d_input, int64_t P, void* stream_);

extern "C" int nmx_mlp_bwd(nmx_mlp_plan* p, void* workspace, const float* params, const float* d_out,
                           float* d_params, int64_t P, void* stream_) {
  return mlp_bwd_impl(p, workspace, params, d_out, d_params, nullptr, P, stream_);
}

extern "C" int nmx_mlp_bwd_input(nmx_mlp_plan* p, void* workspace, const float* params, const float* d_out,
                                 float* d_params, float* d_input, int64_t P, void* stream_) {
  NMX_CHECK_ARG(d_input != nullptr, "d_input non-null");
  return mlp_bwd_impl(p, workspace, params, d_out, d_params, d_input, P, stream_);
}

static int mlp_bwd_impl(nmx_mlp_plan* p, void* workspace, const float* params, const float* d_out, float* d_params,
                        float* d_input, int64_t P, void* stream_) {
  NMX_CHECK_ARG(p && workspace && params && d_out && d_params, "plan, workspace, params, d_out, d_params non-null");
  NMX_CHECK_ARG(P >= 0 && P <= p->max_points, "0 <= P <= max_points");
  cudaStream_t s = (cudaStream_t)stream_;
  NMX_CUDA(cudaMemsetAsync(d_params, 0, p->n_params * sizeof(float), s));
  if (P == 0) return 0;
  if (p->last_fwd_tiny) {  // fused width-64 backward: d_input is [P, in_pos] (nmx_mlp_input_grad_cols)
    Ctx c;
    c.p = p; c.ws = (uint8_t*)workspace; c.params = params; c.s = s; c.training = true;
    c.act = c.ws + p->weights_bytes + dirpe_bytes(p);
    c.al = act_layout(p, p->max_points, true);
    return launch_tiny_bwd(tiny_desc(p, params, P), c.X0(), p->x0_cols, d_out, d_params, d_input, s);
  }
  Ctx c;
  c.p = p; c.ws = (uint8_t*)workspace; c.params = params; c.s = s; c.training = true;
  c.act = c.ws + p->weights_bytes + dirpe_bytes(p);
  c.al = act_layout(p, p->max_points, true);
  const int W = p->W;
  const int out_cols = p->cfg.use_viewdirs ? 4 : p->cfg.out_ch;
  int rc;
  int cur = 0;  // index of the gradient buffer holding dY of the layer being processed
  const bf16* hl = c.H(p->D - 1);

  auto wgrad = [&](const bf16* dY, int dy_cols, const bf16* X, int x_cols, int x_col, int M, int N, int n_valid,
                   float* dW, int ldw, int w_col, float* db) {
    WgradDesc g{};
    g.dY = dY; g.dy_cols = dy_cols; g.dy_ld = dy_cols; g.dy_col = 0;
    g.X = X; g.x_cols = x_cols; g.x_ld = x_cols; g.x_col = x_col;
    g.P = P; g.M = M; g.N = N; g.dW = dW; g.ldw = ldw; g.w_col = w_col; g.n_valid = n_valid; g.db = db;
    return launch_wgrad(g, s);
  };

  // gradient w.r.t. the (already encoded) position inputs: d_x[rows, pos_pad] (+)= dY[rows, W] * W_l[:, :in_pos]
  // (fp32, columns >= in_pos are zero); k_col selects layer 0 (0) or the skip layer (W) inside the packed wt_in
  auto input_grad = [&](const bf16* dY, int64_t r0, int64_t rows, int k_col, bool accumulate) {
    GemmDesc g{};
    g.A0 = dY; g.a0_rows = rows; g.a0_cols = W; g.a0_ld = W; g.a0_k = W;
    g.B = c.ws + p->wt_in; g.b_rows = p->pos_pad; g.b_cols = 2 * W; g.b_ld = 2 * W; g.b_col = k_col;
    g.M = rows; g.N = p->pos_pad; g.D = d_input + r0 * p->pos_pad; g.ldd = p->pos_pad; g.out_fp32 = 1;
    g.accum = accumulate ? 1 : 0;
    return launch_gemm(g, s);
  };
  const int skip_l = p->cfg.skip_layer >= 0 ? p->cfg.skip_layer + 1 : -1;  // the layer whose input is [x_pos, h]
  if (d_input != nullptr && p->pos_pad > 256) { set_error("input gradient: encoded width <= 256"); return NMX_E_UNSUPPORTED; }

  static const bool dbg_sync = getenv("NMX_DEBUG_SYNC") != nullptr;
  if (dbg_sync) {
    cudaError_t e0 = cudaDeviceSynchronize();
    fprintf(stderr, "[nmx] bwd entry sync: %s | ws=%p act=%p params=%p d_out=%p d_params=%p P=%lld HD=%p GHD=%p hl=%p\n",
            cudaGetErrorString(e0), (void*)c.ws, (void*)c.act, (const void*)params, (const void*)d_out, (void*)d_params,
            (long long)P, (void*)c.HD(), (void*)c.GHD(), (const void*)hl);
  }
  if (chain_bwd_eligible(p)) {
    // NMX_BWD_CHUNKS=K > 1 (experiment) cuts the pass into K point chunks and runs wgrad(k) on a second stream, on its own
    // share of the SMs, while the chain works on chunk k+1.  Measured on B200 this mixing does NOT pay (K=4: 7.4 ms vs
    // 6.1 ms for the fine pass: per-SM bandwidth, 4x the wgrad flushes), so the default is the sequential schedule K = 1.
    static cudaStream_t s2_dev[64] = {};  // second stream of the chunked experiment, one per device, created on demand
    static std::vector<cudaEvent_t> evs_dev[64];
    static int n_chunks_cfg = -1, chain_sms_cfg = -1;
    if (n_chunks_cfg < 0) {
      const char* e1 = getenv("NMX_BWD_CHUNKS");
      const char* e2 = getenv("NMX_BWD_CHAIN_SMS");
      chain_sms_cfg = e2 ? atoi(e2) : 96;
      if (chain_sms_cfg < 8 || chain_sms_cfg > kNumSMs - 8) chain_sms_cfg = 96;
      n_chunks_cfg = e1 ? atoi(e1) : 1;
      if (n_chunks_cfg < 1) n_chunks_cfg = 1;
    }
    int dev_id = 0;
    NMX_CUDA(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= 64) dev_id = 0;
    if (n_chunks_cfg > 1 && s2_dev[dev_id] == nullptr) NMX_CUDA(cudaStreamCreateWithFlags(&s2_dev[dev_id], cudaStreamNonBlocking));
    cudaStream_t s2 = s2_dev[dev_id];
    std::vector<cudaEvent_t>& evs = evs_dev[dev_id];
    const int64_t tiles = (P + 127) / 128;
    int K = n_chunks_cfg;
    if (tiles < (int64_t)4 * kNumSMs * K) K = (int)(tiles / (4 * kNumSMs)) > 0 ? (int)(tiles / (4 * kNumSMs)) : 1;
    while (K > 1 && (int)evs.size() < K + 2) {
      cudaEvent_t e;
      NMX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      evs.push_back(e);
    }
    const int64_t chunk_rows = (tiles + K - 1) / K * 128;
    if (K > 1) {
      NMX_CUDA(cudaEventRecord(evs[K], s));  // d_params zeroed (and everything before) -> second stream may start
      NMX_CUDA(cudaStreamWaitEvent(s2, evs[K], 0));
    }
    // head weight gradients (their data gradients are produced inside the fused chain)
    if ((rc = launch_head_bwd(W / 2, 3, c.HD(), W / 2, params + p->rgb.w_off, d_out, out_cols, 0, P,
                              d_params + p->rgb.w_off, d_params + p->rgb.b_off, nullptr, 0, K > 1 ? s2 : s))) return rc;
    if ((rc = launch_head_bwd(W, 1, hl, W, params + p->alpha.w_off, d_out, out_cols, 3, P, d_params + p->alpha.w_off,
                              d_params + p->alpha.b_off, nullptr, 0, K > 1 ? s2 : s))) return rc;
    for (int k = 0; k < K; ++k) {
      const int64_t r0 = (int64_t)k * chunk_rows;
      const int64_t rows = P - r0 < chunk_rows ? P - r0 : chunk_rows;
      if (rows <= 0) break;
      const bool alone_chain = (K == 1 || k == 0);            // nothing to overlap with yet: use every SM
      const bool alone_wgrad = (K == 1 || k == K - 1);        // last chunk: the chain is finished
      static int pair_bwd_off = -1;
      if (pair_bwd_off < 0) pair_bwd_off = getenv("NMX_DISABLE_CHAIN2B") ? 1 : 0;
      if (K == 1 && !pair_bwd_off && chain2_ok(p, 1, 8) && P >= 4096) {
        // the whole pass on CTA pairs, two tiles in ping-pong (nmx_chain2t.cu)
        Chain2BwdLaunch a;
        memset(&a, 0, sizeof(a));
        a.P = P; a.params = params; a.d_out = d_out;
        a.w_ptr[0] = c.ws + p->wt_dir; a.w_k[0] = W / 2;
        a.w_ptr[1] = c.ws + p->wt_feat; a.w_k[1] = W;
        for (int l = p->D - 1; l >= 1; --l) { a.w_ptr[2 + (p->D - 1 - l)] = c.ws + p->wt_off[l]; a.w_k[2 + (p->D - 1 - l)] = W; }
        a.alpha_w_off = (int)p->alpha.w_off; a.rgb_w_off = (int)p->rgb.w_off;
        a.bits = (const uint32_t*)(c.act + c.al.bits); a.bits_word_major = p->bits_word_major;
        a.save_base = c.G(0); a.save_rows = (int64_t)(p->D + 1) * p->max_points; a.cap = p->max_points;
        a.ghd = c.GHD();
        if ((rc = launch_chain2_bwd(a, s))) return rc;
      } else {
        if (p->bits_word_major) { set_error("nmx_mlp_bwd: the forward ran on CTA pairs (word-major sign bits); the pair backward is disabled or ineligible"); return NMX_E_UNSUPPORTED; }
        if ((rc = backward_chain(c, r0, rows, p->max_points, d_out, alone_chain ? 0 : chain_sms_cfg, s))) return rc;
      }
      cudaStream_t sw = K > 1 ? s2 : s;
      if (K > 1) {
        NMX_CUDA(cudaEventRecord(evs[k], s));
        NMX_CUDA(cudaStreamWaitEvent(s2, evs[k], 0));
      }
      const int wg_ctas = alone_wgrad ? 0 : kNumSMs - chain_sms_cfg;
      auto wg = [&](const bf16* dY, int dy_cols, const bf16* X, int x_cols, int x_col, int M, int N, int n_valid,
                    float* dW, int ldw, int w_col, float* db) {
        WgradDesc g{};
        g.dY = dY + r0 * dy_cols; g.dy_cols = dy_cols; g.dy_ld = dy_cols; g.dy_col = 0;
        g.X = X + r0 * x_cols; g.x_cols = x_cols; g.x_ld = x_cols; g.x_col = x_col;
        g.P = rows; g.M = M; g.N = N; g.dW = dW; g.ldw = ldw; g.w_col = w_col; g.n_valid = n_valid; g.db = db;
        g.max_ctas = wg_ctas;
        return launch_wgrad(g, sw);
      };
      // weight gradients from the saved dY slots: dir layer over [feature | dir PE], feature layer, trunk layers
      float* dWd = d_params + p->dir.w_off;
      float* gfold = (float*)(c.act + c.al.gfold);  // G = d_hd^T h_{D-1} (+ db_dir = column sums of d_hd)
      if (k == 0) NMX_CUDA(cudaMemsetAsync(gfold, 0, (size_t)(W / 2) * W * sizeof(float), sw));
      // one pass over d_hd for both products: G (with h_{D-1}) and the dir-PE columns of dW_dir (second operand)
      auto wg2 = [&](const bf16* dY, int dy_cols, const bf16* X, int M, float* dW, int ldw, int w_col, float* db,
                     int x2_col, int n_valid2, float* dW2, int ldw2, int w2_col) {
        WgradDesc g{};
        g.dY = dY + r0 * dy_cols; g.dy_cols = dy_cols; g.dy_ld = dy_cols; g.dy_col = 0;
        g.X = X + r0 * W; g.x_cols = W; g.x_ld = W; g.x_col = 0;
        g.P = rows; g.M = M; g.N = W; g.dW = dW; g.ldw = ldw; g.w_col = w_col; g.n_valid = W; g.db = db;
        g.X2 = c.X0() + r0 * p->x0_cols; g.x2_cols = p->x0_cols; g.x2_ld = p->x0_cols; g.x2_col = x2_col;
        g.dW2 = dW2; g.ldw2 = ldw2; g.w2_col = w2_col; g.n_valid2 = n_valid2;
        g.max_ctas = wg_ctas;
        return launch_wgrad(g, sw);
      };
      const bool dual_ok = (W == 256) && p->dir_pad == 64 && p->pos_pad == 64 && !getenv("NMX_DISABLE_DUAL_WGRAD");
      // One launch for all layers pays off where the per-launch fixed cost (148 partial-tile flushes + tail, ~11 us x 18)
      // matters: small passes (a 1024-ray shard).  Large passes stream each layer at the HBM roofline either way.
      static int64_t batch_max_points = -1;
      if (batch_max_points < 0) {
        const char* e = getenv("NMX_WGRAD_BATCH_MAX_POINTS");
        // measured on B200: faster up to a 262 144-point pass (C2: 1.286 -> 1.272 ms), slower at 524 288 (C3 coarse pass:
        // 9.85 -> 9.88 ms per step) and beyond
        batch_max_points = e ? atoll(e) : 300000;
        if (getenv("NMX_DISABLE_WGRAD_BATCH")) batch_max_points = 0;
      }
      const bool batched = dual_ok && K == 1 && P <= batch_max_points && p->D + 1 <= kMaxWgradJobs;
      if (batched) {
        // every weight gradient of the pass in ONE launch: the SMs are divided among the layers (nmx_wgrad_batch.cu)
        WgradBatchDesc bd;
        memset(&bd, 0, sizeof(bd));
        const int64_t cap = p->max_points;
        bd.n_tensors = 4;
        bd.t[0] = {c.act + c.al.h0, (int64_t)(p->D + 1) * cap, W};      // saved activations h_0 .. h_{D-1} (+ feature slot)
        bd.t[1] = {c.G(0), (int64_t)(p->D + 1) * cap, W};                // saved data gradients dY_0 .. dY_{D-1}
        bd.t[2] = {c.X0(), P, p->x0_cols};                               // encoded inputs [PE(pos) | PE(dir)]
        bd.t[3] = {c.GHD(), P, W / 2};                                   // d_hd
        bd.P = P;
        int nj = 0;
        {  // dir layer: G = d_hd^T h_{D-1} (folded below) and the dir-PE columns of dW_dir from one read of d_hd
          WgradBatchJob& g = bd.job[nj++];
          g.dy_t = 3; g.dy_row0 = 0; g.dy_col = 0; g.x_t = 0; g.x_row0 = (int64_t)(p->D - 1) * cap; g.x_col = 0;
          g.x2_t = 2; g.x2_row0 = 0; g.x2_col = p->pos_pad;
          g.M = W / 2; g.N = W; g.dW = gfold; g.ldw = W; g.w_col = 0; g.n_valid = W; g.db = d_params + p->dir.b_off;
          g.dW2 = dWd; g.ldw2 = p->dir.in; g.w2_col = W; g.n_valid2 = p->in_dir;
        }
        for (int l = p->D - 1; l >= 0; --l) {
          const LinearRef& r = p->trunk[l];
          WgradBatchJob& g = bd.job[nj++];
          g.dy_t = 1; g.dy_row0 = (int64_t)l * cap; g.dy_col = 0; g.M = W; g.x2_t = -1;
          g.dW = d_params + r.w_off; g.ldw = r.in; g.db = d_params + r.b_off;
          if (l == 0) {
            g.x_t = 2; g.x_row0 = 0; g.x_col = 0; g.N = p->pos_pad; g.w_col = 0; g.n_valid = p->in_pos;
          } else {
            g.x_t = 0; g.x_row0 = (int64_t)(l - 1) * cap; g.x_col = 0; g.N = W; g.n_valid = W;
            if (r.in == W + p->in_pos) {  // skip layer: [x_pos | h] against one read of dY
              g.w_col = p->in_pos;
              g.x2_t = 2; g.x2_row0 = 0; g.x2_col = 0; g.dW2 = g.dW; g.ldw2 = r.in; g.w2_col = 0; g.n_valid2 = p->in_pos;
            } else {
              g.w_col = 0;
            }
          }
        }
        bd.n_jobs = nj;
        if ((rc = launch_wgrad_batch(bd, sw))) return rc;
      } else if (dual_ok) {
        if ((rc = wg2(c.GHD(), W / 2, hl, W / 2, gfold, W, 0, d_params + p->dir.b_off, p->pos_pad, p->in_dir, dWd, p->dir.in, W))) return rc;
      } else {
        if ((rc = wg(c.GHD(), W / 2, hl, W, 0, W / 2, W, W, gfold, W, 0, d_params + p->dir.b_off))) return rc;
        if ((rc = wg(c.GHD(), W / 2, c.X0(), p->x0_cols, p->pos_pad, W / 2, p->dir_pad, p->in_dir, dWd, p->dir.in, W, nullptr))) return rc;
      }
      for (int l = p->D - 1; l >= 0 && !batched; --l) {
        const LinearRef& r = p->trunk[l];
        const bf16* dY = c.G(l);
        float* dW = d_params + r.w_off;
        float* db = d_params + r.b_off;
        const bool skip_in = (r.in == W + p->in_pos);
        if (l == 0) {
          if ((rc = wg(dY, W, c.X0(), p->x0_cols, 0, W, p->pos_pad, p->in_pos, dW, r.in, 0, db))) return rc;
        } else if (skip_in && dual_ok) {  // skip layer: [x_pos | h] operands against one read of dY
          if ((rc = wg2(dY, W, c.H(l - 1), W, dW, r.in, p->in_pos, db, 0, p->in_pos, dW, r.in, 0))) return rc;
        } else if (skip_in) {
          if ((rc = wg(dY, W, c.X0(), p->x0_cols, 0, W, p->pos_pad, p->in_pos, dW, r.in, 0, nullptr))) return rc;
          if ((rc = wg(dY, W, c.H(l - 1), W, 0, W, W, W, dW, r.in, p->in_pos, db))) return rc;
        } else {
          if ((rc = wg(dY, W, c.H(l - 1), W, 0, W, W, W, dW, r.in, 0, db))) return rc;
        }
      }
    }
    if (d_input != nullptr) {  // dY_0 and the skip layer's dY are still in their slots
      bool acc = false;
      if (skip_l > 0 && skip_l < p->D) {
        if ((rc = input_grad(c.G(skip_l), 0, P, W, false))) return rc;
        acc = true;
      }
      if ((rc = input_grad(c.G(0), 0, P, 0, acc))) return rc;
    }
    {  // feature / dir-layer weight gradients from G (after every chunk's wgrad has been accumulated)
      cudaStream_t sw = K > 1 ? s2 : s;
      fold_feature_grads_kernel<<<(W / 64) * (W / 32) + (W / 32) * (W / 32), 256, 0, sw>>>(
          (const float*)(c.act + c.al.gfold), d_params + p->dir.b_off, params + p->feat.w_off, params + p->feat.b_off,
          params + p->dir.w_off, p->dir.in, W, W / 2, d_params + p->dir.w_off, d_params + p->feat.w_off,
          d_params + p->feat.b_off);
      NMX_LAUNCH_CHECK();
    }
    if (K > 1) {  // join: the caller's stream continues only after the second stream's last wgrad
      NMX_CUDA(cudaEventRecord(evs[K + 1], s2));
      NMX_CUDA(cudaStreamWaitEvent(s, evs[K + 1], 0));
    }
    return 0;
  }
  static int64_t l2_rows = -1;
  if (l2_rows < 0) { const char* e = getenv("NMX_BWD_L2_ROWS"); l2_rows = e ? atoll(e) : 0; }
  static int64_t noview_batch_max = -1;  // same cross-over as the chain path's batched launch
  if (noview_batch_max < 0) {
    const char* e = getenv("NMX_WGRAD_BATCH_MAX_POINTS");
    noview_batch_max = getenv("NMX_DISABLE_WGRAD_BATCH") ? 0 : (e ? atoll(e) : 300000);
  }
  if (!p->cfg.use_viewdirs && P <= noview_batch_max && l2_rows == 0 && p->D <= kMaxWgradJobs) {
    // Nets without a view-dir head (image learning, the hash grid's tiny MLP), layer by layer: the data-gradient GEMMs
    // keep every dY in its own slot and ALL weight gradients run as one batched launch afterwards (was one 8-10 us
    // launch per layer: the C1 step is launch-bound).
    if ((rc = launch_head_bwd(W, p->cfg.out_ch, hl, W, params + p->outl.w_off, d_out, out_cols, 0, P, d_params + p->outl.w_off,
                              d_params + p->outl.b_off, c.G(p->D - 1), W, s))) return rc;
    for (int l = p->D - 1; l >= 1; --l) {  // dY_{l-1} = (dY_l W_l[:, h part]) * [h_{l-1} > 0]
      GemmDesc g{};
      g.A0 = c.G(l); g.a0_rows = P; g.a0_cols = W; g.a0_ld = W; g.a0_k = W;
      g.B = c.ws + p->wt_off[l]; g.b_rows = W; g.b_cols = W; g.b_ld = W;
      g.M = P; g.N = W; g.D = c.G(l - 1); g.ldd = W; g.mask = c.H(l - 1); g.ldmask = W;
      if ((rc = launch_gemm(g, s))) return rc;
    }
    if (d_input != nullptr) {
      bool acc = false;
      if (skip_l > 0 && skip_l < p->D) {
        if ((rc = input_grad(c.G(skip_l), 0, P, W, false))) return rc;
        acc = true;
      }
      if ((rc = input_grad(c.G(0), 0, P, 0, acc))) return rc;
    }
    WgradBatchDesc bd;
    memset(&bd, 0, sizeof(bd));
    bd.P = P;
    int nt = 0;
    bd.t[nt++] = {c.X0(), P, p->x0_cols};  // tensor 0: encoded inputs; then one tensor per dY / activation slot (rows = P)
    for (int l = p->D - 1; l >= 0; --l) {
      const LinearRef& r = p->trunk[l];
      WgradBatchJob& g = bd.job[bd.n_jobs++];
      bd.t[nt] = {c.G(l), P, W};
      g.dy_t = nt++; g.dy_row0 = 0; g.dy_col = 0; g.M = W; g.x2_t = -1;
      g.dW = d_params + r.w_off; g.ldw = r.in; g.db = d_params + r.b_off;
      if (l == 0) {
        g.x_t = 0; g.x_row0 = 0; g.x_col = 0; g.N = p->pos_pad; g.w_col = 0; g.n_valid = p->in_pos;
      } else {
        bd.t[nt] = {c.H(l - 1), P, W};
        g.x_t = nt++; g.x_row0 = 0; g.x_col = 0; g.N = W; g.n_valid = W;
        if (r.in == W + p->in_pos) {  // skip layer: [x_pos | h] against one read of dY
          g.w_col = p->in_pos;
          g.x2_t = 0; g.x2_row0 = 0; g.x2_col = 0; g.dW2 = g.dW; g.ldw2 = r.in; g.w2_col = 0; g.n_valid2 = p->in_pos;
        } else {
          g.w_col = 0;
        }
      }
    }
    bd.n_tensors = nt;
    return launch_wgrad_batch(bd, s);
  }
  // Layer-by-layer path (nets the fused chain does not cover).  With NMX_BWD_L2_ROWS=R the pass runs over windows of
  // R points so that the two ping-pong gradient buffers (always the SAME first R rows) stay L2-resident.
  auto run_window = [&](int64_t r0, int64_t rows) -> int {
    const bf16* hl_r = hl + r0 * W;
    const float* d_out_r = d_out + r0 * out_cols;
    int cur = 0;
    auto wgrad = [&](const bf16* dY, int dy_cols, const bf16* X, int x_cols, int x_col, int M, int N, int n_valid,
                     float* dW, int ldw, int w_col, float* db) {
      WgradDesc g{};
      g.dY = dY; g.dy_cols = dy_cols; g.dy_ld = dy_cols; g.dy_col = 0;
      g.X = X; g.x_cols = x_cols; g.x_ld = x_cols; g.x_col = x_col;
      g.P = rows; g.M = M; g.N = N; g.dW = dW; g.ldw = ldw; g.w_col = w_col; g.n_valid = n_valid; g.db = db;
      return launch_wgrad(g, s);
    };
    if (p->cfg.use_viewdirs) {
      // rgb head: d_hd_pre = (d_rgb W_rgb) * [hd > 0]; dW_rgb, db_rgb
      if ((rc = launch_head_bwd(W / 2, 3, (c.HD() + r0 * (W / 2)), W / 2, params + p->rgb.w_off, d_out_r, out_cols, 0, rows,
                                d_params + p->rgb.w_off, d_params + p->rgb.b_off, c.GHD(), W / 2, s))) return rc;
      // alpha head: dW_alpha, db_alpha (its d_h is the rank-1 term of the feature dgrad epilogue)
      if ((rc = launch_head_bwd(W, 1, hl_r, W, params + p->alpha.w_off, d_out_r, out_cols, 3, rows, d_params + p->alpha.w_off,
                                d_params + p->alpha.b_off, nullptr, 0, s))) return rc;
      // dir layer: wgrad over [feature | dir PE], bias, dgrad to feature
      float* dWd = d_params + p->dir.w_off;
      if ((rc = wgrad(c.GHD(), W / 2, (c.FEAT() + r0 * W), W, 0, W / 2, W, W, dWd, p->dir.in, 0, d_params + p->dir.b_off))) return rc;
      if ((rc = wgrad(c.GHD(), W / 2, (c.X0() + r0 * p->x0_cols), p->x0_cols, p->pos_pad, W / 2, p->dir_pad, p->in_dir, dWd, p->dir.in, W, nullptr))) return rc;
      GemmDesc g{};
      g.A0 = c.GHD(); g.a0_rows = rows; g.a0_cols = W / 2; g.a0_ld = W / 2; g.a0_k = W / 2;
      g.B = c.ws + p->wt_dir; g.b_rows = W; g.b_cols = W / 2; g.b_ld = W / 2;
      g.M = rows; g.N = W; g.D = c.G(0); g.ldd = W;
      if ((rc = launch_gemm(g, s))) return rc;
      // feature layer: wgrad, bias, dgrad (+ alpha rank-1 term, ReLU mask of h_{D-1})
      if ((rc = wgrad(c.G(0), W, hl_r, W, 0, W, W, W, d_params + p->feat.w_off, W, 0, d_params + p->feat.b_off))) return rc;
      GemmDesc f{};
      f.A0 = c.G(0); f.a0_rows = rows; f.a0_cols = W; f.a0_ld = W; f.a0_k = W;
      f.B = c.ws + p->wt_feat; f.b_rows = W; f.b_cols = W; f.b_ld = W;
      f.M = rows; f.N = W; f.D = c.G(1); f.ldd = W; f.mask = hl_r; f.ldmask = W;
      f.row_vec = d_out_r + 3; f.row_stride = out_cols; f.col_vec = params + p->alpha.w_off;
      if ((rc = launch_gemm(f, s))) return rc;
      cur = 1;
    } else {
      // no-view head: d_h = (d_out W_out) * [h > 0]; dW_out, db_out
      if ((rc = launch_head_bwd(W, p->cfg.out_ch, hl_r, W, params + p->outl.w_off, d_out_r, out_cols, 0, rows,
                                d_params + p->outl.w_off, d_params + p->outl.b_off, c.G(0), W, s))) return rc;
      cur = 0;
    }
    for (int l = p->D - 1; l >= 0; --l) {
      const LinearRef& r = p->trunk[l];
      const bf16* dY = c.G(cur);
      float* dW = d_params + r.w_off;
      bool skip_in = (r.in == W + p->in_pos);
      float* db = d_params + r.b_off;
      if (l == 0) {
        if ((rc = wgrad(dY, W, (c.X0() + r0 * p->x0_cols), p->x0_cols, 0, W, p->pos_pad, p->in_pos, dW, r.in, 0, db))) return rc;
      } else if (skip_in) {
        if ((rc = wgrad(dY, W, (c.X0() + r0 * p->x0_cols), p->x0_cols, 0, W, p->pos_pad, p->in_pos, dW, r.in, 0, nullptr))) return rc;
        if ((rc = wgrad(dY, W, (c.H(l - 1) + r0 * W), W, 0, W, W, W, dW, r.in, p->in_pos, db))) return rc;
      } else {
        if ((rc = wgrad(dY, W, (c.H(l - 1) + r0 * W), W, 0, W, W, W, dW, r.in, 0, db))) return rc;
      }
      if (d_input != nullptr) {
        if (l == skip_l && l > 0) { if ((rc = input_grad(dY, r0, rows, W, false))) return rc; }
        if (l == 0) { if ((rc = input_grad(dY, r0, rows, 0, skip_l > 0 && skip_l < p->D))) return rc; }
      }
      if (l >= 1) {
        GemmDesc g{};
        g.A0 = dY; g.a0_rows = rows; g.a0_cols = W; g.a0_ld = W; g.a0_k = W;
        g.B = c.ws + p->wt_off[l]; g.b_rows = W; g.b_cols = W; g.b_ld = W;
        g.M = rows; g.N = W; g.D = c.G(cur ^ 1); g.ldd = W; g.mask = (c.H(l - 1) + r0 * W); g.ldmask = W;
        if ((rc = launch_gemm(g, s))) return rc;
        cur ^= 1;
      }
    }
    return 0;
  };
  if (l2_rows > 0 && l2_rows % 128 == 0) {
    for (int64_t r0 = 0; r0 < P; r0 += l2_rows)
      if ((rc = run_window(r0, P - r0 < l2_rows ? P - r0 : l2_rows))) return rc;
    return 0;
  }
  return run_window(0, P);
}
