// K2c: multiresolution hash-grid encoding (MultiHashEncoding, encoding/multi_hash.py:13-137), canonical
// semantics of SURVEY 8a row 9 / DESIGN.md:
//   p = x * N_l (fp32) ; corners from (floor(p), ceil(p)) as int32 ; per-level table l ;
//   h = (cx*1 ^ cy*2654435761 ^ cz*805459861) & (T-1) in uint32 wraparound ; ALL levels hashed ;
//   offset = p - floor(p) weights the CEIL corner ; interpolation order x (03,12,56,47) -> y -> z as :123-131.
//
// Mapping: one thread per (point, level), level fastest.  The 16 lanes of a point share its 12-byte position
// (one broadcast load), each lane gathers its level's 8 corners as vector loads of F floats (8 B for F=2), and
// the [P, L*F] output row is written as one fully coalesced run per point.  Gradient scatter uses the
// vectorised red.global.add.v2.f32 / v4.f32 (atomicAdd(float2* / float4*)): one L2 atomic per corner, or per x edge when
// its two corners share an aligned 16-byte slot; warps whose lanes share a cell (coarse levels, samples of one ray) first
// merge lanes that hit the same table entry.
// HBM/L2 roofline: fwd 1164 B/pt, bwd 2188 B/pt at L=16, F=2.
#include "nmx_common.cuh"

using namespace nmx;

namespace {

__device__ __forceinline__ uint32_t hash3(int32_t x, int32_t y, int32_t z, uint32_t mask) {
  return (((uint32_t)x * 1u) ^ ((uint32_t)y * 2654435761u) ^ ((uint32_t)z * 805459861u)) & mask;
}

__global__ void hash_kernel(const int32_t* __restrict__ c, int32_t* __restrict__ idx, int64_t M, uint32_t mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
    idx[i] = (int32_t)hash3(c[i * 3 + 0], c[i * 3 + 1], c[i * 3 + 2], mask);
}

struct Corners {
  uint32_t idx[8];
  float ox, oy, oz;
};

// corner order of multi_hash.py:102-109: which of (x,y,z) takes the ceil coordinate
__device__ __forceinline__ void corners_of(float px, float py, float pz, uint32_t mask, Corners& c) {
  float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  int32_t x0 = (int32_t)fx, y0 = (int32_t)fy, z0 = (int32_t)fz;
  int32_t x1 = (int32_t)ceilf(px), y1 = (int32_t)ceilf(py), z1 = (int32_t)ceilf(pz);
  c.idx[0] = hash3(x1, y1, z1, mask);
  c.idx[1] = hash3(x1, y0, z1, mask);
  c.idx[2] = hash3(x0, y0, z1, mask);
  c.idx[3] = hash3(x0, y1, z1, mask);
  c.idx[4] = hash3(x1, y1, z0, mask);
  c.idx[5] = hash3(x1, y0, z0, mask);
  c.idx[6] = hash3(x0, y0, z0, mask);
  c.idx[7] = hash3(x0, y1, z0, mask);
  c.ox = __fsub_rn(px, fx);
  c.oy = __fsub_rn(py, fy);
  c.oz = __fsub_rn(pz, fz);
}

// a*t + b*(1-t) without FMA contraction (bit-parity with the fp32 oracle)
__device__ __forceinline__ float lerp_ref(float a, float b, float t, float omt) {
  return __fadd_rn(__fmul_rn(a, t), __fmul_rn(b, omt));
}

template <int F>
struct Vec;
template <>
struct Vec<1> { using T = float; };
template <>
struct Vec<2> { using T = float2; };
template <>
struct Vec<4> { using T = float4; };

template <int F>
__global__ void __launch_bounds__(256)
hashgrid_fwd_kernel(const float* __restrict__ x, const float* __restrict__ tables, const float* __restrict__ res,
                    float* __restrict__ out, int32_t* __restrict__ idx_out, int64_t total, int L, uint32_t mask) {
  using V = typename Vec<F>::T;
  const size_t T = (size_t)mask + 1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = t / L;
    int l = (int)(t - p * L);
    float r = __ldg(res + l);
    float px = __fmul_rn(__ldg(x + p * 3 + 0), r);
    float py = __fmul_rn(__ldg(x + p * 3 + 1), r);
    float pz = __fmul_rn(__ldg(x + p * 3 + 2), r);
    Corners c;
    corners_of(px, py, pz, mask, c);
    const V* tab = reinterpret_cast<const V*>(tables) + (size_t)l * T;
    V hv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) hv[k] = __ldg(tab + c.idx[k]);
    // (Fetching the two corners of an x edge with one 16-byte load when they share an aligned slot -- x0 even, see the
    // backward kernel -- was measured and is 7 % SLOWER here: 99 vs 91 us per 262 144 points; the extra selects cost
    // more than the saved L1 requests.  The same pairing halves the atomics of the backward pass and stays there.)
    if (idx_out != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) idx_out[t * 8 + k] = (int32_t)c.idx[k];
    }
    float ox1 = __fsub_rn(1.0f, c.ox), oy1 = __fsub_rn(1.0f, c.oy), oz1 = __fsub_rn(1.0f, c.oz);
    float o[F];
    const float* h = reinterpret_cast<const float*>(hv);
#pragma unroll
    for (int f = 0; f < F; ++f) {
      float h03 = lerp_ref(h[0 * F + f], h[3 * F + f], c.ox, ox1);
      float h12 = lerp_ref(h[1 * F + f], h[2 * F + f], c.ox, ox1);
      float h56 = lerp_ref(h[5 * F + f], h[6 * F + f], c.ox, ox1);
      float h47 = lerp_ref(h[4 * F + f], h[7 * F + f], c.ox, ox1);
      float h0312 = lerp_ref(h03, h12, c.oy, oy1);
      float h4756 = lerp_ref(h47, h56, c.oy, oy1);
      o[f] = lerp_ref(h0312, h4756, c.oz, oz1);
    }
    *reinterpret_cast<V*>(out + t * F) = *reinterpret_cast<V*>(o);
  }
}

template <int F>
__device__ __forceinline__ void red_add(float* addr, const float* v);
template <>
__device__ __forceinline__ void red_add<1>(float* addr, const float* v) { atomicAdd(addr, v[0]); }
template <>
__device__ __forceinline__ void red_add<2>(float* addr, const float* v) {
  atomicAdd(reinterpret_cast<float2*>(addr), make_float2(v[0], v[1]));
}
template <>
__device__ __forceinline__ void red_add<4>(float* addr, const float* v) {
  atomicAdd(reinterpret_cast<float4*>(addr), make_float4(v[0], v[1], v[2], v[3]));
}

constexpr int kBwdPts = 128;  // points per block tile in the backward kernel

// Backward: the block stages a [128 points x L*F] tile of d_out (coalesced) and the 128 positions in shared
// memory, then each warp takes (level, 32-point group) units so that the 32 lanes of a warp work on the SAME
// level: lanes that hit the same table entry (frequent on coarse levels, where a cell holds many samples of a
// ray) are merged with match.any + shuffles and issue ONE vector atomic.
template <int F>
__global__ void __launch_bounds__(256)
hashgrid_bwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ d_out,
                    float* __restrict__ d_tables, int64_t P, int L, uint32_t mask) {
  using V = typename Vec<F>::T;
  extern __shared__ float smem[];
  const size_t T = (size_t)mask + 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int LF = L * F;
  const int stride = LF + (F == 1 ? 1 : F);
  float* s_g = smem;                      // [kBwdPts][stride]
  float* s_x = smem + kBwdPts * stride;   // [kBwdPts][3]
  const int cxs[8] = {1, 1, 0, 0, 1, 1, 0, 0};
  const int cys[8] = {1, 0, 0, 1, 1, 0, 0, 1};
  const int czs[8] = {1, 1, 1, 1, 0, 0, 0, 0};
  const int64_t ntiles = (P + kBwdPts - 1) / kBwdPts;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t p0 = tile * kBwdPts;
    const int npts = (int)min((int64_t)kBwdPts, P - p0);
    __syncthreads();
    for (int i = threadIdx.x; i < npts * LF; i += blockDim.x) {
      int pp = i / LF, c = i - pp * LF;
      s_g[pp * stride + c] = d_out[p0 * LF + i];
    }
    for (int i = threadIdx.x; i < npts * 3; i += blockDim.x) s_x[i] = x[p0 * 3 + i];
    __syncthreads();
    const int groups = kBwdPts / 32;
    for (int unit = warp; unit < L * groups; unit += blockDim.x / 32) {
      const int l = unit / groups;
      const int pp = (unit - l * groups) * 32 + lane;
      const bool active = pp < npts;
      const int pq = active ? pp : 0;
      float r = __ldg(res + l);
      Corners c;
      corners_of(__fmul_rn(s_x[pq * 3 + 0], r), __fmul_rn(s_x[pq * 3 + 1], r), __fmul_rn(s_x[pq * 3 + 2], r), mask, c);
      V gv = *reinterpret_cast<const V*>(s_g + pq * stride + l * F);
      const float* g = reinterpret_cast<const float*>(&gv);
      float wx[2] = {__fsub_rn(1.0f, c.ox), c.ox};
      float wy[2] = {__fsub_rn(1.0f, c.oy), c.oy};
      float wz[2] = {__fsub_rn(1.0f, c.oz), c.oz};
      float* tab = d_tables + (size_t)l * T * F;
      // One match on the CELL decides the path for the whole warp: when no two lanes sit in the same cell (random points,
      // fine levels) the per-corner match / shuffle merging below cannot pay and is skipped.
      const unsigned cell_peers = __match_any_sync(0xffffffffu, active ? c.idx[6] : 0xffffff00u + (unsigned)lane);
      const bool solo = __all_sync(0xffffffffu, __popc(cell_peers) == 1);
      if (solo) {
        if constexpr (F == 2) {
          // The two corners of an x edge hash to (x0 ^ A) and (x1 ^ A) (the x prime is 1): whenever x0 is even they are
          // the two halves of one aligned 16-byte slot and take ONE 16-byte vector atomic instead of two 8-byte ones
          const bool pair_ok = ((c.idx[3] ^ c.idx[0]) == 1u);
          constexpr int ka[4] = {0, 1, 5, 4}, kb[4] = {3, 2, 6, 7};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float wa = wx[1] * wy[cys[kb[j]]] * wz[czs[kb[j]]], wb = wx[0] * wy[cys[kb[j]]] * wz[czs[kb[j]]];
            if (!active) continue;
            if (pair_ok) {
              const bool b_hi = (c.idx[kb[j]] & 1u) != 0u;
              const float wlo = b_hi ? wa : wb, whi = b_hi ? wb : wa;
              atomicAdd(reinterpret_cast<float4*>(tab + (size_t)(c.idx[kb[j]] & ~1u) * 2),
                        make_float4(wlo * g[0], wlo * g[1], whi * g[0], whi * g[1]));
            } else {
              atomicAdd(reinterpret_cast<float2*>(tab + (size_t)c.idx[ka[j]] * 2), make_float2(wa * g[0], wa * g[1]));
              atomicAdd(reinterpret_cast<float2*>(tab + (size_t)c.idx[kb[j]] * 2), make_float2(wb * g[0], wb * g[1]));
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (!active) continue;
            const float w = wx[cxs[k]] * wy[cys[k]] * wz[czs[k]];
            float v[F];
#pragma unroll
            for (int f = 0; f < F; ++f) v[f] = w * g[f];
            red_add<F>(tab + (size_t)c.idx[k] * F, v);
          }
        }
        continue;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float w = active ? wx[cxs[k]] * wy[cys[k]] * wz[czs[k]] : 0.0f;
        float v[F];
#pragma unroll
        for (int f = 0; f < F; ++f) v[f] = w * g[f];
        uint32_t key = active ? c.idx[k] : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, key);
        int leader = __ffs(peers) - 1;
        if (__popc(peers) > 1) {
          unsigned rem = peers;
          float acc[F];
#pragma unroll
          for (int f = 0; f < F; ++f) acc[f] = 0.0f;
          while (rem) {
            int src = __ffs(rem) - 1;
            rem &= rem - 1;
#pragma unroll
            for (int f = 0; f < F; ++f) acc[f] += __shfl_sync(peers, v[f], src);
          }
#pragma unroll
          for (int f = 0; f < F; ++f) v[f] = acc[f];
        }
        if (lane == leader && active) red_add<F>(tab + (size_t)c.idx[k] * F, v);
      }
    }
  }
}

}  // namespace

extern "C" int nmx_hashgrid_hash(const int32_t* coords, int32_t* idx, int64_t M, int log2_T, void* stream) {
  NMX_CHECK_ARG(M >= 0 && log2_T >= 1 && log2_T <= 27, "M >= 0, 1 <= log2_T <= 27");
  if (M == 0) return 0;
  hash_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(coords, idx, M, (1u << log2_T) - 1u);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_hashgrid_fwd(const float* x, const float* tables, const float* scaled_res, float* out,
                                int32_t* idx_out, int64_t P, int L, int F, int log2_T, void* stream) {
  NMX_CHECK_ARG(P >= 0 && L >= 1 && L <= 32 && log2_T >= 1 && log2_T <= 27, "P >= 0, 1 <= L <= 32, 1 <= log2_T <= 27");
  NMX_CHECK_ARG(F == 1 || F == 2 || F == 4, "F in {1,2,4}");
  if (P == 0) return 0;
  int64_t total = P * L;
  uint32_t mask = (1u << log2_T) - 1u;
  int blocks = grid_for(total, 256, 16);
  cudaStream_t s = (cudaStream_t)stream;
  if (F == 1) hashgrid_fwd_kernel<1><<<blocks, 256, 0, s>>>(x, tables, scaled_res, out, idx_out, total, L, mask);
  else if (F == 2) hashgrid_fwd_kernel<2><<<blocks, 256, 0, s>>>(x, tables, scaled_res, out, idx_out, total, L, mask);
  else hashgrid_fwd_kernel<4><<<blocks, 256, 0, s>>>(x, tables, scaled_res, out, idx_out, total, L, mask);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_hashgrid_bwd(const float* x, const float* scaled_res, const float* d_out, float* d_tables,
                                int64_t P, int L, int F, int log2_T, void* stream) {
  NMX_CHECK_ARG(P >= 0 && L >= 1 && L <= 32 && log2_T >= 1 && log2_T <= 27, "P >= 0, 1 <= L <= 32, 1 <= log2_T <= 27");
  NMX_CHECK_ARG(F == 1 || F == 2 || F == 4, "F in {1,2,4}");
  if (P == 0) return 0;
  uint32_t mask = (1u << log2_T) - 1u;
  int64_t ntiles = (P + kBwdPts - 1) / kBwdPts;
  int blocks = (int)(ntiles < (int64_t)kNumSMs * 8 ? ntiles : (int64_t)kNumSMs * 8);
  int stride = L * F + (F == 1 ? 1 : F);
  size_t smem = (size_t)(kBwdPts * stride + kBwdPts * 3) * sizeof(float);
  cudaStream_t s = (cudaStream_t)stream;
#define NMX_HG_BWD(FF)                                                                                         \
  do {                                                                                                         \
    if (smem > 48 * 1024)                                                                                      \
      NMX_CUDA(cudaFuncSetAttribute(hashgrid_bwd_kernel<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hashgrid_bwd_kernel<FF><<<blocks, 256, smem, s>>>(x, scaled_res, d_out, d_tables, P, L, mask);             \
  } while (0)
  if (F == 1) NMX_HG_BWD(1);
  else if (F == 2) NMX_HG_BWD(2);
  else NMX_HG_BWD(4);
#undef NMX_HG_BWD
  NMX_LAUNCH_CHECK();
  return 0;
}
