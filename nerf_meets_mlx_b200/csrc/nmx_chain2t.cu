// TRAINING forward of the 8 x 256 view-dir NeRF (models/NeRF.py:201-243) on CTA PAIRS with two tiles in ping-pong:
// nmx_chain2.cu's schedule (tcgen05 cta_group::2, each CTA stages half of every weight slab, the leader's MMA warp
// alternates layer l of tile X / layer l of tile Y so that one tile's epilogue runs under the other tile's MMAs) PLUS
// everything the backward pass needs:
//   * every trunk activation h_0 .. h_7 and the dir-layer activation hd are TMA-stored (the shared-memory tile that is
//     the next layer's A operand doubles as the staging buffer) by the slot's service warp,
//   * the ReLU sign bits of each (1 bit per activation, one contiguous 4 KB tile per 128 rows and layer),
//   * the encoded input tile X0 = [PE(pos) | PE(dir)] for the weight gradients of the first / skip / dir layers.
// Why pairs: in the one-tile training chain (nmx_chain.cu) the store of a 64 KB tile-layer, the epilogue and the next
// layer's MMAs are serialised per CTA, and the kernel runs at ~35 % tensor duty and ~3.2 TB/s of stores -- half of what
// either pipe can do (write-only HBM bandwidth is 6.3 TB/s, profiles/r2_bw_probe.txt).  With two tiles in flight the
// stores of tile X's layer l drain while the tensor pipe runs tile Y's layer l and X's layer l + 1.
//
// Barriers added to nmx_chain2.cu's set (each CTA has its own copy, all local):
//   st_ready[t][c]  8 arrivals: the warps that own chunk c have written slot t's chunk (same event as act[t][c], but also
//                   signalled for the LAST layer, which feeds no MMA) -> the store warp may issue its TMA store
//   store_done[t]   1 arrival by slot t's service warp: the bulk group of slot t's previous saved layer has finished
//                   READING shared memory -> the epilogue may overwrite slot t's chunks and sign-bit staging tile
// Service warp of slot t (warps 16 / 17 of each CTA): issues slot t's stores in layer order and, while it polls for the
// next finished chunk, encodes the NEXT tile's PE(pos) chunk one 32-row group at a time (after the skip layer has released
// the chunk), so neither job delays the other by more than one row group (~2 k clocks).
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_chain.cuh"
#include "nmx_chain_dev.cuh"
#include "nmx_gemm.cuh"
#include <cstring>

using namespace nmx;
using namespace nmx::sm100;
using namespace nmx::chain_dev;

namespace {

constexpr int kNL = 10;          // 8 trunk layers, feature layer, dir layer
constexpr int kSkipL = 5;        // the trunk layer whose input is [PE(pos), h]
constexpr int kFeatL = 8;        // feature layer: no activation, not saved (its weight gradients fold, nmx_mlp.cu)
constexpr int kEpiWarps = 16;
constexpr int kEncWarp0 = 16;    // warps 16, 17: service warps of slot 0 / slot 1: input encoder AND store issuer
                                 // (a 21st warp would put 6 warps on one scheduler: 80 instead of 96 registers per thread)
constexpr int kTmaWarp = 18;
constexpr int kMmaWarp = 19;
constexpr int kThreads = 20 * 32;
constexpr int kStages = 3;
constexpr int kChunk = 128 * 64 * 2;   // one 128-row x 64-col bf16 chunk
constexpr int kHalfSlab = 128 * 64 * 2;
constexpr int kBitsTile = 128 * 32;    // sign bits of a 128-row x 256-col tile

struct SmemT {
  static constexpr int kActOff = 0;                               // [2 slots][5 chunks]: 4 activation chunks + PE(pos)
  static constexpr int kRingOff = kActOff + 2 * 5 * kChunk;
  static constexpr int kBiasOff = kRingOff + kStages * kHalfSlab; // [kNL][256] fp32
  static constexpr int kBitsOff = kBiasOff + kNL * 256 * 4;       // [2 slots] sign-bit staging tiles
  static constexpr int kBarOff = kBitsOff + 2 * kBitsTile;
  static constexpr int kNumBars = 2 * kStages + 2 + 2 + 8 + 2 + 2 + 2 + 8 + 8 + 2;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kAlloc = kTmemPtrOff + 16;                 // no alignment slack: the base must be 1024 B aligned
  static_assert(kAlloc <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

struct Params {
  int P;
  const float* params;
  float* out;                // [P, 4] fp32 (rgb, sigma)
  const float* rays; int ray_stride; const float* z;
  long long p0, b0; int n_per_ray;
  const float* dir_bias;     // [rays, 128] fp32: W_dir[:, W:] PE(dir) per ray
  const float* consts;       // bias [kNL][256], w_alpha [256], w_rgb [3][128] (16 B aligned)
  uint32_t* bits;            // sign-bit store [(D + 1) * cap rows][8] uint32
  int cap;                   // rows per slot of the activation / sign-bit stores
  int alpha_b_off, rgb_b_off;
  int dbg;  // NMX_EXPERIMENTS builds only (NMX_CHAIN2T_DBG): 1 no activation / sign-bit stores, 2 no X0 stores, 4 no store_done wait
};
struct Maps {
  CUtensorMap w[kNL];
  CUtensorMap save;          // [(D + 1) * cap, 256] activation store, box 128 x 64
  CUtensorMap hd;            // [P, 128], box 128 x 64
  CUtensorMap x0;            // [P, 128], box 128 x 64
};

__device__ __forceinline__ int layer_slabs(int l) { return l == 0 ? 1 : (l == kSkipL ? 5 : 4); }
__device__ __forceinline__ int slab_src(int l, int s) { return l == 0 ? 4 : (l == kSkipL ? (s == 0 ? 4 : s - 1) : s); }
__device__ __forceinline__ int layer_N(int l) { return l == kNL - 1 ? 128 : 256; }

// epi_cols with the head weights read from GLOBAL memory (L1-resident 2.5 KB): the 8 KB of sign-bit staging leave no
// room for them in shared memory.  HEAD 1: alpha (w [256]); HEAD 2: rgb (w [3][128]).
template <int HEAD>
__device__ __forceinline__ void head_partial(const uint32_t (&pk)[4], const float* __restrict__ w, const int col, float& a0,
                                             float (&rgbp)[3]) {
  float xr[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    xr[2 * e] = __uint_as_float(pk[e] << 16);
    xr[2 * e + 1] = __uint_as_float(pk[e] & 0xffff0000u);
  }
  if (HEAD == 1) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + col));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + col + 4));
    a0 += xr[0] * w0.x + xr[1] * w0.y + xr[2] * w0.z + xr[3] * w0.w + xr[4] * w1.x + xr[5] * w1.y + xr[6] * w1.z + xr[7] * w1.w;
  } else {
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + o * 128 + col));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + o * 128 + col + 4));
      rgbp[o] += xr[0] * w0.x + xr[1] * w0.y + xr[2] * w0.z + xr[3] * w0.w + xr[4] * w1.x + xr[5] * w1.y + xr[6] * w1.z +
                 xr[7] * w1.w;
    }
  }
}

// 32 columns [c0, c0 + 32) of chunk c: bf16(act(acc + bias)) -> swizzled smem chunk, sign bits -> staging word,
// optional head partial sums.  Same arithmetic as chain_dev::epi_cols.
template <bool RELU, int HEAD, bool BITS>
__device__ __forceinline__ void epi32(const uint32_t (&r)[32], const int c, const int c0, const int piece0,
                                      const uint32_t bias_addr, const uint32_t act_row_addr, const uint32_t swz,
                                      const float* __restrict__ hw, float& a0, float (&rgbp)[3], const uint32_t bits_addr) {
  const uint32_t so = act_row_addr + (uint32_t)c * kChunk;
  uint32_t bits = 0;
#pragma unroll
  for (int p4 = 0; p4 < 4; ++p4) {
    const float4 b0 = lds128(bias_addr + (uint32_t)(c0 + p4 * 8) * 4u);
    const float4 b1 = lds128(bias_addr + (uint32_t)(c0 + p4 * 8 + 4) * 4u);
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint64_t x = add_f32x2(pack64(r[p4 * 8 + 2 * e], r[p4 * 8 + 2 * e + 1]),
                                   pack64(__float_as_uint(bv[2 * e]), __float_as_uint(bv[2 * e + 1])));
      pk[e] = cvt_bf16x2<RELU>(x);
      if (RELU && BITS) bits |= nz_mask_bf16x2(pk[e]) & (0x00010001u << (p4 * 4 + e));
    }
    sts128(so + ((((uint32_t)(piece0 + p4)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
    if (HEAD != 0) head_partial<HEAD>(pk, hw, c0 + p4 * 8, a0, rgbp);
  }
  if (RELU && BITS) asm volatile("st.shared.u32 [%0], %1;" ::"r"(bits_addr), "r"(bits) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_chain2_train_kernel(const __grid_constant__ Maps maps, const Params prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SW128 operand tiles need 1024 B alignment; the layout has no slack
  float* s_bias = reinterpret_cast<float*>(smem + SmemT::kBiasOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SmemT::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;    // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint64_t* act_ready = tempty + 2;     // [2][4]
  uint64_t* pos_full = act_ready + 8;   // [2]
  uint64_t* pos_empty = pos_full + 2;   // [2]
  uint64_t* tempty_peer = pos_empty + 2;    // [2]
  uint64_t* act_peer = tempty_peer + 2;     // [2][4]
  uint64_t* st_ready = act_peer + 8;        // [2][4]
  uint64_t* store_done = st_ready + 8;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SmemT::kTmemPtrOff);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int num_pt = (prm.P + 255) / 256;  // pair tiles of 256 points

  if (warp == kTmaWarp && lane == 0) {
    for (int l = 0; l < kNL; ++l) tma_prefetch_desc(&maps.w[l]);
    tma_prefetch_desc(&maps.save);
    tma_prefetch_desc(&maps.hd);
    tma_prefetch_desc(&maps.x0);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&tfull[t], 1);
      mbar_init(&tempty[t], kEpiWarps);
      mbar_init(&tempty_peer[t], 1);
      for (int c = 0; c < 4; ++c) {
        mbar_init(&act_ready[t * 4 + c], kEpiWarps / 2);
        mbar_init(&act_peer[t * 4 + c], 1);
        mbar_init(&st_ready[t * 4 + c], kEpiWarps / 2);
      }
      mbar_init(&pos_full[t], 2);
      mbar_init(&pos_empty[t], 1);
      mbar_init(&store_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == kEncWarp0) tmem_alloc_pair<512>(tmem_ptr);
  for (int i = threadIdx.x; i < kNL * 256; i += kThreads) s_bias[i] = __ldg(prm.consts + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t act_base = smem_u32(smem + SmemT::kActOff);

  if (warp == kTmaWarp) {
    // ====================================================== weight producer: this CTA's half of every slab
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int l = 0; l < kNL; ++l) {
          const int half_rows = layer_N(l) / 2;
          const uint32_t bytes = (uint32_t)half_rows * 128u;
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            for (int s = 0; s < layer_slabs(l); ++s) {
              mbar_wait(&empty[stage], phase ^ 1);
              if (leader_cta) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
              tma_load_2d_pair(smem + SmemT::kRingOff + stage * kHalfSlab, &maps.w[l], &full[stage], s * 64,
                               (int)rank * half_rows);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ====================================================== MMA issuer (leader CTA only); CTA 1: barrier relay
    if (leader_cta) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t lc[2] = {0, 0};    // layers issued so far per slot (accumulator generations)
      uint32_t gen[2] = {0, 0};   // act_ready generations produced so far per slot
      uint32_t tc[2] = {0, 0};    // tiles started per slot
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t ring16 = smem_u32(smem + SmemT::kRingOff) >> 4;
      constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;  // SBO 1024 B, version 1, SWIZZLE_128B
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int l = 0; l < kNL; ++l) {
          const uint32_t idesc = make_idesc_bf16(256, layer_N(l), 0, 0);
          const int ns = layer_slabs(l);
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            mbar_wait(&tempty[t], (lc[t] & 1u) ^ 1u);
            mbar_wait_cluster(&tempty_peer[t], (lc[t] & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            for (int s = 0; s < ns; ++s) {
              const int src = slab_src(l, s);
              mbar_wait(&full[stage], phase);
              if (src == 4) mbar_wait_cluster(&pos_full[t], tc[t] & 1u);
              else {
                mbar_wait(&act_ready[t * 4 + src], (gen[t] - 1u) & 1u);
                mbar_wait_cluster(&act_peer[t * 4 + src], (gen[t] - 1u) & 1u);
              }
              tc_fence_after();
              const uint32_t a16 = smem16 + (uint32_t)((t * 5 + src) * (kChunk >> 4));
              const uint64_t adesc = kDescHi | (uint64_t)((a16 & 0x3fffu) | 0x10000u);
              const uint64_t bdesc = kDescHi | (uint64_t)(((ring16 + (uint32_t)stage * (kHalfSlab >> 4)) & 0x3fffu) | 0x10000u);
              if (elect_one()) {
                umma_bf16_pair(d_tmem, adesc, bdesc, idesc, s > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                umma_commit_pair(&empty[stage]);
                if (s == ns - 1) {
                  umma_commit_pair(&tfull[t]);
                  if (l == kSkipL) umma_commit_pair(&pos_empty[t]);
                }
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            ++lc[t];
            if (l < kNL - 1) ++gen[t];
            if (l == kNL - 1) ++tc[t];
          }
        }
      }
    } else {
      const uint32_t tempty_peer_leader = mapa_u32(smem_u32(&tempty_peer[0]), 0);
      const uint32_t act_peer_leader = mapa_u32(smem_u32(&act_peer[0]), 0);
      uint32_t lc[2] = {0, 0}, gen[2] = {0, 0};
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int l = 0; l < kNL; ++l) {
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            if (l < kNL - 1) {
              for (int c = 0; c < 4; ++c) {
                mbar_wait(&act_ready[t * 4 + c], gen[t] & 1u);
                if (lane == 0) mbar_arrive_cluster_relaxed(act_peer_leader + (uint32_t)(t * 4 + c) * 8u);
                __syncwarp();
              }
              ++gen[t];
            }
            mbar_wait(&tempty[t], lc[t] & 1u);
            if (lane == 0) mbar_arrive_cluster_relaxed(tempty_peer_leader + (uint32_t)t * 8u);
            __syncwarp();
            ++lc[t];
          }
        }
      }
    }
  } else if (warp >= kEncWarp0 && warp < kEncWarp0 + 2) {
    // ====================================================== service warp of slot t: input encoder + store issuer
    const int t = warp - kEncWarp0;
    const uint32_t pos_addr = act_base + (uint32_t)((t * 5 + 4) * kChunk);
    const uint32_t posf_leader = mapa_u32(smem_u32(&pos_full[t]), 0);
    // one 32-row group of a tile's input encoding: PE(pos) row -> chunk 4
    auto encode_group = [&](long long row0, int rr) {
      const int row_local = rr * 32 + lane;
      const long long row = row0 + row_local;
      encode_pos_row(prm.rays, prm.ray_stride, prm.z, prm.p0 + row, prm.n_per_ray, row < prm.P,
                     pos_addr + (uint32_t)row_local * 128u, (uint32_t)(row_local & 7));
    };
    // publish an encoded tile: fence, signal the leader's MMA warp, TMA-store the PE(pos) chunk to X0[:, 0:64]
    auto publish = [&](long long row0) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(posf_leader);
        if (row0 < prm.P && !NMX_DBG(prm, 2)) {
          tma_store_2d(&maps.x0, smem + SmemT::kActOff + (t * 5 + 4) * kChunk, 0, (int)row0);
          tma_store_commit();
        }
      }
      __syncwarp();
    };
    auto poll = [&](uint64_t* bar, uint32_t parity) -> bool {  // non-blocking, warp-uniform
      uint32_t ok = 0;
      if (lane == 0) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
      }
      return __shfl_sync(0xffffffffu, ok, 0) != 0;
    };
    uint32_t ph01 = 0, ph23 = 0;  // st_ready phases consumed (chunks 0,1 / chunks 2,3)
    uint32_t tcnt = 0;            // tiles of this slot so far
    {  // the slot's first tile: nothing to overlap with yet
      const int tile = cluster_id + t * num_clusters;
      if (tile < num_pt) {
        const long long row0 = (long long)tile * 256 + (long long)rank * 128;
#pragma unroll 1
        for (int rr = 0; rr < 4; ++rr) encode_group(row0, rr);
        publish(row0);
      }
    }
    for (int j = 0;; ++j, ++tcnt) {
      const int tile = cluster_id + (2 * j + t) * num_clusters;
      if (tile >= num_pt) break;
      const int row0 = tile * 256 + (int)rank * 128;
      const bool live = row0 < prm.P;  // a half tile entirely beyond P must not touch the stores
      const int ntile = cluster_id + (2 * (j + 1) + t) * num_clusters;
      const long long nrow0 = (long long)ntile * 256 + (long long)rank * 128;
      int enc_rr = ntile < num_pt ? 0 : 4;  // row groups of the next tile still to encode
      bool chunk_free = false;              // the skip layer of THIS tile has released the PE(pos) chunk
      for (int l = 0; l < kNL; ++l) {
        if (l == kFeatL) continue;
        const int nck = layer_N(l) / 64;
        for (int c = 0; c < nck; ++c) {
          const uint32_t par = (c < 2 ? ph01 : ph23) & 1u;
          while (!poll(&st_ready[t * 4 + c], par)) {
            bool worked = false;
            if (enc_rr < 4 && l > kSkipL) {
              if (!chunk_free) chunk_free = poll(&pos_empty[t], tcnt & 1u);
              if (chunk_free) { encode_group(nrow0, enc_rr++); worked = true; }
            }
            if (!worked) __nanosleep(64);  // nothing to do: do not burn issue slots (and power: the step runs at the cap)
          }
          if (live && lane == 0 && !NMX_DBG(prm, 1)) {
            const uint8_t* src = smem + SmemT::kActOff + (t * 5 + c) * kChunk;
            if (l < kFeatL) tma_store_2d(&maps.save, src, c * 64, l * prm.cap + row0);
            else tma_store_2d(&maps.hd, src, c * 64, row0);
          }
          __syncwarp();
        }
        ++ph01;
        if (nck == 4) ++ph23;
        if (lane == 0) {
          if (live && !NMX_DBG(prm, 1)) {
            const int slot = l < kFeatL ? l : kFeatL;  // sign-bit slots 0..7 = h_l, slot 8 = hd
            bulk_store_1d(prm.bits + ((size_t)slot * prm.cap + (size_t)row0) * 8, smem + SmemT::kBitsOff + t * kBitsTile,
                          kBitsTile);
          }
          tma_store_commit();
          tma_store_wait_read<0>();  // (also covers the X0 store of this tile's PE(pos) chunk)
          mbar_arrive(&store_done[t]);
        }
        __syncwarp();
      }
      if (ntile < num_pt) {  // whatever is left of the next tile's encoding, then hand it to the MMA warp
        if (!chunk_free) mbar_wait(&pos_empty[t], tcnt & 1u);
#pragma unroll 1
        for (; enc_rr < 4; ++enc_rr) encode_group(nrow0, enc_rr);
        publish(nrow0);
      }
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  } else {
    // ====================================================== epilogue (warps 0..15 of both CTAs)
    const int q = warp & 3;       // TMEM lane quadrant
    const int part = warp >> 2;   // which 32 columns of a half step
    const int row_local = q * 32 + lane;
    const uint32_t swz = (uint32_t)(row_local & 7);
    const uint32_t bias_base = smem_u32(s_bias);
    const float* w_alpha = prm.consts + kNL * 256;
    const float* w_rgb = w_alpha + 256;
    float hb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (part == 0) {
      hb[0] = __ldg(prm.params + prm.rgb_b_off + 0);
      hb[1] = __ldg(prm.params + prm.rgb_b_off + 1);
      hb[2] = __ldg(prm.params + prm.rgb_b_off + 2);
      hb[3] = __ldg(prm.params + prm.alpha_b_off);
    }
    uint32_t lc[2] = {0, 0};
    uint32_t sd_groups[2] = {0, 0}, sd_waited[2] = {0, 0};  // saved layers produced / store_done phases observed per slot
    float alpha_p[2] = {0.0f, 0.0f};
    for (int j = 0;; ++j) {
      const int tx = cluster_id + (2 * j) * num_clusters;
      if (tx >= num_pt) break;
      const bool valid_y = tx + num_clusters < num_pt;
      for (int l = 0; l < kNL; ++l) {
        const int nsteps = layer_N(l) / 64;
        const uint32_t bias_addr = bias_base + (uint32_t)l * 1024u;
        for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
          const int tile = tx + t * num_clusters;
          const long long row = (long long)tile * 256 + (long long)rank * 128 + row_local;
          const uint32_t act_row_addr = act_base + (uint32_t)(t * 5 * kChunk) + (uint32_t)row_local * 128u;
          // sign-bit staging tile, WORD-major [8 words][128 rows]: a warp's 32 rows write 32 consecutive words (the
          // row-major layout of the one-tile chain is an 8-way bank conflict per store)
          const uint32_t bits_row = smem_u32(smem + SmemT::kBitsOff + t * kBitsTile) + (uint32_t)row_local * 4u;
          const uint32_t tacc = tmem_base + (uint32_t)t * 256u + ((uint32_t)(q * 32) << 16);
          float a0 = 0.0f;
          float rgbp[3] = {0.0f, 0.0f, 0.0f};
          const float* dbias = nullptr;
          if (l == kNL - 1) {
            const long long rr = row < prm.P ? row : (long long)prm.P - 1;
            dbias = prm.dir_bias + ((prm.p0 + rr) / prm.n_per_ray - prm.b0) * 128;
          }
          mbar_wait(&tfull[t], lc[t] & 1u);
          tc_fence_after();
          // the store of this slot's previous saved layer must have finished reading the chunks / the sign-bit tile
          if (sd_waited[t] < sd_groups[t] && !NMX_DBG(prm, 4)) {
            mbar_wait(&store_done[t], (sd_groups[t] - 1u) & 1u);
            sd_waited[t] = sd_groups[t];
          }
          const int nh = nsteps / 2;
          const int sub = part & 1;
          uint32_t r32[32];
          tmem_ld_32x32(tacc + (uint32_t)((part >> 1) * 64 + sub * 32), r32);
#pragma unroll 1
          for (int h = 0; h < nh; ++h) {
            const int c = 2 * h + (part >> 1);
            const int c0 = c * 64 + sub * 32;
            const uint32_t bits_addr = bits_row + (uint32_t)(4 * h + part) * 512u;
            tmem_ld_wait_regs<32>(r32);
            if (h == nh - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[t]);
            }
            if (l == kNL - 1) {  // per-ray view-dir term of the dir layer
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(dbias + c0) + i4);
                r32[i4 * 4 + 0] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 0]) + d.x);
                r32[i4 * 4 + 1] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 1]) + d.y);
                r32[i4 * 4 + 2] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 2]) + d.z);
                r32[i4 * 4 + 3] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 3]) + d.w);
              }
              epi32<true, 2, true>(r32, c, c0, 4 * sub, bias_addr, act_row_addr, swz, w_rgb, a0, rgbp, bits_addr);
            } else if (l == 7) {
              epi32<true, 1, true>(r32, c, c0, 4 * sub, bias_addr, act_row_addr, swz, w_alpha, a0, rgbp, bits_addr);
            } else if (l == kFeatL) {
              epi32<false, 0, false>(r32, c, c0, 4 * sub, bias_addr, act_row_addr, swz, nullptr, a0, rgbp, bits_addr);
            } else {
              epi32<true, 0, true>(r32, c, c0, 4 * sub, bias_addr, act_row_addr, swz, nullptr, a0, rgbp, bits_addr);
            }
            if (h + 1 < nh) tmem_ld_32x32(tacc + (uint32_t)(c0 + 128), r32);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (l < kNL - 1) mbar_arrive(&act_ready[t * 4 + c]);
              if (l != kFeatL) mbar_arrive(&st_ready[t * 4 + c]);
            }
          }
          if (l != kFeatL) ++sd_groups[t];
          if (l == 7) alpha_p[t] = a0;
          if (l == kNL - 1) {
            // heads: combine the four column parts' partial sums through shared memory (chunk 3 of this slot: the dir
            // layer neither reads nor writes it, and its last store -- layer 7's -- was waited for above)
            float* xchg = reinterpret_cast<float*>(smem + SmemT::kActOff + (t * 5 + 3) * kChunk);
            float4 v4 = make_float4(rgbp[0], rgbp[1], rgbp[2], alpha_p[t]);
            if (part > 0) *reinterpret_cast<float4*>(xchg + ((part - 1) * 128 + row_local) * 4) = v4;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            if (part == 0) {
#pragma unroll
              for (int pp = 0; pp < 3; ++pp) {
                const float4 o = *reinterpret_cast<const float4*>(xchg + (pp * 128 + row_local) * 4);
                v4.x += o.x; v4.y += o.y; v4.z += o.z; v4.w += o.w;
              }
              if (row < prm.P)
                *reinterpret_cast<float4*>(prm.out + (size_t)row * 4) =
                    make_float4(v4.x + hb[0], v4.y + hb[1], v4.z + hb[2], v4.w + hb[3]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          }
          ++lc[t];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kEncWarp0) tmem_dealloc_pair<512>(tmem_base);
}

// ================================================================================================ backward
// Data-gradient chain of the same net on CTA pairs, two tiles in ping-pong (the pair version of nmx_chain.cu MODE 1):
//   step A : d_hd = (d_rgb W_rgb) * [hd > 0] on the CUDA cores -> chunks 0,1 (A operand of layer 0) and the d_hd store
//   layer 0: d_feature = d_hd W_dir[:, :W]                                   (K = 128, not stored)
//   layer 1: dY_7 = (d_feature W_feat + d_sigma (x) w_alpha) * [h_7 > 0]     -> slot 7 of the dY store
//   layer b: dY_{8-b} = (dY_{9-b} W_{9-b}[:, h part]) * [h_{8-b} > 0], b = 2..8 -> slots 6..0
// Every dY is TMA-stored for the weight-gradient kernels; the ReLU masks come from the forward's sign bits (one 4 KB
// tile per step, bulk-loaded into a double-buffered staging area by the slot's service warp, which also issues the stores).
constexpr int kNB = 9;

struct SmemB {
  static constexpr int kActOff = 0;                               // [2 slots][4 chunks]
  static constexpr int kRingOff = kActOff + 2 * 4 * kChunk;
  static constexpr int kBitsOff = kRingOff + kStages * kHalfSlab; // [2 slots][2 buffers] sign-bit tiles
  static constexpr int kWaOff = kBitsOff + 4 * kBitsTile;         // w_alpha [256] fp32
  static constexpr int kWrgbOff = kWaOff + 256 * 4;               // w_rgb [3][128] fp32
  static constexpr int kBarOff = kWrgbOff + 3 * 128 * 4;
  static constexpr int kNumBars = 2 * kStages + 2 + 2 + 8 + 2 + 8 + 8 + 2 + 4 + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kAlloc = kTmemPtrOff + 16;
  static_assert(kAlloc <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

struct ParamsB {
  int P;
  const float* params;
  const float* d_out;        // [P, 4] fp32 (d_rgb, d_sigma)
  const uint32_t* bits;      // sign-bit store [(D + 1) * cap rows][8]
  int cap;
  int alpha_w_off, rgb_w_off;
  int bits_wm;               // sign-bit tiles are word-major [8][128] (pair forward) instead of row-major [128][8]
};
struct MapsB {
  CUtensorMap w[kNB];        // transposed bf16 weights [256, K_b], box 128 x 64
  CUtensorMap save;          // dY store [(D + 1) * cap, 256], box 128 x 64
  CUtensorMap hd;            // d_hd [P, 128], box 128 x 64
};

__device__ __forceinline__ bool bar_poll(uint64_t* bar, uint32_t parity, int lane) {  // non-blocking, warp-uniform
  uint32_t ok = 0;
  if (lane == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_chain2_bwd_kernel(const __grid_constant__ MapsB maps, const ParamsB prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  float* s_wa = reinterpret_cast<float*>(smem + SmemB::kWaOff);
  float* s_wrgb = reinterpret_cast<float*>(smem + SmemB::kWrgbOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SmemB::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;        // [2]
  uint64_t* tempty = tfull + 2;             // [2]
  uint64_t* act_ready = tempty + 2;         // [2][4]
  uint64_t* tempty_peer = act_ready + 8;    // [2]
  uint64_t* act_peer = tempty_peer + 2;     // [2][4]
  uint64_t* st_ready = act_peer + 8;        // [2][4]
  uint64_t* store_done = st_ready + 8;      // [2]
  uint64_t* bits_full = store_done + 2;     // [2 slots][2 buffers]
  uint64_t* bits_empty = bits_full + 4;     // [2 slots][2 buffers]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SmemB::kTmemPtrOff);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int num_pt = (prm.P + 255) / 256;

  if (warp == kTmaWarp && lane == 0) {
    for (int b = 0; b < kNB; ++b) tma_prefetch_desc(&maps.w[b]);
    tma_prefetch_desc(&maps.save);
    tma_prefetch_desc(&maps.hd);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&tfull[t], 1);
      mbar_init(&tempty[t], kEpiWarps);
      mbar_init(&tempty_peer[t], 1);
      for (int c = 0; c < 4; ++c) {
        mbar_init(&act_ready[t * 4 + c], kEpiWarps / 2);
        mbar_init(&act_peer[t * 4 + c], 1);
        mbar_init(&st_ready[t * 4 + c], kEpiWarps / 2);
      }
      mbar_init(&store_done[t], 1);
      for (int k = 0; k < 2; ++k) {
        mbar_init(&bits_full[t * 2 + k], 1);
        mbar_init(&bits_empty[t * 2 + k], kEpiWarps);
      }
    }
    fence_barrier_init();
  }
  if (warp == kEncWarp0) tmem_alloc_pair<512>(tmem_ptr);
  for (int i = threadIdx.x; i < 256; i += kThreads) s_wa[i] = __ldg(prm.params + prm.alpha_w_off + i);
  for (int i = threadIdx.x; i < 384; i += kThreads) s_wrgb[i] = __ldg(prm.params + prm.rgb_w_off + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t act_base = smem_u32(smem + SmemB::kActOff);

  if (warp == kTmaWarp) {
    // ====================================================== weight producer: this CTA's half of every slab
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int b = 0; b < kNB; ++b) {
          const int ns = b == 0 ? 2 : 4;
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            for (int s = 0; s < ns; ++s) {
              mbar_wait(&empty[stage], phase ^ 1);
              if (leader_cta) mbar_arrive_expect_tx(&full[stage], 2 * kHalfSlab);
              tma_load_2d_pair(smem + SmemB::kRingOff + stage * kHalfSlab, &maps.w[b], &full[stage], s * 64, (int)rank * 128);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (leader_cta) {
      // ==================================================== MMA issuer
      int stage = 0;
      uint32_t phase = 0;
      uint32_t lc[2] = {0, 0};     // accumulator generations per slot
      uint32_t g01[2] = {0, 0};    // act_ready generations CONSUMED so far, chunks 0,1 (step A + layers 0..7 per tile)
      uint32_t g23[2] = {0, 0};    // ... chunks 2,3 (layers 0..7 per tile)
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t ring16 = smem_u32(smem + SmemB::kRingOff) >> 4;
      constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;
      const uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int b = 0; b < kNB; ++b) {
          const int ns = b == 0 ? 2 : 4;
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            mbar_wait(&tempty[t], (lc[t] & 1u) ^ 1u);
            mbar_wait_cluster(&tempty_peer[t], (lc[t] & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            for (int s = 0; s < ns; ++s) {
              mbar_wait(&full[stage], phase);
              const uint32_t par = (s < 2 ? g01[t] : g23[t]) & 1u;
              mbar_wait(&act_ready[t * 4 + s], par);
              mbar_wait_cluster(&act_peer[t * 4 + s], par);
              tc_fence_after();
              const uint32_t a16 = smem16 + (uint32_t)((t * 4 + s) * (kChunk >> 4));
              const uint64_t adesc = kDescHi | (uint64_t)((a16 & 0x3fffu) | 0x10000u);
              const uint64_t bdesc = kDescHi | (uint64_t)(((ring16 + (uint32_t)stage * (kHalfSlab >> 4)) & 0x3fffu) | 0x10000u);
              if (elect_one()) {
                umma_bf16_pair(d_tmem, adesc, bdesc, idesc, s > 0 ? 1u : 0u);
#pragma unroll
                for (int k = 1; k < 4; ++k) umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
                umma_commit_pair(&empty[stage]);
                if (s == ns - 1) umma_commit_pair(&tfull[t]);
              }
              __syncwarp();
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            ++lc[t];
            ++g01[t];
            if (b > 0) ++g23[t];
          }
        }
      }
    } else {
      // ==================================================== CTA 1: relay of the local epilogue barriers to the leader
      const uint32_t tempty_peer_leader = mapa_u32(smem_u32(&tempty_peer[0]), 0);
      const uint32_t act_peer_leader = mapa_u32(smem_u32(&act_peer[0]), 0);
      uint32_t lc[2] = {0, 0}, g01[2] = {0, 0}, g23[2] = {0, 0};
      auto relay_act = [&](int t, int c, uint32_t par) {
        mbar_wait(&act_ready[t * 4 + c], par);
        if (lane == 0) mbar_arrive_cluster_relaxed(act_peer_leader + (uint32_t)(t * 4 + c) * 8u);
        __syncwarp();
      };
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int t = 0; t < (valid_y ? 2 : 1); ++t) {  // step A
          relay_act(t, 0, g01[t] & 1u);
          relay_act(t, 1, g01[t] & 1u);
          ++g01[t];
        }
        for (int b = 0; b < kNB; ++b) {
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            if (b < kNB - 1) {
              relay_act(t, 0, g01[t] & 1u);
              relay_act(t, 1, g01[t] & 1u);
              relay_act(t, 2, g23[t] & 1u);
              relay_act(t, 3, g23[t] & 1u);
              ++g01[t];
              ++g23[t];
            }
            mbar_wait(&tempty[t], lc[t] & 1u);
            if (lane == 0) mbar_arrive_cluster_relaxed(tempty_peer_leader + (uint32_t)t * 8u);
            __syncwarp();
            ++lc[t];
          }
        }
      }
    }
  } else if (warp >= kEncWarp0 && warp < kEncWarp0 + 2) {
    // ====================================================== service warp of slot t: sign-bit loader + store issuer
    // Per tile: masked steps m = 0..8 (step A: hd bits, layer 1: h_7, layers 2..8: h_6..h_0) and store groups g = 0..8
    // (step A: d_hd, layers 1..8: dY_7..dY_0).  Both queues are polled without blocking.
    const int t = warp - kEncWarp0;
    uint32_t m_total = 0;          // masked steps loaded so far (buffer = m_total & 1)
    uint32_t ph01 = 0, ph23 = 0;   // st_ready phases consumed
    for (int j = 0;; ++j) {
      const int tile = cluster_id + (2 * j + t) * num_clusters;
      if (tile >= num_pt) break;
      const int row0 = tile * 256 + (int)rank * 128;
      const bool live = row0 < prm.P;
      int m = 0;    // next masked step of this tile to load
      int g = 0;    // next store group of this tile
      int c = 0;    // next chunk of store group g
      while (m < 9 || g < 9) {
        bool progressed = false;
        if (m < 9) {
          const uint32_t buf = m_total & 1u;
          // the buffer's previous user (two masked steps ago) must have been read by all sixteen epilogue warps
          if (m_total < 2 || bar_poll(&bits_empty[t * 2 + buf], ((m_total >> 1) - 1u) & 1u, lane)) {
            if (lane == 0) {
              if (live) {
                const int slot = 8 - m;  // m = 0: hd bits (slot 8); m = 1..8: h_7 .. h_0
                mbar_arrive_expect_tx(&bits_full[t * 2 + buf], kBitsTile);
                bulk_load_1d(smem + SmemB::kBitsOff + (t * 2 + buf) * kBitsTile,
                             prm.bits + ((size_t)slot * prm.cap + (size_t)row0) * 8, kBitsTile, &bits_full[t * 2 + buf]);
              } else {
                mbar_arrive(&bits_full[t * 2 + buf]);
              }
            }
            __syncwarp();
            ++m;
            ++m_total;
            progressed = true;
          }
        }
        if (g < 9) {
          const int nck = g == 0 ? 2 : 4;
          if (bar_poll(&st_ready[t * 4 + c], (c < 2 ? ph01 : ph23) & 1u, lane)) {
            progressed = true;
            if (live && lane == 0) {
              const uint8_t* src = smem + SmemB::kActOff + (t * 4 + c) * kChunk;
              if (g == 0) tma_store_2d(&maps.hd, src, c * 64, row0);
              else tma_store_2d(&maps.save, src, c * 64, (8 - g) * prm.cap + row0);
            }
            __syncwarp();
            if (++c == nck) {
              if (lane == 0) {
                tma_store_commit();
                tma_store_wait_read<0>();
                mbar_arrive(&store_done[t]);
              }
              __syncwarp();
              ++ph01;
              if (nck == 4) ++ph23;
              c = 0;
              ++g;
            }
          }
        }
        if (!progressed) __nanosleep(64);
      }
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  } else {
    // ====================================================== epilogue (warps 0..15 of both CTAs)
    const int q = warp & 3;
    const int part = warp >> 2;
    const int row_local = q * 32 + lane;
    const uint32_t swz = (uint32_t)(row_local & 7);
    const uint32_t wa_addr = smem_u32(s_wa), wrgb_addr = smem_u32(s_wrgb);
    const int sub = part & 1;
    uint32_t lc[2] = {0, 0};
    uint32_t sd_groups[2] = {0, 0}, sd_waited[2] = {0, 0};
    uint32_t mstep[2] = {0, 0};  // masked steps so far per slot (buffer = mstep & 1, parity = (mstep >> 1) & 1)
    float ds[2] = {0.0f, 0.0f};  // d_sigma of this thread's row in the two slots (alpha rank-1 term of layer 1)
    auto wait_store = [&](int t) {
      if (sd_waited[t] < sd_groups[t]) {
        mbar_wait(&store_done[t], (sd_groups[t] - 1u) & 1u);
        sd_waited[t] = sd_groups[t];
      }
    };
    auto bits_word = [&](int t, int w) -> uint32_t {  // this row's sign-bit word w of slot t's current masked step
      const uint32_t buf = mstep[t] & 1u;
      mbar_wait(&bits_full[t * 2 + buf], (mstep[t] >> 1) & 1u);
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v)
                   : "r"(smem_u32(smem + SmemB::kBitsOff + (t * 2 + buf) * kBitsTile) +
                         (prm.bits_wm ? (uint32_t)w * 512u + (uint32_t)row_local * 4u : (uint32_t)row_local * 32u + (uint32_t)w * 4u))
                   : "memory");
      return v;
    };
    auto bits_release = [&](int t) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&bits_empty[t * 2 + (mstep[t] & 1u)]);
      ++mstep[t];
    };
    for (int j = 0;; ++j) {
      const int tx = cluster_id + (2 * j) * num_clusters;
      if (tx >= num_pt) break;
      const bool valid_y = tx + num_clusters < num_pt;
      // ---- step A of both slots: d_hd = (d_rgb W_rgb) * [hd > 0] -> chunks 0,1 (32 columns per warp)
      for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
        const int tile = tx + t * num_clusters;
        const long long row = (long long)tile * 256 + (long long)rank * 128 + row_local;
        float4 dr = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (row < prm.P) dr = __ldg(reinterpret_cast<const float4*>(prm.d_out) + row);
        ds[t] = dr.w;
        const uint32_t act_row_addr = act_base + (uint32_t)(t * 4 * kChunk) + (uint32_t)row_local * 128u;
        wait_store(t);
        const uint32_t bw = bits_word(t, part);
        bits_release(t);
        const int c = part >> 1;
        const uint32_t so = act_row_addr + (uint32_t)c * kChunk;
#pragma unroll
        for (int p4 = 0; p4 < 4; ++p4) {
          const int jj = part * 32 + p4 * 8;
          float v[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const float4 w0 = lds128(wrgb_addr + (uint32_t)(0 * 128 + jj + hh * 4) * 4u);
            const float4 w1 = lds128(wrgb_addr + (uint32_t)(1 * 128 + jj + hh * 4) * 4u);
            const float4 w2 = lds128(wrgb_addr + (uint32_t)(2 * 128 + jj + hh * 4) * 4u);
            v[hh * 4 + 0] = dr.x * w0.x + dr.y * w1.x + dr.z * w2.x;
            v[hh * 4 + 1] = dr.x * w0.y + dr.y * w1.y + dr.z * w2.y;
            v[hh * 4 + 2] = dr.x * w0.z + dr.y * w1.z + dr.z * w2.z;
            v[hh * 4 + 3] = dr.x * w0.w + dr.y * w1.w + dr.z * w2.w;
          }
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = pack_bf16(v[2 * e], v[2 * e + 1]) & (((bw >> (p4 * 4 + e)) & 0x00010001u) * 0xFFFFu);
          sts128(so + ((((uint32_t)(4 * sub + p4)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&act_ready[t * 4 + c]);
          mbar_arrive(&st_ready[t * 4 + c]);
        }
        ++sd_groups[t];
      }
      // ---- the nine layers
      for (int b = 0; b < kNB; ++b) {
        for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
          const uint32_t act_row_addr = act_base + (uint32_t)(t * 4 * kChunk) + (uint32_t)row_local * 128u;
          const uint32_t tacc = tmem_base + (uint32_t)t * 256u + ((uint32_t)(q * 32) << 16);
          const uint64_t ds2 = pack64(__float_as_uint(ds[t]), __float_as_uint(ds[t]));
          mbar_wait(&tfull[t], lc[t] & 1u);
          tc_fence_after();
          wait_store(t);
          uint32_t bw[2] = {0u, 0u};
          if (b >= 1) {
            bw[0] = bits_word(t, part);
            bw[1] = bits_word(t, 4 + part);
            bits_release(t);
          }
          uint32_t r32[32];
          tmem_ld_32x32(tacc + (uint32_t)((part >> 1) * 64 + sub * 32), r32);
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int c = 2 * h + (part >> 1);
            const int c0 = c * 64 + sub * 32;
            const uint32_t bwh = h == 0 ? bw[0] : bw[1];
            tmem_ld_wait_regs<32>(r32);
            if (h == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[t]);
            }
            const uint32_t so = act_row_addr + (uint32_t)c * kChunk;
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) {
              float wa[8];
              if (b == 1) {
                const float4 a0 = lds128(wa_addr + (uint32_t)(c0 + p4 * 8) * 4u);
                const float4 a1 = lds128(wa_addr + (uint32_t)(c0 + p4 * 8 + 4) * 4u);
                wa[0] = a0.x; wa[1] = a0.y; wa[2] = a0.z; wa[3] = a0.w; wa[4] = a1.x; wa[5] = a1.y; wa[6] = a1.z; wa[7] = a1.w;
              }
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                uint64_t x = pack64(r32[p4 * 8 + 2 * e], r32[p4 * 8 + 2 * e + 1]);
                if (b == 1) x = fma_f32x2(ds2, pack64(__float_as_uint(wa[2 * e]), __float_as_uint(wa[2 * e + 1])), x);
                uint32_t v = cvt_bf16x2<false>(x);
                if (b >= 1) v &= ((bwh >> (p4 * 4 + e)) & 0x00010001u) * 0xFFFFu;
                pk[e] = v;
              }
              sts128(so + ((((uint32_t)(4 * sub + p4)) ^ swz) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
            if (h == 0) tmem_ld_32x32(tacc + (uint32_t)(c0 + 128), r32);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (b < kNB - 1) mbar_arrive(&act_ready[t * 4 + c]);
              if (b >= 1) mbar_arrive(&st_ready[t * 4 + c]);
            }
          }
          if (b >= 1) ++sd_groups[t];
          ++lc[t];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kEncWarp0) tmem_dealloc_pair<512>(tmem_base);
}

// X0[:, 64:128] = PE(dir) of the point's ray (the dir-layer weight gradient's second operand): one 16 B copy per thread
// from the per-ray table.  Done here rather than by the chain's service warps: their plain stores delayed the store
// issue (measured 0.14 ms per fine pass against ~0.03 ms for this kernel).
__global__ void __launch_bounds__(256)
fill_x0_dir_kernel(const uint4* __restrict__ dir_pe, uint4* __restrict__ x0, long long P, int n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P * 8; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i >> 3;
    const int k = (int)(i & 7);
    x0[p * 16 + 8 + k] = __ldg(dir_pe + (p / n) * 8 + k);
  }
}

}  // namespace

namespace nmx {

int64_t chain2_train_scratch_bytes(int64_t n_rays) { return 13312 + n_rays * 512; }

int launch_chain2_train(const Chain2TrainLaunch& a, cudaStream_t stream) {
  if (a.P <= 0) return 0;
  Maps maps;
  Params prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  for (int l = 0; l < kNL; ++l) {
    const int N = l == kNL - 1 ? 128 : 256;
    if ((rc = make_tmap_bf16_2d(&maps.w[l], a.w_ptr[l], (uint64_t)N, (uint64_t)a.w_k[l], (uint64_t)a.w_k[l], (uint32_t)(N / 2)))) return rc;
  }
  if ((rc = make_tmap_bf16_2d(&maps.save, a.save_base, (uint64_t)a.save_rows, 256, 256, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&maps.hd, a.hd, (uint64_t)a.P, 128, 128, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&maps.x0, a.x0, (uint64_t)a.P, 128, 128, 128))) return rc;
  prm.P = (int)a.P; prm.params = a.params; prm.out = a.out;
  prm.rays = a.rays; prm.ray_stride = a.ray_stride; prm.z = a.z; prm.p0 = 0; prm.b0 = 0; prm.n_per_ray = a.n_per_ray;
  prm.alpha_b_off = a.alpha_b_off; prm.rgb_b_off = a.rgb_b_off;
  { static int dbg = -1; if (dbg < 0) dbg = experiment_env("NMX_CHAIN2T_DBG"); prm.dbg = dbg; }
  prm.bits = a.bits; prm.cap = (int)a.cap;
  float* consts = a.scratch;          // [constants block (13 KB) | per-ray dir bias]
  float* dir_bias = a.scratch + 3328;
  prm.consts = consts; prm.dir_bias = dir_bias;
  const long long n_rays = (a.P + a.n_per_ray - 1) / a.n_per_ray;
  {
    RayPrep rp;
    memset(&rp, 0, sizeof(rp));
    rp.params = a.params;
    for (int l = 0; l < kNL; ++l) rp.bias_off[l] = a.bias_off[l];
    rp.alpha_w_off = a.alpha_w_off; rp.rgb_w_off = a.rgb_w_off;
    rp.rays = a.rays; rp.ray_stride = a.ray_stride; rp.b0 = 0; rp.B = n_rays; rp.n_freqs_dir = a.n_freqs_dir;
    rp.dir_w_off = a.dir_w_off; rp.dir_ldw = a.dir_ldw; rp.consts = consts; rp.dir_bias = dir_bias; rp.dir_pe = a.dir_pe;
    if ((rc = launch_ray_prep(rp, stream))) return rc;
  }
  fill_x0_dir_kernel<<<grid_for(a.P * 8, 256, 16), 256, 0, stream>>>((const uint4*)a.dir_pe, (uint4*)a.x0, a.P, a.n_per_ray);
  NMX_LAUNCH_CHECK();
  static bool attr[64] = {};
  if (once_per_device(attr))
    NMX_CUDA(cudaFuncSetAttribute(mlp_chain2_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemT::kAlloc));
  const int num_pt = (int)((a.P + 255) / 256);
  int clusters = (num_pt + 1) / 2;
  if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
  if (clusters < 1) clusters = 1;
  double flops = 0.0;
  for (int l = 0; l < kNL; ++l) flops += 2.0 * a.P * (l == kNL - 1 ? 128 : 256) * 64.0 * (l == 0 ? 1 : (l == kSkipL ? 5 : 4));
  prof_begin(3, flops, stream);
  mlp_chain2_train_kernel<<<clusters * 2, kThreads, SmemT::kAlloc, stream>>>(maps, prm);
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_chain2_bwd(const Chain2BwdLaunch& a, cudaStream_t stream) {
  if (a.P <= 0) return 0;
  MapsB maps;
  ParamsB prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  for (int b = 0; b < kNB; ++b)
    if ((rc = make_tmap_bf16_2d(&maps.w[b], a.w_ptr[b], 256, (uint64_t)a.w_k[b], (uint64_t)a.w_k[b], 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&maps.save, a.save_base, (uint64_t)a.save_rows, 256, 256, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&maps.hd, a.ghd, (uint64_t)a.P, 128, 128, 128))) return rc;
  prm.P = (int)a.P; prm.params = a.params; prm.d_out = a.d_out; prm.bits = a.bits; prm.cap = (int)a.cap;
  prm.alpha_w_off = a.alpha_w_off; prm.rgb_w_off = a.rgb_w_off; prm.bits_wm = a.bits_word_major;
  static bool attr[64] = {};
  if (once_per_device(attr))
    NMX_CUDA(cudaFuncSetAttribute(mlp_chain2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemB::kAlloc));
  const int num_pt = (int)((a.P + 255) / 256);
  int clusters = (num_pt + 1) / 2;
  if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
  if (clusters < 1) clusters = 1;
  prof_begin(4, 2.0 * a.P * 256.0 * 64.0 * (2 + 8 * 4), stream);
  mlp_chain2_bwd_kernel<<<clusters * 2, kThreads, SmemB::kAlloc, stream>>>(maps, prm);
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmx
