// Fused MLP forward chain on CTA PAIRS with TWO tiles in ping-pong (inference path of the 8 x 256 view-dir NeRF,
// models/NeRF.py:201-243 evaluated through run_model, models/NeRF.py:25-48, with the Embedder PE of
// models/embedding.py:35-71 computed in the kernel).
//
// Why: inside one 128-point tile the layers are strictly dependent, so after every layer the tensor pipe waits for the
// accumulator -> epilogue -> shared memory -> next MMA hand-off (about 1200 of 3250 clocks per layer in nmx_chain.cu).
// A second, independent tile fills that time -- but two tiles do not fit in one SM next to full 32 KB weight slabs.  With
// tcgen05 cta_group::2 the two CTAs of a cluster share every MMA (M = 256: 128 rows from each CTA) and each CTA stages
// only HALF of the weight slab, so per CTA: 2 tiles x 5 chunks of activations (160 KB) + a 3 x 16 KB weight ring fit.
// (A 4-stage ring with biases / head weights read from global memory instead was measured: no gain on the weight side,
// slower epilogue.)
//
// A cluster works on two 256-point "pair tiles" X and Y at a time.  The leader's MMA warp alternates
//   layer l of X, layer l of Y, layer l+1 of X, ...
// each into that slot's own 256-column accumulator (2 x 256 = all 512 TMEM columns), so while the sixteen epilogue warps
// of both CTAs drain and re-encode X's accumulator, the tensor pipe runs Y's layer, and vice versa.
//
// The view-dir input of the dir layer is per RAY, not per point: its contribution W_dir[:, W:] PE(dir) is precomputed
// per ray (dir_bias_kernel, same bf16 operand rounding as the tensor-core path) and added in the dir layer's epilogue, so
// no view-dir operand tile is needed in shared memory.
//
// Barriers (same offsets in both CTAs; "leader" = the copy in cluster rank 0 is the one waited on):
//   full[s]   leader   1 arrival (expect_tx) + the bytes of BOTH CTAs' half-slab TMA loads
//   empty[s]  each     leader's multicast tcgen05.commit after the slab's MMAs
//   tfull[t]  each     multicast commit after the last slab of a layer of slot t
//   tempty[t] each     16 local arrivals: every epilogue warp of the CTA has read slot t's accumulator; CTA 1's otherwise
//                      idle MMA warp relays its completion to the leader's tempty_peer[t] with ONE remote arrival
//   act[t][c] each     8 local arrivals (the warps that own chunk c): slot t's next-layer input chunk is in shared memory;
//                      relayed likewise
//                      (act_peer[t][c]); the leader's MMA warp waits for its local and the peer barrier
//   posf[t]   leader   2 arrivals: both CTAs' encoder warps have written slot t's PE(pos) chunk
//   pose[t]   each     multicast commit after the skip layer (last reader of the PE(pos) chunk)
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_chain.cuh"
#include "nmx_chain_dev.cuh"
#include "nmx_gemm.cuh"
#include <cstring>
#include <cstdlib>

using namespace nmx;
using namespace nmx::sm100;
using namespace nmx::chain_dev;

namespace {

constexpr int kNL = 10;          // 8 trunk layers, feature layer, dir layer
constexpr int kSkipL = 5;        // the trunk layer whose input is [PE(pos), h]
constexpr int kEpiWarps = 16;
constexpr int kEncWarp0 = 16;    // warps 16, 17: input encoders of slot 0 / slot 1 (warp 16 also owns TMEM alloc)
constexpr int kTmaWarp = 18;
constexpr int kMmaWarp = 19;
constexpr int kThreads = 20 * 32;
constexpr int kStages = 3;
constexpr int kChunk = 128 * 64 * 2;   // one 128-row x 64-col bf16 chunk
constexpr int kHalfSlab = 128 * 64 * 2;

struct Smem2 {
  static constexpr int kActOff = 0;                               // [2 slots][5 chunks]: 4 activation chunks + PE(pos)
  static constexpr int kRingOff = kActOff + 2 * 5 * kChunk;
  static constexpr int kBiasOff = kRingOff + kStages * kHalfSlab; // [kNL][256] fp32
  static constexpr int kWaOff = kBiasOff + kNL * 256 * 4;         // w_alpha [256] fp32
  static constexpr int kWrgbOff = kWaOff + 256 * 4;               // w_rgb [3][128] fp32
  static constexpr int kBarOff = kWrgbOff + 3 * 128 * 4;
  static constexpr int kNumBars = 2 * kStages + 2 + 2 + 8 + 2 + 2 + 2 + 8;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kAlloc = kTotal + 1024;
  static_assert(kAlloc <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};

struct Chain2Params {
  int P;                     // points in this launch
  const float* params;       // packed fp32 parameters (biases, tiny heads)
  float* out;                // [P, 4] fp32 (rgb, sigma)
  int bias_off[kNL];
  int alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off;
  const float* rays; int ray_stride; const float* z;
  long long p0, b0; int n_per_ray;
  const float* dir_bias;     // [rays of this launch, 128] fp32
  const float* consts;       // 16 B aligned: bias [kNL][256], w_alpha [256], w_rgb [3][128] (chain2_consts_kernel)
  // timing-only experiments (NMX_CHAIN2_DBG, results are garbage): 1 encoders skip their work, 2 epilogue math skipped,
  // 4 no weight TMA, 16 / 32 only the leader / only CTA 1 loads weights, 64 both load with local completion, 128 both load
  // the same half, 256 non-suspending poll of the ring's empty barriers, 512 epilogue = barrier traffic only, 1024 one
  // tensor map for all layers, 4096 no barrier relay from CTA 1, 8192 no proxy fence
  int dbg;
};
struct Chain2Maps { CUtensorMap w[kNL]; };

__device__ __forceinline__ int layer_slabs(int l) { return l == 0 ? 1 : (l == kSkipL ? 5 : 4); }
// shared-memory chunk (0..3 activation, 4 PE(pos)) read by K slab s of layer l
__device__ __forceinline__ int slab_src(int l, int s) { return l == 0 ? 4 : (l == kSkipL ? (s == 0 ? 4 : s - 1) : s); }
__device__ __forceinline__ int layer_N(int l) { return l == kNL - 1 ? 128 : 256; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_chain2_kernel(const __grid_constant__ Chain2Maps maps, const Chain2Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* s_bias = reinterpret_cast<float*>(smem + Smem2::kBiasOff);
  float* s_wa = reinterpret_cast<float*>(smem + Smem2::kWaOff);
  float* s_wrgb = reinterpret_cast<float*>(smem + Smem2::kWrgbOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Smem2::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;    // [2]
  uint64_t* tempty = tfull + 2;         // [2]
  uint64_t* act_ready = tempty + 2;     // [2][4]
  uint64_t* pos_full = act_ready + 8;   // [2]
  uint64_t* pos_empty = pos_full + 2;   // [2]
  uint64_t* tempty_peer = pos_empty + 2;    // [2]    leader: CTA 1's forwarder warp relays "all 16 local warps arrived"
  uint64_t* act_peer = tempty_peer + 2;     // [2][4]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem2::kTmemPtrOff);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader_cta = rank == 0;
  const int num_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const int num_pt = (prm.P + 255) / 256;  // pair tiles of 256 points

  if (warp == kTmaWarp && lane == 0) {
    for (int l = 0; l < kNL; ++l) tma_prefetch_desc(&maps.w[l]);
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
      }
      mbar_init(&pos_full[t], 2);
      mbar_init(&pos_empty[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == kEncWarp0) tmem_alloc_pair<512>(tmem_ptr);
  for (int i = threadIdx.x; i < kNL * 256 + 256 + 384; i += kThreads) s_bias[i] = __ldg(prm.consts + i);  // bias | w_alpha | w_rgb
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t act_base = smem_u32(smem + Smem2::kActOff);

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
              if NMX_DBG(prm, 256) {  // experiment: non-suspending poll
                uint32_t ok;
                do {
                  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                               : "=r"(ok) : "r"(smem_u32(&empty[stage])), "r"(phase ^ 1) : "memory");
                } while (!ok);
              } else {
                mbar_wait(&empty[stage], phase ^ 1);
              }
              if NMX_DBG(prm, 4) {  // experiment: no weight traffic (results are garbage)
                if (leader_cta) mbar_arrive(&full[stage]);
              } else if NMX_DBG(prm, 16) {  // experiment: only the leader loads its half (results are garbage)
                if (leader_cta) {
                  mbar_arrive_expect_tx(&full[stage], bytes);
                  tma_load_2d_pair(smem + Smem2::kRingOff + stage * kHalfSlab, &maps.w[l], &full[stage], s * 64, 0);
                }
              } else if NMX_DBG(prm, 64) {  // experiment: both load, CTA 1's completion stays local (leader does not wait for it)
                mbar_arrive_expect_tx(&full[stage], bytes);
                tma_load_2d(smem + Smem2::kRingOff + stage * kHalfSlab, &maps.w[l], &full[stage], s * 64, (int)rank * half_rows);
              } else if NMX_DBG(prm, 32) {  // experiment: only CTA 1 loads its half
                if (leader_cta) mbar_arrive_expect_tx(&full[stage], bytes);
                else tma_load_2d_pair(smem + Smem2::kRingOff + stage * kHalfSlab, &maps.w[l], &full[stage], s * 64, half_rows);
              } else if NMX_DBG(prm, 1024) {  // experiment: ONE tensor map for every layer (descriptor-cache test)
                if (leader_cta) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                tma_load_2d_pair(smem + Smem2::kRingOff + stage * kHalfSlab, &maps.w[l == kNL - 1 ? kNL - 1 : 1], &full[stage],
                                 (s & 3) * 64, (int)rank * half_rows);
              } else {
                if (leader_cta) mbar_arrive_expect_tx(&full[stage], 2 * bytes);
                tma_load_2d_pair(smem + Smem2::kRingOff + stage * kHalfSlab, &maps.w[l], &full[stage], s * 64,
                                 NMX_DBG(prm, 128) ? 0 : (int)rank * half_rows);
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ====================================================== MMA issuer (leader CTA only)
    if (leader_cta) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t lc[2] = {0, 0};    // layers issued so far per slot (accumulator generations)
      uint32_t gen[2] = {0, 0};   // act_ready generations produced so far per slot
      uint32_t tc[2] = {0, 0};    // tiles started per slot
      const uint32_t smem16 = smem_u32(smem) >> 4;
      const uint32_t ring16 = smem_u32(smem + Smem2::kRingOff) >> 4;
      constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;  // SBO 1024 B, version 1, SWIZZLE_128B
      for (int j = 0;; ++j) {
        const int tx = cluster_id + (2 * j) * num_clusters;
        if (tx >= num_pt) break;
        const bool valid_y = tx + num_clusters < num_pt;
        for (int l = 0; l < kNL; ++l) {
          const uint32_t idesc = make_idesc_bf16(256, layer_N(l), 0, 0);
          const int ns = layer_slabs(l);
          for (int t = 0; t < (valid_y ? 2 : 1); ++t) {
            mbar_wait(&tempty[t], (lc[t] & 1u) ^ 1u);  // the previous layer's epilogue has drained this accumulator
            if (!NMX_DBG(prm, 4096)) mbar_wait_cluster(&tempty_peer[t], (lc[t] & 1u) ^ 1u);  // ... in CTA 1 as well
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
            for (int s = 0; s < ns; ++s) {
              const int src = slab_src(l, s);
              mbar_wait(&full[stage], phase);
              if (src == 4) mbar_wait_cluster(&pos_full[t], tc[t] & 1u);
              else {
                mbar_wait(&act_ready[t * 4 + src], (gen[t] - 1u) & 1u);
                if (!NMX_DBG(prm, 4096)) mbar_wait_cluster(&act_peer[t * 4 + src], (gen[t] - 1u) & 1u);
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
    }
    if (!leader_cta && !NMX_DBG(prm, 4096)) {
      // ---- CTA 1: this warp relays the local epilogue barriers to the leader, ONE remote arrival per barrier phase
      // (sixteen warps arriving remotely with cluster-scope release semantics each was measurably slow)
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
    // ====================================================== input encoder of slot (warp - kEncWarp0): PE(pos) -> chunk 4
    const int t = warp - kEncWarp0;
    const uint32_t pos_addr = act_base + (uint32_t)((t * 5 + 4) * kChunk);
    const uint32_t posf_leader = mapa_u32(smem_u32(&pos_full[t]), 0);
    uint32_t tcnt = 0;
    for (int j = 0;; ++j, ++tcnt) {
      const int tile = cluster_id + (2 * j + t) * num_clusters;
      if (tile >= num_pt) break;
      if (tcnt > 0) mbar_wait(&pos_empty[t], (tcnt - 1u) & 1u);
#pragma unroll 1
      for (int rr = 0; rr < (NMX_DBG(prm, 1) ? 0 : 4); ++rr) {
        const int row_local = rr * 32 + lane;
        const long long row = (long long)tile * 256 + (long long)rank * 128 + row_local;
        encode_pos_row(prm.rays, prm.ray_stride, prm.z, prm.p0 + row, prm.n_per_ray, row < prm.P,
                       pos_addr + (uint32_t)row_local * 128u, (uint32_t)(row_local & 7));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(posf_leader);
    }
  } else {
    // ====================================================== epilogue (warps 0..15 of both CTAs)
    const int q = warp & 3;       // TMEM lane quadrant
    const int part = warp >> 2;   // 16-column share of every 64-column chunk
    const int row_local = q * 32 + lane;
    const uint32_t swz = (uint32_t)(row_local & 7);
    const uint32_t bias_base = smem_u32(s_bias);
    const uint32_t wa_addr = smem_u32(s_wa), wrgb_addr = smem_u32(s_wrgb);
    float hb[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (part == 0) {
      hb[0] = __ldg(prm.params + prm.rgb_b_off + 0);
      hb[1] = __ldg(prm.params + prm.rgb_b_off + 1);
      hb[2] = __ldg(prm.params + prm.rgb_b_off + 2);
      hb[3] = __ldg(prm.params + prm.alpha_b_off);
    }
    uint32_t lc[2] = {0, 0};
    float alpha_p[2] = {0.0f, 0.0f};  // alpha-head partial sums of the two slots (layer 7 -> written after layer 9)
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
          const uint32_t tacc = tmem_base + (uint32_t)t * 256u + ((uint32_t)(q * 32) << 16);
          float hp[8];
#pragma unroll
          for (int o = 0; o < 8; ++o) hp[o] = 0.0f;
          float rgbp[3] = {0.0f, 0.0f, 0.0f};
          const float* dbias = nullptr;
          if (l == kNL - 1) {
            const long long rr = row < prm.P ? row : (long long)prm.P - 1;
            dbias = prm.dir_bias + ((prm.p0 + rr) / prm.n_per_ray - prm.b0) * 128;
          }
          if (!NMX_DBG(prm, 2048)) mbar_wait(&tfull[t], lc[t] & 1u);  // 2048: experiment, no tfull polling (garbage results)
          tc_fence_after();
          if NMX_DBG(prm, 512) {  // experiment: barrier traffic only (no TMEM loads, no proxy fences)
            for (int h = 0; h < nsteps / 2; ++h) {
              __syncwarp();
              if (l < kNL - 1 && lane == 0) mbar_arrive(&act_ready[t * 4 + 2 * h + (part >> 1)]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[t]);
            ++lc[t];
            continue;
          }
          // Half steps: chunks (0, 1) then (2, 3); every warp takes 32 columns of one chunk per step.  With two tiles in
          // flight the time to the FIRST chunk no longer matters (the other tile's MMAs cover it); the total does, and
          // two 32-column steps cost less than four 16-column steps (fewer fence / barrier round trips per warp).
          const int nh = nsteps / 2;
          const int sub = part & 1;
          uint32_t r32[32];
          tmem_ld_32x32(tacc + (uint32_t)((part >> 1) * 64 + sub * 32), r32);
#pragma unroll 1
          for (int h = 0; h < nh; ++h) {
            const int c = 2 * h + (part >> 1);
            const int c0 = c * 64 + sub * 32;
            tmem_ld_wait_regs<32>(r32);
            if (h == nh - 1) {
              // the accumulator is now completely in registers: hand it back before the math of the last chunks
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[t]);
            }
            if NMX_DBG(prm, 2) {
            } else if (l == kNL - 1) {  // per-ray view-dir term of the dir layer
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4) {
                const float4 d = __ldg(reinterpret_cast<const float4*>(dbias + c0) + i4);
                r32[i4 * 4 + 0] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 0]) + d.x);
                r32[i4 * 4 + 1] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 1]) + d.y);
                r32[i4 * 4 + 2] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 2]) + d.z);
                r32[i4 * 4 + 3] = __float_as_uint(__uint_as_float(r32[i4 * 4 + 3]) + d.w);
              }
              epi_cols<true, 2, 4, false, false>(r32, c, c0, 4 * sub, 0, bias_addr, act_row_addr, swz, wrgb_addr, 0, hp, rgbp, 0u);
            } else if (l == 7) {
              epi_cols<true, 1, 4, false, false>(r32, c, c0, 4 * sub, 0, bias_addr, act_row_addr, swz, wa_addr, 1, hp, rgbp, 0u);
            } else if (l == 8) {
              epi_cols<false, 0, 4, false, false>(r32, c, c0, 4 * sub, 0, bias_addr, act_row_addr, swz, 0u, 0, hp, rgbp, 0u);
            } else {
              epi_cols<true, 0, 4, false, false>(r32, c, c0, 4 * sub, 0, bias_addr, act_row_addr, swz, 0u, 0, hp, rgbp, 0u);
            }
            if (h + 1 < nh) tmem_ld_32x32(tacc + (uint32_t)(c0 + 128), r32);
            if (!NMX_DBG(prm, 8192)) fence_proxy_async_smem();  // 8192: timing experiment without the proxy fence
            __syncwarp();
            if (l < kNL - 1 && lane == 0) mbar_arrive(&act_ready[t * 4 + c]);
          }
          if (l == 7) alpha_p[t] = hp[0];
          if (l == kNL - 1) {
            // heads: combine the four column parts' partial sums through shared memory (chunk 3 of this slot is free:
            // the dir layer reads it no more and writes chunks 0, 1 only) and write raw = (rgb, sigma)
            float* xchg = reinterpret_cast<float*>(smem + Smem2::kActOff + (t * 5 + 3) * kChunk);
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
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // xchg is reused by the next tile's layers
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

// Everything the pair chains need per launch besides the weights, in ONE kernel (was three launches):
//  * consts: aligned copy of what the epilogue reads per column: bias [kNL][256] (zero padded), w_alpha [256],
//    w_rgb [3][128] (blocks 0..25 write 128 entries each);
//  * per ray: the view-dir encoding PE(dir) (models/embedding.py:35-71, bands k^2, bf16-rounded like every MLP operand),
//    optionally stored as a bf16 [rays, 64] table (training: X0[:, 64:128] is filled from it), and the view-dir term of
//    the dir layer out[b, o] = sum_k bf16(W_dir[o, W + k]) * PE(dir_b)[k] (fp32 accumulate, increasing k).  A block keeps
//    the 27 x 128 bf16-rounded weight columns in shared memory and walks over rays.
constexpr int kPrepWStride = 65;  // floats per output row of the staged view-dir weights (odd: conflict-free both ways)

__global__ void __launch_bounds__(128)
ray_prep_kernel(const RayPrep a) {
  __shared__ float w_s[128 * kPrepWStride];
  __shared__ float pe[2][64];
  const int o = threadIdx.x;
  {
    const int i = blockIdx.x * 128 + o;
    if (i < kNL * 256 + 256 + 384) {
      float v;
      if (i < kNL * 256) {
        const int l = i >> 8, c = i & 255;
        v = c < (l == kNL - 1 ? 128 : 256) ? a.params[a.bias_off[l] + c] : 0.0f;
      } else if (i < kNL * 256 + 256) {
        v = a.params[a.alpha_w_off + (i - kNL * 256)];
      } else {
        v = a.params[a.rgb_w_off + (i - kNL * 256 - 256)];
      }
      a.consts[i] = v;
    }
  }
  const int in_dir = 3 + 6 * a.n_freqs_dir;
  {  // stage W_dir[:, W:W+in_dir] (bf16-rounded): a warp reads one output row per step, lanes along k (coalesced), eight
     // rows in flight (was one strided load per thread and k: 27 dependent round trips before the first ray)
    const float* Wd = a.params + a.dir_w_off + 256;
    const int lane = o & 31, warp = o >> 5;
#pragma unroll 8
    for (int r = warp; r < 128; r += 4) {
      if (lane < in_dir) w_s[r * kPrepWStride + lane] = __bfloat162float(__float2bfloat16_rn(__ldg(Wd + (size_t)r * a.dir_ldw + lane)));
      if (lane + 32 < in_dir)
        w_s[r * kPrepWStride + lane + 32] = __bfloat162float(__float2bfloat16_rn(__ldg(Wd + (size_t)r * a.dir_ldw + lane + 32)));
    }
  }
  __nv_bfloat16* table = reinterpret_cast<__nv_bfloat16*>(a.dir_pe);
  int buf = 0;
  for (long long b = blockIdx.x; b < a.B; b += gridDim.x, buf ^= 1) {
    if (o < 64) {
      float v = 0.0f;
      if (o < in_dir) {
        const float* d = a.rays + (a.b0 + b) * a.ray_stride + (a.ray_stride - 3);
        if (o < 3) v = d[o];
        else {
          const int qq = o - 3, k = qq / 6, rr = qq - 6 * k, fn = rr / 3, dd = rr - 3 * fn;
          const float ang = __fmul_rn(d[dd], (float)(k * k));
          v = fn ? cosf(ang) : sinf(ang);
        }
      }
      const __nv_bfloat16 vb = __float2bfloat16_rn(v);
      pe[buf][o] = __bfloat162float(vb);
      if (table != nullptr) table[b * 64 + o] = vb;
    }
    __syncthreads();  // (double-buffered pe: one barrier per ray; the first one also covers the staged weights)
    float acc = 0.0f;
    for (int k = 0; k < in_dir; ++k) acc += w_s[o * kPrepWStride + k] * pe[buf][k];
    a.dir_bias[b * 128 + o] = acc;
  }
}

}  // namespace

namespace nmx {

int launch_ray_prep(const RayPrep& rp, cudaStream_t stream) {
  if (rp.B <= 0) return 0;
  // ~8 rays per block amortise the 14 KB weight staging; at most 4 blocks per SM
  long long blocks = (rp.B + 7) / 8;
  if (blocks > (long long)kNumSMs * 4) blocks = (long long)kNumSMs * 4;
  if (blocks < 26) blocks = 26;  // the constants block is written 128 entries per block
  ray_prep_kernel<<<(unsigned)blocks, 128, 0, stream>>>(rp);
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_chain2(const Chain2Launch& a, cudaStream_t stream) {
  if (a.P <= 0) return 0;
  Chain2Maps maps;
  Chain2Params prm;
  memset(&prm, 0, sizeof(prm));
  int rc;
  for (int l = 0; l < kNL; ++l) {
    const int N = l == kNL - 1 ? 128 : 256;
    if ((rc = make_tmap_bf16_2d(&maps.w[l], a.w_ptr[l], (uint64_t)N, (uint64_t)a.w_k[l], (uint64_t)a.w_k[l], (uint32_t)(N / 2)))) return rc;
    prm.bias_off[l] = a.bias_off[l];
  }
  prm.P = (int)a.P; prm.params = a.params; prm.out = a.out;
  prm.alpha_w_off = a.alpha_w_off; prm.alpha_b_off = a.alpha_b_off; prm.rgb_w_off = a.rgb_w_off; prm.rgb_b_off = a.rgb_b_off;
  prm.rays = a.rays; prm.ray_stride = a.ray_stride; prm.z = a.z; prm.p0 = a.p0; prm.b0 = a.p0 / a.n_per_ray;
  prm.n_per_ray = a.n_per_ray; prm.dir_bias = a.dir_bias;
  { static int dbg = -1; if (dbg < 0) dbg = experiment_env("NMX_CHAIN2_DBG"); prm.dbg = dbg; }
  // scratch: [constants block (13 KB, 256 B aligned) | per-ray dir bias]
  float* consts = a.dir_bias;
  float* dir_bias = a.dir_bias + 3328;
  prm.consts = consts; prm.dir_bias = dir_bias;
  const long long b1 = (a.p0 + a.P - 1) / a.n_per_ray;
  const long long n_rays = b1 - prm.b0 + 1;
  {
    RayPrep rp;
    memset(&rp, 0, sizeof(rp));
    rp.params = a.params;
    for (int l = 0; l < kNL; ++l) rp.bias_off[l] = a.bias_off[l];
    rp.alpha_w_off = a.alpha_w_off; rp.rgb_w_off = a.rgb_w_off;
    rp.rays = a.rays; rp.ray_stride = a.ray_stride; rp.b0 = prm.b0; rp.B = n_rays; rp.n_freqs_dir = a.n_freqs_dir;
    rp.dir_w_off = a.dir_w_off; rp.dir_ldw = a.dir_ldw; rp.consts = consts; rp.dir_bias = dir_bias; rp.dir_pe = nullptr;
    if ((rc = launch_ray_prep(rp, stream))) return rc;
  }
  static bool attr[64] = {};
  if (once_per_device(attr)) {
    NMX_CUDA(cudaFuncSetAttribute(mlp_chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2::kAlloc));
  }
  const int num_pt = (int)((a.P + 255) / 256);
  int clusters = (num_pt + 1) / 2;
  if (clusters > kNumSMs / 2) clusters = kNumSMs / 2;
  if (clusters < 1) clusters = 1;
  double flops = 0.0;
  for (int l = 0; l < kNL; ++l) flops += 2.0 * a.P * (l == kNL - 1 ? 128 : 256) * 64.0 * (l == 0 ? 1 : (l == kSkipL ? 5 : 4));
  prof_begin(2, flops, stream);
  mlp_chain2_kernel<<<clusters * 2, kThreads, Smem2::kAlloc, stream>>>(maps, prm);
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmx
