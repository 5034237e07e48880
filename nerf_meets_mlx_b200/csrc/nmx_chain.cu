// K3 fused forward chain: the whole NeRF MLP (models/NeRF.py:201-243) for a 128-point tile in ONE persistent kernel.
//
// Activations never leave the SM between layers: the epilogue of layer l writes relu(acc + b) as bf16 straight into
// the 128B-swizzled shared-memory tile that is the A operand of layer l+1 (in place, chunk by chunk, so the MMAs of
// layer l+1 start as soon as the first 64-column chunk exists), accumulators ping-pong between the two halves of
// TMEM, and the weights (2.4 MB bf16 per net, L2-resident) stream through a TMA ring in 64-wide K slabs.  The skip
// concat [input_pos, h] and the view-dir concat [feature, input_dir] are extra K slabs read from the resident
// encoded-input tile.  The N=1 / N=3 / N<=8 heads (alpha, rgb, output_linear) are evaluated by the epilogue threads
// from the values they already hold in registers.  When training, each finished chunk is also TMA-stored to the
// saved-activation buffers (the smem tile doubles as the staging buffer).
//
// Roles (320 threads): warp 0 = TMA producer (weight ring + encoded-input tile), warp 1 = MMA issuer (one lane),
// warps 2..9 = epilogue, two warps per TMEM lane quadrant; epilogue group g = (warp-2)/4 owns columns [128g, 128g+128).
// Roofline: tensor pipe (inference: ~0 HBM traffic; training: 512 B/point/layer of activation stores).
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_chain.cuh"

using namespace nmx;
using namespace nmx::sm100;

namespace {

constexpr int kThreads = 320;
constexpr int kStages = 3;
constexpr int kSlabBytes = 256 * 64 * 2;   // one 64-wide K slab of a 256-row weight matrix
constexpr int kChunkBytes = 128 * 64 * 2;  // one 128-row x 64-col bf16 activation chunk

struct Smem {
  static constexpr int kActOff = 0;                              // 4 chunks (128 x 256 bf16)
  static constexpr int kX0Off = kActOff + 4 * kChunkBytes;       // pos chunk, dir chunk
  static constexpr int kRingOff = kX0Off + 2 * kChunkBytes;
  static constexpr int kBiasOff = kRingOff + kStages * kSlabBytes;       // [kMaxChainLayers][256] fp32
  static constexpr int kW7Off = kBiasOff + kMaxChainLayers * 256 * 4;    // head-7 weights [8][256] fp32
  static constexpr int kWrgbOff = kW7Off + 8 * 256 * 4;                  // rgb weights [3][128] fp32
  static constexpr int kXchgOff = kWrgbOff + 3 * 128 * 4;                // [128][12] fp32 partial sums
  static constexpr int kBarOff = kXchgOff + 128 * 12 * 4;
  // barriers: full[S], empty[S], tfull[2], tempty[2], act_ready[4], x0pos_full, x0pos_empty, x0dir_full, x0dir_empty
  static constexpr int kNumBars = 2 * kStages + 2 + 2 + 4 + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kAlloc = kTotal + 1024;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float round_bf16(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }

__global__ void __launch_bounds__(kThreads, 1)
mlp_chain_fwd_kernel(const __grid_constant__ ChainMaps maps, const ChainParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_act = smem + Smem::kActOff;
  uint8_t* s_x0 = smem + Smem::kX0Off;
  uint8_t* s_ring = smem + Smem::kRingOff;
  float* s_bias = reinterpret_cast<float*>(smem + Smem::kBiasOff);
  float* s_w7 = reinterpret_cast<float*>(smem + Smem::kW7Off);
  float* s_wrgb = reinterpret_cast<float*>(smem + Smem::kWrgbOff);
  float* s_xchg = reinterpret_cast<float*>(smem + Smem::kXchgOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Smem::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* act_ready = tempty + 2;
  uint64_t* x0pos_full = act_ready + 4;
  uint64_t* x0pos_empty = x0pos_full + 1;
  uint64_t* x0dir_full = x0pos_empty + 1;
  uint64_t* x0dir_empty = x0dir_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Smem::kTmemPtrOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = (prm.P + 127) / 128;
  const int NL = prm.n_layers;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < NL; ++l) tma_prefetch_desc(&maps.w[l]);
    tma_prefetch_desc(&maps.x0);
    tma_prefetch_desc(&maps.save);
    tma_prefetch_desc(&maps.hd);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&act_ready[i], 1);
    mbar_init(x0pos_full, 1);
    mbar_init(x0pos_empty, 1);
    mbar_init(x0dir_full, 1);
    mbar_init(x0dir_empty, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  // stage biases and the register-head weights (fp32) once per CTA
  for (int i = threadIdx.x; i < NL * 256; i += kThreads) {
    int l = i >> 8, c = i & 255;
    s_bias[i] = (c < prm.L[l].N) ? prm.params[prm.L[l].bias_off + c] : 0.0f;
  }
  for (int i = threadIdx.x; i < prm.head7_n * 256; i += kThreads) s_w7[i] = prm.params[prm.head7_w_off + i];
  if (prm.rgb_layer >= 0)
    for (int i = threadIdx.x; i < 3 * 128; i += kThreads) s_wrgb[i] = prm.params[prm.rgb_w_off + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ====================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        if (it == 0) {
          mbar_arrive_expect_tx(x0pos_full, kChunkBytes);
          tma_load_2d(s_x0, &maps.x0, x0pos_full, 0, tile * 128);
          if (prm.uses_dir) {
            mbar_arrive_expect_tx(x0dir_full, kChunkBytes);
            tma_load_2d(s_x0 + kChunkBytes, &maps.x0, x0dir_full, prm.x0_dir_col, tile * 128);
          }
        }
        for (int l = 0; l < NL; ++l) {
          const int N = prm.L[l].N;
          for (int s = 0; s < prm.L[l].n_slabs; ++s) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* dst = s_ring + stage * kSlabBytes;
            mbar_arrive_expect_tx(&full[stage], (uint32_t)N * 128);
            tma_load_2d(dst, &maps.w[l], &full[stage], s * 64, 0);
            if (N > 128) tma_load_2d(dst + 128 * 128, &maps.w[l], &full[stage], s * 64, 128);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (l == 1 && it > 0 && prm.uses_dir) {
            // this tile's view-dir chunk: the previous tile's dir-layer MMAs must have finished reading the buffer
            mbar_wait(x0dir_empty, (uint32_t)((it - 1) & 1));
            mbar_arrive_expect_tx(x0dir_full, kChunkBytes);
            tma_load_2d(s_x0 + kChunkBytes, &maps.x0, x0dir_full, prm.x0_dir_col, tile * 128);
          }
          if (l == prm.pos_prefetch_layer) {
            const int ntile = tile + gridDim.x;
            if (ntile < num_tiles) {  // next tile's position chunk, once this tile's last reader (skip layer) is done
              mbar_wait(x0pos_empty, (uint32_t)(it & 1));
              mbar_arrive_expect_tx(x0pos_full, kChunkBytes);
              tma_load_2d(s_x0, &maps.x0, x0pos_full, 0, ntile * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================================== MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      uint32_t lcount = 0;
      uint32_t rc[4] = {0, 0, 0, 0};
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        bool pos_waited = false, dir_waited = false;
        for (int l = 0; l < NL; ++l, ++lcount) {
          const int as = lcount & 1;
          const uint32_t aphase = (lcount >> 1) & 1;
          const int N = prm.L[l].N;
          const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
          mbar_wait(&tempty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * 256;
          for (int s = 0; s < prm.L[l].n_slabs; ++s) {
            const int src = prm.L[l].src[s];
            uint32_t a_addr;
            if (src == kSrcPos) {
              if (!pos_waited) { mbar_wait(x0pos_full, (uint32_t)(it & 1)); pos_waited = true; }
              a_addr = smem_u32(s_x0);
            } else if (src == kSrcDir) {
              if (!dir_waited) { mbar_wait(x0dir_full, (uint32_t)(it & 1)); dir_waited = true; }
              a_addr = smem_u32(s_x0 + kChunkBytes);
            } else {
              mbar_wait(&act_ready[src], rc[src] & 1);
              rc[src]++;
              a_addr = smem_u32(s_act + src * kChunkBytes);
            }
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + stage * kSlabBytes);
            const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
            const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (s | k) != 0);
            umma_commit(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[as]);
          if (l == prm.pos_last_layer) umma_commit(x0pos_empty);
          if (l == prm.dir_layer) umma_commit(x0dir_empty);
        }
      }
    }
  } else {
    // ====================================================== epilogue (8 warps)
    const int q = warp & 3;
    const int g = (warp - 2) >> 2;              // column group
    const int row_local = q * 32 + lane;
    const bool gleader = ((warp - 2) & 3) == 0 && lane == 0;  // one thread per group issues barriers' side effects
    const uint32_t swz = (uint32_t)(row_local & 7);
    const int bar_id = 1 + g;
    uint32_t lcount = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int row = tile * 128 + row_local;
      float hp[8];  // head-7 partial dot products over this thread's columns
#pragma unroll
      for (int o = 0; o < 8; ++o) hp[o] = 0.0f;
      float rgbp[3] = {0.0f, 0.0f, 0.0f};
      for (int l = 0; l < NL; ++l, ++lcount) {
        const int as = lcount & 1;
        const uint32_t aphase = (lcount >> 1) & 1;
        const int N = prm.L[l].N;
        const int relu = prm.L[l].relu;
        const int nck = (N == 256) ? 2 : 1;
        mbar_wait(&tfull[as], aphase);
        tc_fence_after();
        if (prm.save) {  // stores issued from this group's chunks must have finished reading shared memory
          if (gleader) tma_store_wait_read<0>();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        }
        const float* bias = s_bias + l * 256;
        for (int ci = 0; ci < nck; ++ci) {
          const int c = (N == 256) ? (2 * g + ci) : g;
          uint8_t* so = s_act + c * kChunkBytes + row_local * 128;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = c * 64 + h * 32;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + as * 256 + c0 + ((uint32_t)(q * 32) << 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float x = __uint_as_float(r[p4 * 8 + e]) + bias[c0 + p4 * 8 + e];
                v[e] = relu ? fmaxf(x, 0.0f) : x;
              }
              const uint32_t piece = ((uint32_t)(h * 4 + p4) ^ swz) << 4;
              *reinterpret_cast<uint4*>(so + piece) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              if (l == prm.head7_layer) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float xr = round_bf16(v[e]);
#pragma unroll
                  for (int o = 0; o < 8; ++o)
                    if (o < prm.head7_n) hp[o] += xr * s_w7[o * 256 + c0 + p4 * 8 + e];
                }
              }
              if (l == prm.rgb_layer) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float xr = round_bf16(v[e]);
                  const int col = c0 + p4 * 8 + e;
                  rgbp[0] += xr * s_wrgb[col];
                  rgbp[1] += xr * s_wrgb[128 + col];
                  rgbp[2] += xr * s_wrgb[256 + col];
                }
              }
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (gleader) {
            if (prm.L[l].feeds_next) mbar_arrive(&act_ready[c]);
            if (prm.save && prm.L[l].save_kind == 1) {
              tma_store_2d(&maps.save, s_act + c * kChunkBytes, c * 64, prm.L[l].save_row0 + tile * 128);
              tma_store_commit();
            } else if (prm.save && prm.L[l].save_kind == 2) {
              tma_store_2d(&maps.hd, s_act + c * kChunkBytes, c * 64, tile * 128);
              tma_store_commit();
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
      }
      // ---- register heads: combine the two column groups' partial sums and write the raw outputs
      float* xr = s_xchg + row_local * 12;
      if (g == 1) {
#pragma unroll
        for (int o = 0; o < 8; ++o) xr[o] = hp[o];
        xr[8] = rgbp[0];
        xr[9] = rgbp[1];
        xr[10] = rgbp[2];
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (g == 0 && row < prm.P) {
        float* o_row = prm.out + (size_t)row * prm.out_cols;
        if (prm.rgb_layer >= 0) {
          const float* pb = prm.params;
          float4 o4;
          o4.x = rgbp[0] + xr[8] + pb[prm.rgb_b_off + 0];
          o4.y = rgbp[1] + xr[9] + pb[prm.rgb_b_off + 1];
          o4.z = rgbp[2] + xr[10] + pb[prm.rgb_b_off + 2];
          o4.w = hp[0] + xr[0] + pb[prm.head7_b_off];
          *reinterpret_cast<float4*>(o_row) = o4;
        } else {
#pragma unroll
          for (int o = 0; o < 8; ++o)
            if (o < prm.head7_n) o_row[o] = hp[o] + xr[o] + prm.params[prm.head7_b_off + o];
        }
      }
    }
    if (gleader) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

}  // namespace

namespace nmx {

int launch_chain_fwd(const ChainMaps& maps, const ChainParams& prm, cudaStream_t stream) {
  if (prm.P <= 0) return 0;
  static bool attr = false;
  if (!attr) {
    NMX_CUDA(cudaFuncSetAttribute(mlp_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kAlloc));
    attr = true;
  }
  int tiles = (prm.P + 127) / 128;
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  mlp_chain_fwd_kernel<<<grid, kThreads, Smem::kAlloc, stream>>>(maps, prm);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmx
