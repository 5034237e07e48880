// K3 fused forward chain: the whole NeRF MLP (models/NeRF.py:201-243) for a 128-point tile in ONE persistent kernel.
//
// Activations never leave the SM between layers: the epilogue of layer l writes relu(acc + b) as bf16 straight into
// the 128B-swizzled shared-memory tile that is the A operand of layer l+1 (in place), accumulators ping-pong between
// the two halves of TMEM, and the weights (2.4 MB bf16 per net, L2-resident) stream through a TMA ring.
//
// Pipelining inside one tile (the layers of a tile are strictly dependent): the epilogue turns the accumulator into
// the next layer's input in two steps of two 64-column chunks; the MMA warp starts layer l+1 on chunks 0,1 while the
// epilogue is still producing chunks 2,3.  Every K slab is one "piece" of 4 MMAs M128 x N256 x K16 described by a
// host-built 64-bit table entry in the constant bank, so the issue loop is a few dozen uniform-datapath instructions.
// The skip concat [input_pos, h] and the view-dir concat [feature, input_dir] are extra K slabs read from the
// resident encoded-input tile.  The N=1 / N=3 / N<=8 heads (alpha, rgb, output_linear) are evaluated by the epilogue
// threads from the values they already hold in registers.  When training, a dedicated warp TMA-stores each finished
// chunk to the saved-activation buffers (the smem tile doubles as the staging buffer).
//
// Roles (608 threads): warps 0..15 = epilogue (four warps per TMEM lane quadrant; in each half-phase two of them
// share a 64-column chunk, 32 columns each), warp 16 = activation-store issuer (training only), warp 17 = TMA producer
// (weight ring + encoded-input tile), warp 18 = MMA issuer.
// Roofline: tensor pipe (inference: ~0 HBM traffic; training: 512 B/point/layer of activation stores).
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_chain.cuh"
#include "nmx_chain_dev.cuh"

#ifndef NMX_EPI_QUARTER
#define NMX_EPI_QUARTER 1  // forward epilogue granularity: 1 = one 64-column chunk per step (16 warps x 16 columns)
#endif
#include "nmx_gemm.cuh"

using namespace nmx;
using namespace nmx::sm100;
using namespace nmx::chain_dev;

namespace {

constexpr int kEpiWarps = 16;               // 4 per TMEM lane quadrant
constexpr int kEpiCols = 32;                // columns one epilogue warp handles per half-phase
// Warp roles.  The issue scheduler favours the highest warp id on a sub-partition, so the latency-critical single
// issuers (MMA, TMA) sit above the sixteen epilogue warps they share the SM with.
constexpr int kEpiWarp0 = 0;                // warps 0..15 epilogue (TMEM lane quadrant = warp % 4)
constexpr int kStoreWarp = 16;              // activation-store issuer (training) + TMEM alloc/dealloc
constexpr int kTmaWarp = 17;
constexpr int kMmaWarp = 18;
constexpr int kMaskWarp = 19;             // backward: ReLU sign-bit loader; forward: fused input encoder
constexpr int kThreads = 20 * 32;         // (inference: the idle store warp is the second input encoder)
constexpr int kStages = 3;
constexpr int kSlabBytes = 256 * 64 * 2;   // one piece of weights: up to 256 output rows x 64-wide K slab
constexpr int kChunkBytes = 128 * 64 * 2;  // one 128-row x 64-col bf16 activation chunk
constexpr int kBitsBytes = 128 * 32;       // ReLU sign bits of one 128-row x 256-col activation tile

// Shared-memory layout.  Forward: activation tile, resident encoded-input chunks (pos, dir), weight ring, biases and
// head weights (+ a 4 KB ReLU sign-bit staging tile aliased onto rows 1..4 of the head-7 weight block, which a
// view-dir net does not use).  Backward: activation tile (dY), two 4 KB sign-bit slots refilled by bulk copies one
// step ahead, weight ring, w_alpha / w_rgb.
template <int MODE>
struct SmemT {
  static constexpr int kActOff = 0;                                       // 4 chunks (128 x 256 bf16)
  static constexpr int kX0Off = kActOff + 4 * kChunkBytes;                // fwd: pos chunk, dir chunk; bwd: 2 sign-bit slots
  static constexpr int kRingOff = kX0Off + (MODE == 0 ? 2 * kChunkBytes : 2 * kBitsBytes);
  static constexpr int kBiasOff = kRingOff + kStages * kSlabBytes;        // fwd: [kMaxChainLayers][256] fp32
  static constexpr int kW7Off = kBiasOff + (MODE == 0 ? kMaxChainLayers * 256 * 4 : 0);   // [8][256] / bwd [1][256] fp32
  static constexpr int kWrgbOff = kW7Off + (MODE == 0 ? 8 : 1) * 256 * 4; // rgb weights [3][128] fp32
  static constexpr int kXchgOff = kWrgbOff + 3 * 128 * 4;                 // fwd: [3][128][4] fp32 partial sums
  static constexpr int kBarOff = kXchgOff + (MODE == 0 ? 3 * 128 * 4 * 4 : 0);
  // barriers: full[S], empty[S], tfull[2 stages][2 halves], tempty[2], act_ready[4], x0pos_full, x0pos_empty,
  //           x0dir_full, x0dir_empty, store_done, bits_full[2], bits_empty[2]
  static constexpr int kNumBars = 2 * kStages + 4 + 2 + 4 + 4 + 1 + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kAlloc = kTotal + 1024;
  static_assert(kAlloc <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
};
using Smem = SmemT<0>;

// NMX_CHAIN_DBG bit 2: CTA 0 records (clock64, globaltimer) at pipeline events of its first kTraceTiles tiles
constexpr int kTraceTiles = 6;
__device__ long long g_trace[2][kTraceTiles][kMaxChainLayers][4][2];
__device__ __forceinline__ void trace(bool on, int role, int it, int l, int ev) {
  if (on && it < kTraceTiles) {
    long long c = clock64();
    unsigned long long g;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    g_trace[role][it][l][ev][0] = c;
    g_trace[role][it][l][ev][1] = (long long)g;
  }
}

// Epilogue of one layer for one warp.  Half-phase h (accumulator columns [128h, 128h+128) = chunks 2h, 2h+1):
// this warp takes 32 columns of one of the two chunks: TMEM -> (+bias, ReLU) -> bf16 -> swizzled smem chunk (A operand
// of the next layer / TMA-store source), optional register heads.
// HEAD: 0 none, 1 alpha (one output), 2 rgb, 3 output_linear (up to 8 outputs).
template <bool RELU, int HEAD, bool BITS>
__device__ __forceinline__ void epi_layer(const uint32_t tacc, const int n_halves, uint64_t* tfull2, const uint32_t aphase,
                                          const int part, const uint32_t bias_addr, const uint32_t act_row_addr,
                                          const uint32_t swz, uint64_t* act_ready, const bool signal, const int lane,
                                          const uint32_t hw_addr, const int head_n, float (&hp)[8], float (&rgbp)[3],
                                          const bool skip_math, const bool tr_on, const int it, const int l,
                                          const uint32_t bits_row_addr) {
#if NMX_EPI_QUARTER
  // Quarter steps: all 16 epilogue warps work on ONE 64-column chunk at a time (16 columns per warp), so the first
  // K-slab of the next layer is released after a quarter of the epilogue instead of half of it.  The TMEM load of
  // step s+1 is issued as soon as the math of step s has consumed the registers, i.e. it is in flight during the
  // proxy fence and the barrier arrival of step s.
  const int nsteps = 2 * n_halves;
  (void)skip_math;  // debug switch of the half-step variant only
  uint32_t r[16];
  mbar_wait(&tfull2[0], aphase);
  trace(tr_on, 1, it, l, 0);
  tc_fence_after();
  tmem_ld_32x16(tacc + (uint32_t)(part * 16), r);
#pragma unroll 1
  for (int st = 0; st < nsteps; ++st) {
    tmem_ld_wait_regs<16>(r);
    if (st == 0) trace(tr_on, 1, it, l, 1);
    epi_cols<RELU, HEAD, 2, BITS>(r, st, st * 64 + part * 16, 2 * part, 8 * (part & 1), bias_addr, act_row_addr, swz, hw_addr,
                            head_n, hp, rgbp, bits_row_addr ? bits_row_addr + (uint32_t)(2 * st + (part >> 1)) * 4u : 0u);
    if (st == 1) {  // columns 128.. belong to the second commit (same instant for N = 256; keeps the phases in step)
      mbar_wait(&tfull2[1], aphase);
      tc_fence_after();
    }
    if (st + 1 < nsteps) tmem_ld_32x16(tacc + (uint32_t)((st + 1) * 64 + part * 16), r);
    fence_proxy_async_smem();
    __syncwarp();
    if (signal && lane == 0) mbar_arrive(&act_ready[st]);
  }
#else
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    mbar_wait(&tfull2[h], aphase);
    if (h == 0) trace(tr_on, 1, it, l, 0);
    if (h < n_halves) {
      tc_fence_after();
      const int c = 2 * h + (part >> 1);
      const int sub = part & 1;
      const int c0 = c * 64 + sub * 32;
      if (!skip_math) {
        uint32_t r[32];
        tmem_ld_32x32(tacc + (uint32_t)c0, r);
        tmem_ld_wait_regs<32>(r);
        if (h == 0) trace(tr_on, 1, it, l, 1);
        epi_cols<RELU, HEAD, 4, BITS>(r, c, c0, sub * 4, 0, bias_addr, act_row_addr, swz, hw_addr, head_n, hp, rgbp,
                                bits_row_addr ? bits_row_addr + (uint32_t)(4 * h + part) * 4u : 0u);
      }
      __syncwarp();
      if (signal && lane == 0) mbar_arrive(&act_ready[c]);
    }
  }
#endif
}

// backward epilogue of 32 columns: bf16((acc [+ d_sigma * w_alpha]) masked by the ReLU sign bits of the saved
// activation (`bits`: bit e / 16+e = columns 2e / 2e+1 of this thread's 32 columns)).
// NP4 / piece0 / bit0 as in epi_cols (4: 32 columns = a whole sign-bit word; 2: 16 columns = pairs bit0..bit0+7 of it).
template <int EPI, int NP4>
__device__ __forceinline__ void bwd_cols(const uint32_t (&r)[8 * NP4], const int c, const int c0, const int piece0,
                                         const int bit0, const uint32_t act_row_addr, const uint32_t swz,
                                         const uint32_t bits, const uint32_t wa_addr, const float ds) {
  const uint32_t so = act_row_addr + (uint32_t)c * kChunkBytes;
  const uint64_t ds2 = pack64(__float_as_uint(ds), __float_as_uint(ds));
#pragma unroll
  for (int p4 = 0; p4 < NP4; ++p4) {
    const uint32_t piece = (((uint32_t)(piece0 + p4)) ^ swz) << 4;
    float wa[8];
    if (EPI == 2) {
      const float4 a0 = lds128(wa_addr + (uint32_t)(c0 + p4 * 8) * 4u);
      const float4 a1 = lds128(wa_addr + (uint32_t)(c0 + p4 * 8 + 4) * 4u);
      wa[0] = a0.x; wa[1] = a0.y; wa[2] = a0.z; wa[3] = a0.w; wa[4] = a1.x; wa[5] = a1.y; wa[6] = a1.z; wa[7] = a1.w;
    }
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint64_t x = pack64(r[p4 * 8 + 2 * e], r[p4 * 8 + 2 * e + 1]);
      if (EPI == 2) x = fma_f32x2(ds2, pack64(__float_as_uint(wa[2 * e]), __float_as_uint(wa[2 * e + 1])), x);
      uint32_t v = cvt_bf16x2<false>(x);
      if (EPI >= 1) v &= ((bits >> (bit0 + p4 * 4 + e)) & 0x00010001u) * 0xFFFFu;
      pk[e] = v;
    }
    sts128(so + piece, pk[0], pk[1], pk[2], pk[3]);
  }
  if (NP4 == 4) fence_proxy_async_smem();  // quarter steps: the caller fences after issuing the next TMEM load
}

// MODE 0: forward (bias + ReLU epilogue, heads).  MODE 1: backward data-gradient chain (models/NeRF.py backward of
// 201-243): step A computes d_hd = (d_rgb W_rgb) * [hd > 0] on the CUDA cores, then every layer is
// dX = dY W (W^T copies as the K-major B operand), with the ReLU mask of the saved activation (and the alpha head's
// rank-1 term d_sigma (x) w_alpha on the feature layer) in the epilogue; every dY is TMA-stored for the wgrad kernels.
// SAVE (forward): the training variant keeps activations + ReLU sign bits; the inference variant compiles all of that
// out of the epilogue (the epilogue warps are issue-bound: every instruction per column counts).
template <int MODE, bool SAVE>
__global__ void __launch_bounds__(kThreads, 1)
mlp_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams prm) {
  using SL = SmemT<MODE>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SW128 operand tiles need 1024 B alignment; plain pointer arithmetic keeps the shared address space visible.
  // (The backward layout has no slack: there the pad must be 0, which the trap below enforces.)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_act = smem + SL::kActOff;
  uint8_t* s_x0 = smem + SL::kX0Off;
  uint8_t* s_ring = smem + SL::kRingOff;
  float* s_bias = reinterpret_cast<float*>(smem + SL::kBiasOff);
  float* s_w7 = reinterpret_cast<float*>(smem + SL::kW7Off);
  float* s_wrgb = reinterpret_cast<float*>(smem + SL::kWrgbOff);
  float* s_xchg = reinterpret_cast<float*>(smem + SL::kXchgOff);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SL::kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;   // [acc stage][half]
  uint64_t* tempty = tfull + 4;
  uint64_t* act_ready = tempty + 2;
  uint64_t* x0pos_full = act_ready + 4;
  uint64_t* x0pos_empty = x0pos_full + 1;
  uint64_t* x0dir_full = x0pos_empty + 1;
  uint64_t* x0dir_empty = x0dir_full + 1;
  uint64_t* store_done = x0dir_empty + 1;
  uint64_t* bits_full = store_done + 1;
  uint64_t* bits_empty = bits_full + 2;
  // forward: sign-bit staging tile (rows 1..4 of the head-7 weight block); backward: slot 0 of the two bit slots
  uint32_t* s_bits = MODE == 0 ? reinterpret_cast<uint32_t*>(s_w7 + 256) : reinterpret_cast<uint32_t*>(s_x0);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + SL::kTmemPtrOff);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int num_tiles = (prm.P + 127) / 128;
  const int NL = prm.n_layers;
  constexpr bool save = SAVE;

  if (warp == kTmaWarp && lane == 0) {
    for (int l = 0; l < NL; ++l) tma_prefetch_desc(&maps.w[l]);
    tma_prefetch_desc(&maps.x0);
    tma_prefetch_desc(&maps.save);
    tma_prefetch_desc(&maps.hd);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&tfull[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&tempty[i], kEpiWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&act_ready[i], NMX_EPI_QUARTER ? kEpiWarps : kEpiWarps / 2);
    mbar_init(x0pos_full, 1);
    mbar_init(x0pos_empty, 1);
    mbar_init(x0dir_full, 1);
    mbar_init(x0dir_empty, 1);
    mbar_init(store_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bits_full[i], 1);
      mbar_init(&bits_empty[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kStoreWarp) tmem_alloc<512>(tmem_ptr);
  // stage biases and the register-head weights (fp32) once per CTA
  if (MODE == 0) {
    for (int i = threadIdx.x; i < NL * 256; i += kThreads) {
      int l = i >> 8, c = i & 255;
      s_bias[i] = (c < prm.L[l].N) ? prm.params[prm.L[l].bias_off + c] : 0.0f;
    }
  }
  for (int i = threadIdx.x; i < prm.head7_n * 256; i += kThreads) s_w7[i] = prm.params[prm.head7_w_off + i];
  if (prm.rgb_layer >= 0 || MODE == 1)
    for (int i = threadIdx.x; i < 3 * 128; i += kThreads) s_wrgb[i] = prm.params[prm.rgb_w_off + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  // NOTE: the producer / MMA / store loops run warp-uniformly (all 32 lanes take the same path and poll the same
  // barriers); only the asynchronous issue itself is predicated on one elected lane.  That keeps descriptors and
  // addresses in uniform registers instead of a per-instruction register->uniform "waterfall".
  if (warp == kTmaWarp) {
    // ====================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      if (it == 0 && MODE == 0 && !prm.enc_fused) {
        if (elect_one()) {
          mbar_arrive_expect_tx(x0pos_full, kChunkBytes);
          tma_load_2d(s_x0, &maps.x0, x0pos_full, 0, tile * 128);
          if (prm.uses_dir) {
            mbar_arrive_expect_tx(x0dir_full, kChunkBytes);
            tma_load_2d(s_x0 + kChunkBytes, &maps.x0, x0dir_full, prm.x0_dir_col, tile * 128);
          }
        }
        __syncwarp();
      }
      for (int l = 0; l < NL; ++l) {
        const int np = prm.L[l].n_pieces;
        const int N = prm.L[l].N;
        for (int i = 0; i < np; ++i) {
          const uint32_t hi = (uint32_t)(prm.L[l].piece_tab[i] >> 32);  // weight box: k column
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = s_ring + stage * kSlabBytes;
          if (elect_one()) {
            if NMX_DBG(prm, 1) {  // experiment: no weight traffic (operands are whatever the ring holds)
              mbar_arrive(&full[stage]);
            } else {
              mbar_arrive_expect_tx(&full[stage], (uint32_t)N * 128u);
              tma_load_2d(dst, &maps.w[l], &full[stage], (int)(hi & 0xffffu), 0);
              if (N > 128) tma_load_2d(dst + 128 * 128, &maps.w[l], &full[stage], (int)(hi & 0xffffu), 128);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (MODE == 0 && !prm.enc_fused && l == 1 && it > 0 && prm.uses_dir) {
          // this tile's view-dir chunk: the previous tile's dir-layer MMAs must have finished reading the buffer
          mbar_wait(x0dir_empty, (uint32_t)((it - 1) & 1));
          if (elect_one()) {
            mbar_arrive_expect_tx(x0dir_full, kChunkBytes);
            tma_load_2d(s_x0 + kChunkBytes, &maps.x0, x0dir_full, prm.x0_dir_col, tile * 128);
          }
          __syncwarp();
        }
        if (MODE == 0 && !prm.enc_fused && l == prm.pos_prefetch_layer) {
          const int ntile = tile + gridDim.x;
          if (ntile < num_tiles) {  // next tile's position chunk, once this tile's last reader (skip layer) is done
            mbar_wait(x0pos_empty, (uint32_t)(it & 1));
            if (elect_one()) {
              mbar_arrive_expect_tx(x0pos_full, kChunkBytes);
              tma_load_2d(s_x0, &maps.x0, x0pos_full, 0, ntile * 128);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ====================================================== MMA issuer
    // Per piece: one 64-bit table entry from the constant bank (host-built, see launch_chain_fwd), at most two
    // barrier waits, four MMAs and their commits -- everything else was folded into the table.
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    uint32_t lcount = 0;
    uint32_t acbits = 0;  // bit c = parity of the number of signals so far on act_ready[c] (same replay in every role)
    const bool tr_on = NMX_DBG(prm, 4) && blockIdx.x == 0 && lane == 0;
    const uint32_t smem16 = smem_u32(smem) >> 4;
    const uint32_t ring16 = smem_u32(s_ring) >> 4;
    constexpr uint64_t kDescHi = (uint64_t)0x40004040u << 32;  // SBO 1024 B, descriptor version 1, SWIZZLE_128B
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      // The previous tile's last layer signalled act_ready for the store warp only (training) -- no MMA consumed it.
      // An mbarrier parity wait can only tell the current phase from the one before, so this warp must never get two
      // phases ahead of a barrier: observe those signals before waiting for this tile's first ones.
      if (it > 0 && save && !prm.L[NL - 1].feeds_next) {
        const int nck_last = prm.L[NL - 1].N / 64;
        for (int c = 0; c < nck_last; ++c) mbar_wait(&act_ready[c], ((acbits >> c) & 1u) ^ 1u);
      }
      if (MODE == 1) acbits ^= 0x3u;  // step A (d_hd) signals chunks 0,1 before the first layer
      for (int l = 0; l < NL; ++l, ++lcount) {
        const int as = lcount & 1;
        const uint32_t aphase = (lcount >> 1) & 1;
        const int np = prm.L[l].n_pieces;
        const uint32_t idesc = make_idesc_bf16(128, prm.L[l].N, 0, 0);
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        trace(tr_on, 0, it, l, 0);
        const uint32_t tacc = tmem_base + as * 256;
        for (int i = 0; i < np; ++i) {
          const uint64_t e = prm.L[l].piece_tab[i];
          const uint32_t lo = (uint32_t)e;
          mbar_wait(&full[stage], phase);
          if (i == 0) trace(tr_on, 0, it, l, 1);
          const uint32_t wc = (lo >> 25) & 7u;
          if (wc == 1) {  // first use of an activation chunk written by the previous layer's epilogue
            const uint32_t c = (lo >> 28) & 3u;
            mbar_wait(&act_ready[c], ((acbits >> c) & 1u) ^ 1u);
          } else if (wc == 2) {
            mbar_wait(x0pos_full, (uint32_t)(it & 1));
          } else if (wc == 3) {
            mbar_wait(x0dir_full, (uint32_t)(it & 1));
          }
          tc_fence_after();
          if (i == 0) trace(tr_on, 0, it, l, 2);
          const uint64_t adesc = kDescHi | (uint64_t)(((smem16 + (lo & 0xffffu)) & 0x3fffu) | 0x10000u);
          const uint64_t bdesc = kDescHi | (uint64_t)(((ring16 + (uint32_t)stage * (kSlabBytes >> 4)) & 0x3fffu) | 0x10000u);
          const uint32_t d_tmem = tacc + ((lo >> 16) & 0xffu);
          if (elect_one()) {
            umma_bf16(d_tmem, adesc, bdesc, idesc, (lo >> 24) & 1u);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
            umma_commit(&empty[stage]);
            if (lo & (1u << 30)) umma_commit(&tfull[as * 2 + 0]);
            if (lo & (1u << 31)) umma_commit(&tfull[as * 2 + 1]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (MODE == 0) {
          if (elect_one()) {
            if (l == prm.pos_last_layer) umma_commit(x0pos_empty);
            if (l == prm.dir_layer) umma_commit(x0dir_empty);
          }
          __syncwarp();
        }
        trace(tr_on, 0, it, l, 3);
        if (prm.L[l].feeds_next || save) acbits ^= (prm.L[l].N > 128 ? 0xFu : 0x3u);
      }
    }
  } else if (warp == kStoreWarp && (MODE == 1 || save || !prm.enc_fused)) {
    // ====================================================== activation-store issuer (training)
    if (save) {
      uint32_t acbits = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        if (MODE == 1) {  // step A: d_hd (chunks 0,1) -> its own [P, 128] buffer
          for (int c = 0; c < 2; ++c) {
            mbar_wait(&act_ready[c], (acbits >> c) & 1u);
            if (elect_one() && !NMX_DBG(prm, 16)) {
              tma_store_2d(&maps.hd, s_act + c * kChunkBytes, c * 64, tile * 128);
              tma_store_commit();
            }
            __syncwarp();
          }
          acbits ^= 0x3u;
          if (elect_one()) {
            tma_store_wait_read<0>();
            mbar_arrive(store_done);
          }
          __syncwarp();
        }
        for (int l = 0; l < NL; ++l) {
          const int nck = prm.L[l].N / 64;
          const int kind = prm.L[l].save_kind;
          const int row0 = prm.L[l].save_row0 + tile * 128;
          for (int c = 0; c < nck; ++c) {
            mbar_wait(&act_ready[c], (acbits >> c) & 1u);
            if (elect_one() && !NMX_DBG(prm, 16)) {
              if (kind == 1 && NMX_DBG(prm, 32))  // experiment: chunk-major planes, each store one contiguous 16 KB block
                tma_store_2d(&maps.save, s_act + c * kChunkBytes, 0, prm.L[l].save_row0 * 4 + c * prm.cap + tile * 128);
              else if (kind == 1) tma_store_2d(&maps.save, s_act + c * kChunkBytes, c * 64, row0);
              else if (kind == 2) tma_store_2d(&maps.hd, s_act + c * kChunkBytes, c * 64, tile * 128);
              tma_store_commit();
            }
            __syncwarp();
          }
          acbits ^= (nck == 4 ? 0xFu : 0x3u);
          if (MODE == 0 && prm.bits != nullptr && prm.L[l].bits_row0 >= 0) {
            if (elect_one() && !NMX_DBG(prm, 16)) {  // the layer's ReLU sign bits: one contiguous 4 KB tile
              bulk_store_1d(prm.bits + ((size_t)prm.L[l].bits_row0 + (size_t)tile * 128) * 8, s_bits, kBitsBytes);
              tma_store_commit();
            }
            __syncwarp();
          }
          if (elect_one()) {
            tma_store_wait_read<0>();  // the epilogue of the next layer may overwrite the chunks
            mbar_arrive(store_done);
          }
          __syncwarp();
        }
      }
      if (elect_one()) tma_store_wait<0>();
      __syncwarp();
    }
  } else if (warp == kMaskWarp || (MODE == 0 && warp == kStoreWarp && !save && prm.enc_fused)) {
    // ====================================================== ReLU sign-bit loader, backward only
    // Masked step s (step A, then every masked layer) reads slot s & 1: a contiguous 4 KB tile of sign bits, bulk-copied
    // one step ahead; a slot is refilled once all sixteen epilogue warps have arrived on bits_empty.
    if (MODE == 0 && prm.enc_fused) {
      // ---- fused input encoder (forward), one tile ahead of the MMA warp.  Training: this warp alone (four rows per
      // lane; the tile period is long enough).  Inference: the idle store warp takes rows 64..127.
      const int n_enc = save ? 1 : 2;
      const int ew = warp == kMaskWarp ? 0 : 1;
      const int rows_per_lane = 4 / n_enc;
      const bool leader = ew == 0 && lane == 0;  // arrives on the barriers and issues the x0 stores (training)
      auto enc_sync = [&]() {
        if (n_enc == 2) asm volatile("bar.sync 2, 64;" ::: "memory");
        else __syncwarp();
      };
      const uint32_t x0_addr = smem_u32(s_x0);
      const uint4* dpe = reinterpret_cast<const uint4*>(prm.dir_pe);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        // position chunk: its last reader of the previous tile is the skip layer
        if (it > 0) mbar_wait(x0pos_empty, (uint32_t)((it - 1) & 1));
        if (save && it > 0) {  // the x0 stores of the previous tile read these chunks
          if (leader) tma_store_wait_read<0>();
          enc_sync();
        }
#pragma unroll 1
        for (int rr = 0; rr < rows_per_lane; ++rr) {
          const int row_local = ew * 64 + rr * 32 + lane;
          const int row = tile * 128 + row_local;
          encode_pos_row(prm.rays, prm.ray_stride, prm.z, prm.p0 + row, prm.n_per_ray, row < prm.P,
                         x0_addr + (uint32_t)row_local * 128u, (uint32_t)(row_local & 7));
        }
        fence_proxy_async_smem();
        enc_sync();
        if (leader) {
          mbar_arrive(x0pos_full);
          if (save) {  // wgrad of the first / skip layer reads the encoded tile from HBM
            tma_store_2d(&maps.x0, s_x0, 0, tile * 128);
            tma_store_commit();
          }
        }
        if (prm.uses_dir) {
          // view-dir chunk: gathered from the per-ray table; its last reader of the previous tile is the dir layer
          if (it > 0) mbar_wait(x0dir_empty, (uint32_t)((it - 1) & 1));
#pragma unroll 1
          for (int rr = 0; rr < rows_per_lane; ++rr) {
            const int row_local = ew * 64 + rr * 32 + lane;
            const int row = tile * 128 + row_local;
            const uint32_t ra = x0_addr + kChunkBytes + (uint32_t)row_local * 128u;
            const uint32_t swz = (uint32_t)(row_local & 7);
            uint4 d[8];
            if (row < prm.P) {
              const long long b = (prm.p0 + row) / prm.n_per_ray - prm.b0;
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] = __ldg(dpe + b * 8 + j);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) d[j] = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sts128(ra + ((((uint32_t)j) ^ swz) << 4), d[j].x, d[j].y, d[j].z, d[j].w);
          }
          fence_proxy_async_smem();
          enc_sync();
          if (leader) {
            mbar_arrive(x0dir_full);
            if (save) {
              tma_store_2d(&maps.x0, s_x0 + kChunkBytes, prm.x0_dir_col, tile * 128);
              tma_store_commit();
            }
          }
        }
      }
      if (save && leader) tma_store_wait<0>();
      __syncwarp();
    }
    if (MODE == 1 && warp == kMaskWarp && !NMX_DBG(prm, 8)) {
      uint32_t step = 0;
      auto fill = [&](int row0) {
        const uint32_t slot = step & 1u;
        if (step >= 2) mbar_wait(&bits_empty[slot], ((step >> 1) - 1) & 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&bits_full[slot], kBitsBytes);
          bulk_load_1d(reinterpret_cast<uint8_t*>(s_bits) + slot * kBitsBytes, prm.bits + (size_t)row0 * 8, kBitsBytes,
                       &bits_full[slot]);
        }
        __syncwarp();
        ++step;
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        fill(prm.hd_bits_row0 + tile * 128);  // step A
        for (int l = 0; l < NL; ++l)
          if (prm.L[l].epi >= 1) fill(prm.L[l].bits_row0 + tile * 128);
      }
    }
  } else {
    // ====================================================== epilogue (warps 0..15)
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    const int part = (warp - kEpiWarp0) >> 2;        // 0..3: chunk (part >> 1) of the half, 32-column sub-block (part & 1)
    const int row_local = q * 32 + lane;
    const uint32_t swz = (uint32_t)(row_local & 7);
    const uint32_t act_row_addr = smem_u32(s_act) + (uint32_t)row_local * 128u;
    if constexpr (MODE == 0) {
      const uint32_t bias_base = smem_u32(s_bias);
      const uint32_t w7_addr = smem_u32(s_w7), wrgb_addr = smem_u32(s_wrgb);
      const bool skip_math = NMX_DBG(prm, 2) != 0;
      uint32_t lcount = 0;
      const bool tr_on = NMX_DBG(prm, 4) && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0;
      const bool vd = prm.rgb_layer >= 0;
      float hb[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // head biases of the view-dir net (rgb, alpha), loaded once
      if (vd && part == 0) {
        hb[0] = __ldg(prm.params + prm.rgb_b_off + 0);
        hb[1] = __ldg(prm.params + prm.rgb_b_off + 1);
        hb[2] = __ldg(prm.params + prm.rgb_b_off + 2);
        hb[3] = __ldg(prm.params + prm.head7_b_off);
      }
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int row = tile * 128 + row_local;
        float hp[8];  // head-7 partial dot products over this thread's columns
  #pragma unroll
        for (int o = 0; o < 8; ++o) hp[o] = 0.0f;
        float rgbp[3] = {0.0f, 0.0f, 0.0f};
        for (int l = 0; l < NL; ++l, ++lcount) {
          const int as = lcount & 1;
          const uint32_t aphase = (lcount >> 1) & 1;
          const int n_halves = prm.L[l].N / 128;
          const bool signal = prm.L[l].feeds_next || save;
          // the TMA stores of the previous layer's chunks must have finished reading shared memory
          if (save && lcount > 0) mbar_wait(store_done, (lcount - 1) & 1);
          trace(tr_on, 1, it, l, 3);
          const uint32_t tacc = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(q * 32) << 16);
          const uint32_t bias_addr = bias_base + (uint32_t)l * 1024u;
          // sign-bit staging row of this thread (training forward of a net whose backward is the fused chain)
          const uint32_t bra = (save && prm.bits != nullptr && prm.L[l].bits_row0 >= 0)
                                   ? smem_u32(s_bits) + (uint32_t)row_local * 32u : 0u;
          uint64_t* tf = &tfull[as * 2];
          if (l == prm.head7_layer && prm.head7_n == 1)
            epi_layer<true, 1, SAVE>(tacc, n_halves, tf, aphase, part, bias_addr, act_row_addr, swz, act_ready, signal, lane,
                               w7_addr, 1, hp, rgbp, skip_math, tr_on, it, l, bra);
          else if (l == prm.head7_layer)
            epi_layer<true, 3, SAVE>(tacc, n_halves, tf, aphase, part, bias_addr, act_row_addr, swz, act_ready, signal, lane,
                               w7_addr, prm.head7_n, hp, rgbp, skip_math, tr_on, it, l, bra);
          else if (l == prm.rgb_layer)
            epi_layer<true, 2, SAVE>(tacc, n_halves, tf, aphase, part, bias_addr, act_row_addr, swz, act_ready, signal, lane,
                               wrgb_addr, 0, hp, rgbp, skip_math, tr_on, it, l, bra);
          else if (prm.L[l].relu)
            epi_layer<true, 0, SAVE>(tacc, n_halves, tf, aphase, part, bias_addr, act_row_addr, swz, act_ready, signal, lane, 0,
                               0, hp, rgbp, skip_math, tr_on, it, l, bra);
          else
            epi_layer<false, 0, false>(tacc, n_halves, tf, aphase, part, bias_addr, act_row_addr, swz, act_ready, signal, lane, 0,
                                0, hp, rgbp, skip_math, tr_on, it, l, bra);
          trace(tr_on, 1, it, l, 2);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
        // ---- register heads: combine the column parts' partial sums (4 values per pass) and write the raw outputs
        const int nvals = vd ? 4 : prm.head7_n;
        float tot[8];
  #pragma unroll
        for (int o = 0; o < 8; ++o) tot[o] = 0.0f;
        for (int pass = 0; pass * 4 < nvals; ++pass) {
          float v4[4];
  #pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = 0.0f;
  #pragma unroll
            for (int o = 0; o < 8; ++o)
              if (o == pass * 4 + j) x = hp[o];
            v4[j] = x;
          }
          if (vd) { v4[0] = rgbp[0]; v4[1] = rgbp[1]; v4[2] = rgbp[2]; v4[3] = hp[0]; }
          if (part > 0)
            *reinterpret_cast<float4*>(s_xchg + ((part - 1) * 128 + row_local) * 4) = make_float4(v4[0], v4[1], v4[2], v4[3]);
          asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
          if (part == 0) {
  #pragma unroll
            for (int pp = 0; pp < 3; ++pp) {
              const float4 t = *reinterpret_cast<const float4*>(s_xchg + (pp * 128 + row_local) * 4);
              v4[0] += t.x; v4[1] += t.y; v4[2] += t.z; v4[3] += t.w;
            }
  #pragma unroll
            for (int j = 0; j < 4; ++j)
  #pragma unroll
              for (int o = 0; o < 8; ++o)
                if (o == pass * 4 + j) tot[o] = v4[j];
          }
          if ((pass + 1) * 4 < nvals) asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        }
        if (part == 0 && row < prm.P) {
          float* o_row = prm.out + (size_t)row * prm.out_cols;
          if (vd) {
            *reinterpret_cast<float4*>(o_row) = make_float4(tot[0] + hb[0], tot[1] + hb[1], tot[2] + hb[2], tot[3] + hb[3]);
          } else {
  #pragma unroll
            for (int o = 0; o < 8; ++o)
              if (o < prm.head7_n) o_row[o] = tot[o] + __ldg(prm.params + prm.head7_b_off + o);
          }
        }
      }
  
    } else {
      // ---------------------------------------------------- backward data-gradient epilogue
      const uint32_t wa_addr = smem_u32(s_w7);      // w_alpha [256] fp32
      const uint32_t wrgb_addr = smem_u32(s_wrgb);  // w_rgb [3][128] fp32
      const bool tr_on = NMX_DBG(prm, 4) && blockIdx.x == 0 && warp == kEpiWarp0 && lane == 0;
      const bool no_mask = NMX_DBG(prm, 8) != 0;  // experiment: skip the saved-activation reads
      uint32_t lcount = 0, scount = 0;
      uint32_t mstep = 0;  // masked steps so far (step A + masked layers): slot = mstep & 1, parity = (mstep >> 1) & 1
      const uint32_t bits_base = smem_u32(s_bits) + (uint32_t)row_local * 32u;
      auto ld_bits = [&](uint32_t slot, int w) -> uint32_t {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(bits_base + slot * kBitsBytes + (uint32_t)w * 4u) : "memory");
        return v;
      };
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int row = tile * 128 + row_local;
        const bool row_ok = row < prm.P;
        float4 dr = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (row_ok) dr = __ldg(reinterpret_cast<const float4*>(prm.d_out) + row);
        // ---- step A: d_hd = (d_rgb W_rgb) * [hd > 0]  -> chunks 0,1 (A operand of the dir-layer data gradient)
        // Quarter layout as in the layers below: every warp takes 16 columns of chunk 0, then 16 columns of chunk 1.
        if (scount > 0) mbar_wait(store_done, (scount - 1) & 1);
        {
          uint32_t bw[2] = {~0u, ~0u};
          if (!no_mask) {
            mbar_wait(&bits_full[mstep & 1u], (mstep >> 1) & 1u);
            bw[0] = ld_bits(mstep & 1u, part >> 1);
            bw[1] = ld_bits(mstep & 1u, 2 + (part >> 1));
          }
#pragma unroll
          for (int st = 0; st < 2; ++st) {
            const uint32_t so = act_row_addr + (uint32_t)st * kChunkBytes;
#pragma unroll
            for (int p4 = 0; p4 < 2; ++p4) {
              const int j = st * 64 + part * 16 + p4 * 8;
              const uint32_t piece = (((uint32_t)(2 * part + p4)) ^ swz) << 4;
              float v[8];
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const float4 w0 = lds128(wrgb_addr + (uint32_t)(0 * 128 + j + hh * 4) * 4u);
                const float4 w1 = lds128(wrgb_addr + (uint32_t)(1 * 128 + j + hh * 4) * 4u);
                const float4 w2 = lds128(wrgb_addr + (uint32_t)(2 * 128 + j + hh * 4) * 4u);
                v[hh * 4 + 0] = dr.x * w0.x + dr.y * w1.x + dr.z * w2.x;
                v[hh * 4 + 1] = dr.x * w0.y + dr.y * w1.y + dr.z * w2.y;
                v[hh * 4 + 2] = dr.x * w0.z + dr.y * w1.z + dr.z * w2.z;
                v[hh * 4 + 3] = dr.x * w0.w + dr.y * w1.w + dr.z * w2.w;
              }
              uint32_t pk[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                pk[e] = pack_bf16(v[2 * e], v[2 * e + 1]) &
                        (((bw[st] >> (8 * (part & 1) + p4 * 4 + e)) & 0x00010001u) * 0xFFFFu);
              sts128(so + piece, pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[st]);
          }
          if (!no_mask) {
            if (lane == 0) mbar_arrive(&bits_empty[mstep & 1u]);
            ++mstep;
          }
        }
        ++scount;
        for (int l = 0; l < NL; ++l, ++lcount, ++scount) {
          const int as = lcount & 1;
          const uint32_t aphase = (lcount >> 1) & 1;
          const int epi = no_mask ? (prm.L[l].epi == 2 ? 2 : 0) : prm.L[l].epi;
          const bool masked = prm.L[l].epi >= 1 && !no_mask;
          const uint32_t slot = mstep & 1u;
          const int nsteps = prm.L[l].N / 64;
          // the TMA stores of the previous step's chunks must have finished reading shared memory
          mbar_wait(store_done, (scount - 1) & 1);
          const uint32_t tacc = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(q * 32) << 16);
          uint32_t bw[4] = {0u, 0u, 0u, 0u};
          if (masked) {
            mbar_wait(&bits_full[slot], (mstep >> 1) & 1u);
#pragma unroll
            for (int st = 0; st < 4; ++st) bw[st] = ld_bits(slot, 2 * st + (part >> 1));
            ++mstep;
          }
          // ---- four quarter steps: one 64-column chunk per step, 16 columns per warp; the TMEM load of step s+1 is
          //      issued before the proxy fence of step s
          mbar_wait(&tfull[as * 2 + 0], aphase);
          trace(tr_on, 1, it, l, 0);
          tc_fence_after();
          uint32_t r[16];
          tmem_ld_32x16(tacc + (uint32_t)(part * 16), r);
#pragma unroll
          for (int st = 0; st < 4; ++st) {
            if (st < nsteps) {
              const int c0 = st * 64 + part * 16;
              tmem_ld_wait_regs<16>(r);
              if (epi == 2) bwd_cols<2, 2>(r, st, c0, 2 * part, 8 * (part & 1), act_row_addr, swz, bw[st], wa_addr, dr.w);
              else if (epi == 1) bwd_cols<1, 2>(r, st, c0, 2 * part, 8 * (part & 1), act_row_addr, swz, bw[st], wa_addr, 0.0f);
              else bwd_cols<0, 2>(r, st, c0, 2 * part, 8 * (part & 1), act_row_addr, swz, bw[st], wa_addr, 0.0f);
            }
            if (st == 1) {
              mbar_wait(&tfull[as * 2 + 1], aphase);
              tc_fence_after();
            }
            if (st < nsteps) {
              if (st + 1 < nsteps) tmem_ld_32x16(tacc + (uint32_t)((st + 1) * 64 + part * 16), r);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) mbar_arrive(&act_ready[st]);
            }
          }
          if (masked && lane == 0) mbar_arrive(&bits_empty[slot]);  // the slot can be refilled
          trace(tr_on, 1, it, l, 2);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kStoreWarp) tmem_dealloc<512>(tmem_base);
}

}  // namespace

namespace nmx {

template <int MODE>
static int launch_chain(const ChainMaps& maps, const ChainParams& prm_in, cudaStream_t stream) {
  if (prm_in.P <= 0) return 0;
  ChainParams prm = prm_in;
  for (int l = 0; l < prm.n_layers; ++l) {
    ChainLayerDesc& d = prm.L[l];
    uint32_t pk = 0;
    for (int s = 0; s < d.n_slabs; ++s) pk |= (uint32_t)(d.src[s] & 15) << (4 * s);
    d.src_packed = pk;
    // One piece per 64-wide K slab, full layer width (N = 256 MMAs: a tcgen05.mma costs ~80 clocks to issue whatever
    // its N, so narrower pieces are issue-bound -- measured with nmx_diag_mma_rate).
    // Table entry: [0,16) A-operand smem offset >> 4, [16,24) accumulator column, [24] accumulate flag of the first
    // MMA, [25,28) wait code (0 none, 1 act_ready[chunk in 28-29], 2 pos, 3 dir), [30] / [31] commit "half 0 / 1 of
    // the accumulator complete" after the piece, [32,48) weight k column.
    for (int s = 0; s < d.n_slabs; ++s) {
      const int src = d.src[s];
      const uint32_t a_off = src == kSrcPos ? Smem::kX0Off : src == kSrcDir ? Smem::kX0Off + kChunkBytes
                                                                             : Smem::kActOff + src * kChunkBytes;
      uint64_t e = (uint64_t)(a_off >> 4) | ((uint64_t)(s > 0 ? 1 : 0) << 24);
      const uint64_t wc = src == kSrcPos ? 2 : src == kSrcDir ? 3 : 1;
      e |= wc << 25;
      if (wc == 1) e |= (uint64_t)src << 28;
      e |= (uint64_t)(s * 64) << 32;
      if (s == d.n_slabs - 1) e |= (1ull << 30) | (1ull << 31);
      d.piece_tab[s] = e;
    }
    d.n_pieces = d.n_slabs;
  }
  static bool attr[64] = {};
  if (once_per_device(attr)) {
    NMX_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemT<MODE>::kAlloc));
    if (MODE == 0)
      NMX_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemT<0>::kAlloc));
  }
  int tiles = (prm.P + 127) / 128;
  const int cap_ctas = (prm.max_ctas > 0 && prm.max_ctas < kNumSMs) ? prm.max_ctas : kNumSMs;
  int grid = tiles < cap_ctas ? tiles : cap_ctas;
  double flops = 0.0;  // padded flops actually issued to the tensor pipe
  for (int l = 0; l < prm.n_layers; ++l) flops += 2.0 * prm.P * prm.L[l].N * 64.0 * prm.L[l].n_slabs;
  prof_begin(MODE == 1 ? 4 : (prm.save ? 3 : 2), flops, stream);
  if (MODE == 0 && !prm.save) mlp_chain_kernel<0, false><<<grid, kThreads, SmemT<0>::kAlloc, stream>>>(maps, prm);
  else mlp_chain_kernel<MODE, true><<<grid, kThreads, SmemT<MODE>::kAlloc, stream>>>(maps, prm);
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_chain_fwd(const ChainMaps& maps, const ChainParams& prm, cudaStream_t stream) {
  return launch_chain<0>(maps, prm, stream);
}
int launch_chain_bwd(const ChainMaps& maps, const ChainParams& prm, cudaStream_t stream) {
  return launch_chain<1>(maps, prm, stream);
}

}  // namespace nmx

// diagnostics: copies the event trace recorded under NMX_CHAIN_DBG bit 2 (n int64 values) to host memory
extern "C" int nmx_chain_trace_read(long long* out, int n) {
  NMX_CHECK_ARG(out && n > 0 && (size_t)n * 8 <= sizeof(g_trace), "out non-null; 0 < n <= trace size");
  NMX_CUDA(cudaMemcpyFromSymbol(out, g_trace, (size_t)n * 8));
  return 0;
}
