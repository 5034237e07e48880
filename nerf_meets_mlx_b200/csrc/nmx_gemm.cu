// K3 building blocks: bf16 GEMMs on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged
// by TMA into 128B-swizzled shared memory), warp-specialised and persistent.
//
//   gemm_kmajor_kernel : D[M,N] = epi(A[M,K] * B[N,K]^T)      -- layer forward and dgrad (B = W or W^T copy)
//       A may be the K-concatenation of two row-major bf16 tensors (skip connection / view-dir concat).
//       epi: +bias, (+ row_vec (x) col_vec), ReLU or ReLU-mask from a saved activation, -> bf16 (or fp32).
//   wgrad_kernel       : dW[M,N] += dY[P,M]^T * X[P,N]        -- contraction over points, MN-major operands read
//       straight from the row-major activations (no transposes), split over points, fp32 vector-atomic reduce.
//
// Tile: 128 points (UMMA M=128, cta_group::1) x N<=256 (full layer width, so activations are read once) x K=64 per
// stage.  Warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quadrant
// = warp_id % 4).  TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <mutex>
#include <vector>

#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_gemm.cuh"

using namespace nmx;
using namespace nmx::sm100;

namespace nmx {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return NMX_E_DRIVER;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p rows=%llu cols=%llu ld=%llu box=[%u,%u]", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return NMX_E_DRIVER;
  }
  return 0;
}

}  // namespace nmx

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;

struct GemmArgs {
  int a0_col, a0_k, a1_col, a1_k;  // K segments of A (elements; multiples of 64); a1_k may be 0
  int b_col;                       // first K column of B to use
  int M, N;
  const float* bias;      // [N] or null
  void* D;                // bf16 or fp32 [M, ldd]
  int ldd;
  int out_fp32;
  int relu;
  int use_mask;           // output *= (mask > 0), mask = saved post-ReLU activation read through tmMask
  int mask_col, d_col;    // first column of the mask / output tensors
  const float* row_vec;   // optional rank-1 term row_vec[m*row_stride] * col_vec[n] added before the mask
  int row_stride;
  const float* col_vec;
  int accum;              // fp32 output only: D += result (the tile owns its rows: plain read-modify-write)
};

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTileBytes = STAGES * kStageBytes;
  static constexpr int kChunkBytes = BM * 64 * 2;               // one 128-row x 64-col bf16 chunk (SW128 box)
  static constexpr int kOutOff = kTileBytes;                    // 2 output staging chunks (TMA store source)
  static constexpr int kMaskOff = kOutOff + 2 * kChunkBytes;    // 2 ReLU-mask chunks (TMA load destination)
  static constexpr int kBarOff = kMaskOff + 2 * kChunkBytes;    // full[S], empty[S], tfull[2], tempty[2], mfull[2]
  static constexpr int kTmemPtrOff = kBarOff + (2 * STAGES + 6) * 8;
  static constexpr int kBiasOff = kTmemPtrOff + 16;             // bias[BN] + colvec[BN]
  static constexpr int kTotal = kBiasOff + 2 * BN * 4;
  static constexpr int kAlloc = kTotal + 1024;                  // slack for manual 1024 B alignment
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kmajor_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                   const __grid_constant__ CUtensorMap tmMask, const GemmArgs args) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* mfull = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOff);
  float* s_colv = s_bias + BN;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = args.N;
  const int num_tiles = (args.M + BM - 1) / BM;
  const int num_kb = (args.a0_k + args.a1_k) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    tma_prefetch_desc(&tmMask);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
      mbar_init(&mfull[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarp0) tmem_alloc<512>(tmem_ptr);
  for (int i = threadIdx.x; i < BN; i += kThreads) {
    s_bias[i] = (args.bias != nullptr && i < N) ? args.bias[i] : 0.0f;
    s_colv[i] = (args.col_vec != nullptr && i < N) ? args.col_vec[i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = L::kABytes + (uint32_t)N * BK * 2;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::kStageBytes;
          uint8_t* sB = sA + L::kABytes;
          mbar_arrive_expect_tx(&full[stage], tx);
          int k = kb * BK;
          if (k < args.a0_k) tma_load_2d(sA, &tmA0, &full[stage], args.a0_col + k, tile * BM);
          else tma_load_2d(sA, &tmA1, &full[stage], args.a1_col + (k - args.a0_k), tile * BM);
          tma_load_2d(sB, &tmB, &full[stage], args.b_col + k, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(BM, N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + L::kABytes;
          const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // ================================ epilogue warps ================================
    // TMEM -> registers (+bias, rank-1 term, ReLU / ReLU-mask) -> bf16 -> 128B-swizzled smem chunk -> TMA store.
    // Each thread owns one tile row; a 16 B piece c of its 128 B chunk row lands at (c ^ (row & 7)) so that both the
    // st.shared (quarter-warp = 8 rows x 8 distinct pieces) and the TMA engine see conflict-free / canonical layout.
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_local = q * 32 + lane;
    const bool leader = (threadIdx.x == kEpiWarp0 * 32);
    uint8_t* s_out = smem + L::kOutOff;
    uint8_t* s_mask = smem + L::kMaskOff;
    const int nchunks = N / 64;
    const bool use_mask = args.use_mask != 0;
    const uint32_t swz = (uint32_t)(row_local & 7);
    uint32_t g = 0;  // running chunk counter (selects the staging buffer and the mask barrier parity)
    if (use_mask && leader && (int)blockIdx.x < num_tiles) {
      mbar_arrive_expect_tx(&mfull[0], L::kChunkBytes);
      tma_load_2d(s_mask, &tmMask, &mfull[0], args.mask_col, blockIdx.x * BM);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const int row = tile * BM + row_local;
      const bool row_ok = row < args.M;
      const float rv = (args.row_vec != nullptr && row_ok) ? args.row_vec[(size_t)row * args.row_stride] : 0.0f;
      if (args.out_fp32) {
        // test / debug path: direct fp32 stores
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + as * BN + c0 + ((uint32_t)(q * 32) << 16), r);
          tmem_ld_wait();
          if (row_ok) {
            float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(args.D) + (size_t)row * args.ldd + c0);
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float x = __uint_as_float(r[k4 * 4 + e]) + s_bias[c0 + k4 * 4 + e] + rv * s_colv[c0 + k4 * 4 + e];
                v[e] = args.relu ? fmaxf(x, 0.0f) : x;
              }
              if (args.accum) {
                const float4 o = dp[k4];
                v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
              }
              dp[k4] = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
        }
      } else {
        for (int j = 0; j < nchunks; ++j, ++g) {
          const int buf = g & 1;
          uint8_t* so = s_out + buf * L::kChunkBytes + row_local * 128;
          const uint8_t* sm = s_mask + buf * L::kChunkBytes + row_local * 128;
          if (leader) tma_store_wait_read<1>();  // the store that last used s_out[buf] has finished reading it
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (use_mask && leader) {  // prefetch the next mask chunk into the buffer everyone finished reading
            int nj = j + 1, ntile = tile;
            if (nj == nchunks) { nj = 0; ntile = tile + gridDim.x; }
            if (ntile < num_tiles) {
              mbar_arrive_expect_tx(&mfull[buf ^ 1], L::kChunkBytes);
              tma_load_2d(s_mask + (buf ^ 1) * L::kChunkBytes, &tmMask, &mfull[buf ^ 1], args.mask_col + nj * 64, ntile * BM);
            }
          }
          if (use_mask) mbar_wait(&mfull[buf], (g >> 1) & 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int c0 = j * 64 + h * 32;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + as * BN + c0 + ((uint32_t)(q * 32) << 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int p4 = 0; p4 < 4; ++p4) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int c = c0 + p4 * 8 + e;
                float x = __uint_as_float(r[p4 * 8 + e]) + s_bias[c] + rv * s_colv[c];
                v[e] = args.relu ? fmaxf(x, 0.0f) : x;
              }
              const uint32_t piece = ((uint32_t)(h * 4 + p4) ^ swz) << 4;
              if (use_mask) {
                uint4 m = *reinterpret_cast<const uint4*>(sm + piece);
                const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(&m);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (!(__bfloat162float(mb[e]) > 0.0f)) v[e] = 0.0f;
              }
              *reinterpret_cast<uint4*>(so + piece) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]),
                                                                 pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            }
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (leader) {
            tma_store_2d(&tmD, s_out + buf * L::kChunkBytes, args.d_col + j * 64, tile * BM);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
    if (leader) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarp0) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ wgrad
struct WgradArgs {
  int dy_col, x_col;  // first column of dY / X to use
  int P;              // points (contraction length)
  int M, N;           // M = 128 per CTA m-tile (grid.y tiles), N multiple of 64, <= 256
  float* dW;          // fp32 [M_total, ldw], accumulated with atomics at column offset w_col
  int ldw, w_col;
  int n_valid;        // only columns < n_valid are accumulated (padded K inputs)
  float* db;          // optional bias gradient [M_total]: db[m] += sum_p dY[p, m] (ones-column MMA), or null
  int kb_per_cta;     // 64-point blocks per CTA
  // optional second 64-column operand sharing the SAME dY tile (dY is then read once for both products):
  // dW2[m, w2_col + n] += sum_p dY[p, m] X2[p, x2_col + n], n < n_valid2
  int x2_col;
  float* dW2;
  int ldw2, w2_col, n_valid2;
};

template <int STAGES, bool DUAL>
struct WgradSmem {
  static constexpr int kABytes = 2 * 64 * 64 * 2;   // two 64(M) x 64(P) boxes
  static constexpr int kBBytes = 4 * 64 * 64 * 2;   // up to four 64(N) x 64(P) boxes
  static constexpr int kB2Bytes = DUAL ? 64 * 64 * 2 : 0;  // one more 64(N2) x 64(P) box of the second operand
  static constexpr int kStageBytes = kABytes + kBBytes + kB2Bytes;
  static constexpr int kOnesOff = STAGES * kStageBytes;          // constant all-ones 64(N) x 64(P) box for bias grads
  static constexpr int kBarOff = kOnesOff + 8192;
  static constexpr int kTmemPtrOff = kBarOff + (2 * STAGES + 1) * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kAlloc = kTotal + 1024;
};

template <int STAGES, bool DUAL>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
             const __grid_constant__ CUtensorMap tmX2, const WgradArgs args) {
  using L = WgradSmem<STAGES, DUAL>;
  constexpr uint32_t kCol2 = 256;                  // accumulator columns of the second operand (DUAL)
  constexpr uint32_t kColOnes = DUAL ? 320 : 256;  // accumulator columns of the ones-MMA (bias gradient)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::kTmemPtrOff);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = args.N;
  const int nb = N / 64;
  const int m_tile = blockIdx.y;
  const int total_kb = (args.P + 63) / 64;
  const int kb0 = blockIdx.x * args.kb_per_cta;
  const int kb1 = min(total_kb, kb0 + args.kb_per_cta);
  const int num_kb = max(kb1 - kb0, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    if (DUAL) tma_prefetch_desc(&tmX2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == kEpiWarp0) tmem_alloc<512>(tmem_ptr);
  {  // bf16 1.0 everywhere: as an MN-major B operand it makes column 256.. of the accumulator the column sums of dY
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + L::kOnesOff);
    for (int i = threadIdx.x; i < 8192 / 4; i += kThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const bool want_db = args.db != nullptr;
  if (num_kb == 0) {  // nothing to contribute (uniform across the CTA)
    __syncthreads();
    if (warp == kEpiWarp0) tmem_dealloc<512>(tmem_base);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = L::kABytes + (uint32_t)nb * 64 * 64 * 2 + L::kB2Bytes;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sA = smem + stage * L::kStageBytes;
        uint8_t* sB = sA + L::kABytes;
        mbar_arrive_expect_tx(&full[stage], tx);
        const int p = kb * 64;
        tma_load_2d(sA, &tmDY, &full[stage], args.dy_col + m_tile * 128, p);
        tma_load_2d(sA + 8192, &tmDY, &full[stage], args.dy_col + m_tile * 128 + 64, p);
        for (int j = 0; j < nb; ++j) tma_load_2d(sB + j * 8192, &tmX, &full[stage], args.x_col + j * 64, p);
        if (DUAL) tma_load_2d(sB + L::kBBytes, &tmX2, &full[stage], args.x2_col, p);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
      const uint32_t idesc_ones = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t idesc2 = make_idesc_bf16(128, 64, 1, 1);
      const uint64_t odesc = make_smem_desc(smem_u32(smem + L::kOnesOff), 8192, 1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
        const uint32_t b_addr = a_addr + L::kABytes;
        // MN-major SW128: LBO = 8192 B between 64-wide M/N atoms (separate TMA boxes), SBO = 1024 B per 8 points
        const uint64_t adesc = make_smem_desc(a_addr, 8192, 1024);
        const uint64_t bdesc = make_smem_desc(b_addr, 8192, 1024);
        const uint64_t b2desc = make_smem_desc(b_addr + L::kBBytes, 8192, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 points = 16 rows of 128 B = 2048 B -> +128 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
          if (DUAL) umma_bf16(tmem_base + kCol2, adesc + (uint64_t)(k * 128), b2desc + (uint64_t)(k * 128), idesc2, (kb | k) != 0);
          if (want_db) umma_bf16(tmem_base + kColOnes, adesc + (uint64_t)(k * 128), odesc, idesc_ones, (kb | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int m = m_tile * 128 + q * 32 + lane;
    // flush: one vector reduction (red.global.add.v4.f32, 16 B) per 4 columns when the row start is 16 B aligned --
    // 4x fewer L2 atomic operations than scalar adds; the scalar path handles odd leading dimensions (63 / 283 / 319)
    auto flush = [&](uint32_t tcol0, int ncols, float* dW, int ldw, int w_col, int n_valid) {
      float* wrow = dW + (size_t)m * ldw + w_col;
      const bool vec_ok = ((ldw & 3) == 0) && ((w_col & 3) == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tcol0 + c0 + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
        if (m < args.M) {
          if (vec_ok && c0 + 32 <= n_valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + c0 + j),
                           "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                           "f"(__uint_as_float(r[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < n_valid) atomicAdd(wrow + c0 + j, __uint_as_float(r[j]));
          }
        }
      }
    };
    flush(0, N, args.dW, args.ldw, args.w_col, args.n_valid);
    if (DUAL) flush(kCol2, 64, args.dW2, args.ldw2, args.w2_col, args.n_valid2);
    if (want_db) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + kColOnes + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      if (m < args.M) atomicAdd(args.db + m, __uint_as_float(r[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarp0) tmem_dealloc<512>(tmem_base);
}

// column sums of a bf16 matrix (bias gradients): out[n] += sum_p Y[p, col0 + n]
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ Y, int ld, int col0, int N, int64_t P, float* __restrict__ out) {
  // block: 256 threads = 8 row-groups x 32 column-pair lanes; each lane owns 2 adjacent columns per 64-col chunk
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  __shared__ float s[8][64];
  for (int c0 = 0; c0 < N; c0 += 64) {
    float a0 = 0.f, a1 = 0.f;
    int c = c0 + lane * 2;
    if (c < N) {
      for (int64_t p = (int64_t)blockIdx.x * 8 + rg; p < P; p += (int64_t)gridDim.x * 8) {
        __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(Y + p * ld + col0 + c);
        a0 += __bfloat162float(v.x);
        a1 += __bfloat162float(v.y);
      }
    }
    s[rg][lane * 2] = a0;
    s[rg][lane * 2 + 1] = a1;
    __syncthreads();
    if (threadIdx.x < 64 && c0 + threadIdx.x < N) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) t += s[g][threadIdx.x];
      atomicAdd(out + c0 + threadIdx.x, t);
    }
    __syncthreads();
  }
}

}  // namespace

namespace nmx {

// ---- optional live profiling (bench.py roofline): CUDA events around every GEMM / wgrad launch on its stream
struct ProfRec { cudaEvent_t a, b; int kind; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
void prof_begin(int kind, double flops, cudaStream_t s) {
  if (!g_prof_on) return;
  ProfRec r; r.kind = kind; r.flops = flops;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t s) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof.back().b, s);
}

// Host-side launchers (internal C++ API used by nmx_mlp.cu and the C ABI test hook).
int launch_gemm(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0) return 0;
  if (g.N % 64 != 0 || g.N < 64 || g.N > 256) { set_error("gemm: N must be a multiple of 64 in [64,256] (got %d)", g.N); return NMX_E_BADARG; }
  if (g.a0_k % 64 || g.a1_k % 64 || g.a0_k <= 0) { set_error("gemm: K segments must be positive multiples of 64"); return NMX_E_BADARG; }
  if (g.M > 0x7fffffff - 256) { set_error("gemm: M too large"); return NMX_E_BADARG; }
  CUtensorMap tA0, tA1, tB;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tA0, g.A0, g.a0_rows, g.a0_cols, g.a0_ld, BM))) return rc;
  if (g.A1 != nullptr && g.a1_k > 0) {
    if ((rc = make_tmap_bf16_2d(&tA1, g.A1, g.a0_rows, g.a1_cols, g.a1_ld, BM))) return rc;
  } else {
    tA1 = tA0;
  }
  if ((rc = make_tmap_bf16_2d(&tB, g.B, g.b_rows, g.b_cols, g.b_ld, (uint32_t)g.N))) return rc;
  CUtensorMap tD, tM;
  if (!g.out_fp32) {
    if ((rc = make_tmap_bf16_2d(&tD, g.D, g.M, g.d_cols > 0 ? g.d_cols : g.N, g.ldd, BM))) return rc;
  } else {
    tD = tA0;
  }
  if (g.mask != nullptr) {
    if (g.out_fp32) { set_error("gemm: the ReLU-mask epilogue needs bf16 output"); return NMX_E_BADARG; }
    if ((rc = make_tmap_bf16_2d(&tM, g.mask, g.M, g.mask_cols > 0 ? g.mask_cols : g.N, g.ldmask, BM))) return rc;
  } else {
    tM = tA0;
  }
  GemmArgs a;
  a.a0_col = g.a0_col; a.a0_k = g.a0_k; a.a1_col = g.a1_col; a.a1_k = (g.A1 ? g.a1_k : 0);
  a.b_col = g.b_col; a.M = (int)g.M; a.N = g.N; a.bias = g.bias; a.D = g.D; a.ldd = g.ldd; a.out_fp32 = g.out_fp32;
  a.relu = g.relu; a.use_mask = g.mask != nullptr; a.mask_col = g.mask_col; a.d_col = g.d_col; a.row_vec = g.row_vec;
  a.row_stride = g.row_stride; a.col_vec = g.col_vec; a.accum = g.accum;
  int tiles = (int)((g.M + BM - 1) / BM);
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  prof_begin(0, 2.0 * (double)g.M * g.N * (g.a0_k + a.a1_k), stream);
  if (g.N > 128) {
    using L = GemmSmem<256, 3>;
    static bool attr[64] = {};
    if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc)); }
    gemm_kmajor_kernel<256, 3><<<grid, kThreads, L::kAlloc, stream>>>(tA0, tA1, tB, tD, tM, a);
  } else {
    using L = GemmSmem<128, 4>;
    static bool attr[64] = {};
    if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(gemm_kmajor_kernel<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc)); }
    gemm_kmajor_kernel<128, 4><<<grid, kThreads, L::kAlloc, stream>>>(tA0, tA1, tB, tD, tM, a);
  }
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad(const WgradDesc& g, cudaStream_t stream) {
  if (g.P <= 0) return 0;
  if (g.M % 64 || g.M <= 0 || g.N % 64 || g.N > 256 || g.N <= 0) { set_error("wgrad: M %% 64, N %% 64, N <= 256 required (M=%d N=%d)", g.M, g.N); return NMX_E_BADARG; }
  const bool dual = g.X2 != nullptr;
  if (dual && (g.dW2 == nullptr || g.n_valid2 <= 0 || g.n_valid2 > 64)) { set_error("wgrad: second operand needs dW2 and 0 < n_valid2 <= 64"); return NMX_E_BADARG; }
  CUtensorMap tDY, tX, tX2;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tDY, g.dY, g.P, g.dy_cols, g.dy_ld, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tX, g.X, g.P, g.x_cols, g.x_ld, 64))) return rc;
  if (dual) {
    if ((rc = make_tmap_bf16_2d(&tX2, g.X2, g.P, g.x2_cols, g.x2_ld, 64))) return rc;
  } else {
    tX2 = tX;
  }
  WgradArgs a;
  a.dy_col = g.dy_col; a.x_col = g.x_col; a.P = (int)g.P; a.M = g.M; a.N = g.N; a.dW = g.dW; a.ldw = g.ldw; a.w_col = g.w_col;
  a.n_valid = g.n_valid > 0 ? g.n_valid : g.N;
  a.db = g.db;
  a.x2_col = g.x2_col; a.dW2 = g.dW2; a.ldw2 = g.ldw2; a.w2_col = g.w2_col; a.n_valid2 = g.n_valid2;
  int m_tiles = (g.M + 127) / 128;
  int total_kb = (int)((g.P + 63) / 64);
  const int cta_cap = (g.max_ctas > 0 && g.max_ctas < kNumSMs) ? g.max_ctas : kNumSMs;
  int splits = cta_cap / m_tiles;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  a.kb_per_cta = (total_kb + splits - 1) / splits;
  splits = (total_kb + a.kb_per_cta - 1) / a.kb_per_cta;
  prof_begin(1, 2.0 * (double)g.P * g.M * (g.N + (dual ? 64 : 0)), stream);
  if (dual) {
    using L = WgradSmem<3, true>;
    static bool attr[64] = {};
    if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(wgrad_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc)); }
    wgrad_kernel<3, true><<<dim3(splits, m_tiles), kThreads, L::kAlloc, stream>>>(tDY, tX, tX2, a);
  } else {
    using L = WgradSmem<4, false>;
    static bool attr[64] = {};
    if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(wgrad_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kAlloc)); }
    wgrad_kernel<4, false><<<dim3(splits, m_tiles), kThreads, L::kAlloc, stream>>>(tDY, tX, tX2, a);
  }
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

int launch_colsum(const void* Y, int ld, int col0, int N, int64_t P, float* out, cudaStream_t stream) {
  if (P <= 0 || N <= 0) return 0;
  int64_t blocks = (P + 63) / 64;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  colsum_bf16_kernel<<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)Y, ld, col0, N, P, out);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmx

extern "C" int nmx_profile_enable(int on) {
  for (auto& r : nmx::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  nmx::g_prof.clear();
  nmx::g_prof_on = on != 0;
  return 0;
}

extern "C" int nmx_profile_report(int kind, double* total_ms, double* total_flops, int64_t* launches) {
  NMX_CUDA(cudaDeviceSynchronize());
  double ms = 0, fl = 0; int64_t n = 0;
  for (auto& r : nmx::g_prof) {
    if (r.kind != kind) continue;
    float t = 0; cudaEventElapsedTime(&t, r.a, r.b);
    ms += t; fl += r.flops; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = n;
  return 0;
}

// C ABI test hooks -----------------------------------------------------------------------------------------------
extern "C" int nmx_gemm_bf16(const void* A, const void* Bm, const float* bias, void* D, int64_t M, int N, int K,
                             int relu, int d_is_fp32, void* stream) {
  NMX_CHECK_ARG(A && Bm && D && M >= 0 && K > 0, "A, B, D non-null; M >= 0; K > 0");
  nmx::GemmDesc g{};
  g.A0 = A; g.a0_rows = M; g.a0_cols = K; g.a0_ld = K; g.a0_col = 0; g.a0_k = K;
  g.A1 = nullptr; g.a1_k = 0;
  g.B = Bm; g.b_rows = N; g.b_cols = K; g.b_ld = K; g.b_col = 0;
  g.M = M; g.N = N; g.bias = bias; g.D = D; g.ldd = N; g.out_fp32 = d_is_fp32; g.relu = relu;
  return nmx::launch_gemm(g, (cudaStream_t)stream);
}

extern "C" int nmx_wgrad_bf16(const void* dY, const void* X, float* dW, float* db, int64_t P, int M, int N, void* stream) {
  NMX_CHECK_ARG(dY && X && dW && P >= 0, "dY, X, dW non-null; P >= 0");
  nmx::WgradDesc g{};
  g.dY = dY; g.dy_cols = M; g.dy_ld = M; g.dy_col = 0;
  g.X = X; g.x_cols = N; g.x_ld = N; g.x_col = 0;
  g.P = P; g.M = M; g.N = N; g.dW = dW; g.ldw = N; g.w_col = 0; g.n_valid = N; g.db = db;
  return nmx::launch_wgrad(g, (cudaStream_t)stream);
}

extern "C" int nmx_colsum_bf16(const void* Y, float* out, int64_t P, int N, void* stream) {
  NMX_CHECK_ARG(Y && out && P >= 0 && N % 2 == 0, "Y, out non-null; N even");
  return nmx::launch_colsum(Y, N, 0, N, P, out, (cudaStream_t)stream);
}
