// CTA-pair (tcgen05 cta_group::2) GEMM building block: D[M, 256] (fp32) = A[M, K] * B[256, K]^T, bf16 operands.
//
// Two CTAs of a cluster (one TPC) work on ONE 256-row tile: each CTA stages its own 128 rows of A and its own HALF of
// the weight slab (128 of the 256 output columns) per 64-wide K step -- so the weight ring per SM is half as large as
// with cta_group::1 -- and the leader CTA issues M = 256 MMAs on behalf of both; every CTA finds the accumulator rows of
// ITS 128 rows in its own TMEM.  This is the mechanism the two-tile ping-pong MLP chain of DESIGN.md (next step) needs;
// it is exposed as a test hook (nmx_gemm_pair_bf16) and validated against torch in tests/test_gemm_gpu.py.
//
// Protocol (same barrier offsets in both CTAs' shared memory):
//   full[s]   (leader's copy is used): 1 arrival (leader's expect_tx) + the bytes of BOTH CTAs' TMA loads; CTA 1 issues
//             its loads with the barrier address mapped onto the leader's copy (peer bit cleared);
//   empty[s]  (each CTA's own copy): signalled in both CTAs by the leader's multicast tcgen05.commit;
//   tfull[a]  (each CTA's own copy): multicast commit after the tile's last K step;
//   tempty[a] (leader's copy): 8 arrivals = 4 epilogue warps x 2 CTAs (CTA 1 arrives remotely through mapa).
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include <cstdlib>

using namespace nmx;
using namespace nmx::sm100;

namespace {

constexpr int kStages = 4;
constexpr int kThreads = 192;            // warp 0 TMA, warp 1 MMA (leader only), warps 2..5 epilogue
constexpr int kStageBytes = 2 * 16384;   // A [128 x 64] + B half [128 x 64] bf16
struct PairArgs {
  int M, K;
  float* D;
  int ldd;
  int b_only;    // experiment (NMX_PAIR_BONLY): only the weight halves are streamed (A stays whatever is in smem)
  int no_store;  // experiment (NMX_PAIR_NOSTORE): skip the fp32 output stores to time the TMA + MMA pipeline alone
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const PairArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;   // [2]
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_pairs = gridDim.x >> 1, pair0 = blockIdx.x >> 1;
  const int num_tiles = (args.M + 255) / 256;
  const int num_kb = args.K / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair0; t < num_tiles; t += num_pairs) {
        const int row0 = t * 256 + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sA = smem + stage * kStageBytes;
          if (args.b_only) {
            if (leader) mbar_arrive_expect_tx(&full[stage], kStageBytes);
          } else {
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * kStageBytes);  // both CTAs' A rows + B halves
            tma_load_2d_pair(sA, &tmA, &full[stage], kb * 64, row0);
          }
          tma_load_2d_pair(sA + 16384, &tmB, &full[stage], kb * 64, (int)rank * 128);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(256, 256, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tcount = 0;
      for (int t = pair0; t < num_tiles; t += num_pairs, ++tcount) {
        const int as = tcount & 1;
        mbar_wait(&tempty[as], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
          const uint64_t bdesc = make_smem_desc(a_addr + 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          umma_commit_pair(&empty[stage]);
          if (kb == num_kb - 1) umma_commit_pair(&tfull[as]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    const int q = warp & 3;  // TMEM lane quadrant of this warp
    const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty[0]), 0);
    uint32_t tcount = 0;
    for (int t = pair0; t < num_tiles; t += num_pairs, ++tcount) {
      const int as = tcount & 1;
      mbar_wait(&tfull[as], (tcount >> 1) & 1);
      tc_fence_after();
      const int row = t * 256 + (int)rank * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < 256; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + as * 256 + c0 + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
        if (row < args.M && !(args.no_store && r[0] != 0x7fc00001u)) {
          float4* dp = reinterpret_cast<float4*>(args.D + (size_t)row * args.ldd + c0);
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4)
            dp[k4] = make_float4(__uint_as_float(r[k4 * 4]), __uint_as_float(r[k4 * 4 + 1]), __uint_as_float(r[k4 * 4 + 2]),
                                 __uint_as_float(r[k4 * 4 + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_leader0 + (uint32_t)as * 8u);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace

extern "C" int nmx_gemm_pair_bf16(const void* A, const void* Bm, float* D, int64_t M, int K, int max_pairs, void* stream) {
  NMX_CHECK_ARG(A && Bm && D && M > 0 && K > 0 && K % 64 == 0, "A, B, D non-null; M > 0; K a positive multiple of 64");
  NMX_CHECK_ARG(M < (1ll << 31) - 512, "M too large");
  CUtensorMap tA, tB;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tA, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tB, Bm, 256, (uint64_t)K, (uint64_t)K, 128))) return rc;
  PairArgs a;
  a.M = (int)M; a.K = K; a.D = D; a.ldd = 256;
  a.no_store = experiment_env("NMX_PAIR_NOSTORE") ? 1 : 0;
  a.b_only = experiment_env("NMX_PAIR_BONLY") ? 1 : 0;
  const int smem = kStages * kStageBytes + 256 + 1024;
  static bool attr[64] = {};
  if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); }
  int tiles = (int)((M + 255) / 256);
  int pairs = kNumSMs / 2;
  if (max_pairs > 0 && max_pairs < pairs) pairs = max_pairs;
  if (tiles < pairs) pairs = tiles;
  pair_gemm_kernel<<<pairs * 2, kThreads, smem, (cudaStream_t)stream>>>(tA, tB, a);
  NMX_LAUNCH_CHECK();
  return 0;
}
