// K3 weight gradients of a whole backward pass in ONE launch.
//
// dW_l[out, in] += dY_l[P, out]^T X_l[P, in] for every Linear of the net (models/NeRF.py:182-199) is a contraction over
// POINTS: HBM-read bound (1 KB/point/layer), tiny output.  One launch per layer (nmx_gemm.cu `wgrad_kernel`) splits
// the points over all 148 SMs, so every launch ends with 148 partial 128 x 256 fp32 tiles flushed by atomics (38 MB per
// launch, 18 launches per pass) plus a launch tail -- a fixed cost of ~11 us per launch that dominates a 1024-ray shard
// (strong scaling, SURVEY 7 "8-GPU >= 7x").  Here the SMs are divided among the LAYERS instead: job j gets a share of
// the CTAs proportional to the bytes it reads, each of its CTAs contracts a contiguous range of 64-point blocks with
// fp32 accumulators in TMEM (tcgen05.mma, MN-major operands straight from the row-major activations, as wgrad_kernel)
// and flushes ONCE: 148 partial tiles per PASS, one launch, and all layers' reads in flight together.
// Per CTA: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 flush (TMEM lane quadrant = warp % 4).
#include <algorithm>
#include <vector>

#include "nmx_common.cuh"
#include "nmx_sm100.cuh"
#include "nmx_gemm.cuh"

using namespace nmx;
using namespace nmx::sm100;

namespace {

constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int kStages = 4;
constexpr int kABytes = 2 * 64 * 64 * 2;   // two 64(M) x 64(P) boxes of dY
constexpr int kBBytes = 4 * 64 * 64 * 2;   // up to four 64(N) x 64(P) boxes of X
constexpr int kB2Bytes = 64 * 64 * 2;      // one 64(N2) x 64(P) box of the second operand
// Four stages in flight (148 SMs x ~170 KB = 25 MB: ~4 us of HBM latency at full bandwidth; three stages measured 25 %
// slower).  A job without a second operand uses 48 KB stages; with one, 56 KB stages -- 224 KB, which fits only because
// the all-ones operand of the bias-gradient MMA needs just two SW128 atoms (16 points x 128 B).
constexpr int kStagePlain = kABytes + kBBytes;
constexpr int kStageDual = kStagePlain + kB2Bytes;
constexpr int kOnesOff = kStages * kStageDual;
constexpr int kOnesBytes = 2048;
constexpr int kBarOff = kOnesOff + kOnesBytes;
constexpr int kTmemPtrOff = kBarOff + (2 * kStages + 1) * 8;
constexpr int kAlloc = kTmemPtrOff + 16;
static_assert(kAlloc <= 232448, "exceeds the 227 KB shared-memory limit of sm_100");
constexpr uint32_t kCol2 = 256, kColOnes = 320;

struct JobArgs {
  int dy_t, x_t, x2_t;
  int dy_row0, x_row0, x2_row0;
  int dy_col, x_col, x2_col;
  int M, N;
  int m_tiles, cta0, n_ctas, kb_per_cta;
  float* dW; int ldw, w_col, n_valid;
  float* db;
  float* dW2; int ldw2, w2_col, n_valid2;
};
struct BatchArgs {
  int n_jobs, total_kb;
  JobArgs job[kMaxWgradJobs];
};
struct BatchMaps { CUtensorMap t[kMaxWgradTensors]; };

__global__ void __launch_bounds__(kThreads, 1)
wgrad_batch_kernel(const __grid_constant__ BatchMaps maps, const __grid_constant__ BatchArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();  // SW128 operand tiles need 1024 B alignment; the layout has no slack
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kTmemPtrOff);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // which job / M tile / point range this CTA owns (uniform across the CTA)
  int j = 0;
  while (j + 1 < args.n_jobs && (int)blockIdx.x >= args.job[j + 1].cta0) ++j;
  const JobArgs& jb = args.job[j];
  const int local = (int)blockIdx.x - jb.cta0;
  const int m_tile = local % jb.m_tiles;
  const int split = local / jb.m_tiles;
  const int kb0 = split * jb.kb_per_cta;
  const int kb1 = min(args.total_kb, kb0 + jb.kb_per_cta);
  const int num_kb = max(kb1 - kb0, 0);
  const int N = jb.N, nb = N / 64;
  const bool has2 = jb.x2_t >= 0;
  const int stage_bytes = has2 ? kStageDual : kStagePlain;
  const bool want_db = jb.db != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.t[jb.dy_t]);
    tma_prefetch_desc(&maps.t[jb.x_t]);
    if (has2) tma_prefetch_desc(&maps.t[jb.x2_t]);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == kEpiWarp0) tmem_alloc<512>(tmem_ptr);
  {  // bf16 1.0 everywhere: as an MN-major B operand it makes the accumulator column the column sums of dY (db)
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kOnesOff);
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += kThreads) ones[i] = 0x3F803F80u;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (num_kb == 0) {
    __syncthreads();
    if (warp == kEpiWarp0) tmem_dealloc<512>(tmem_base);
    return;
  }

  // The producer and MMA loops run warp-uniformly (all lanes poll the same barriers); only the asynchronous issue itself
  // is predicated on one elected lane, which keeps descriptors in uniform registers (same pattern as nmx_chain.cu).
  if (warp == 0) {
    const CUtensorMap* mdy = &maps.t[jb.dy_t];
    const CUtensorMap* mx = &maps.t[jb.x_t];
    const CUtensorMap* mx2 = &maps.t[has2 ? jb.x2_t : jb.x_t];
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = kABytes + (uint32_t)nb * 8192u + (has2 ? kB2Bytes : 0);
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* sA = smem + stage * stage_bytes;
      uint8_t* sB = sA + kABytes;
      const int p = kb * 64;
      if (elect_one()) {
        mbar_arrive_expect_tx(&full[stage], tx);
        tma_load_2d(sA, mdy, &full[stage], jb.dy_col + m_tile * 128, jb.dy_row0 + p);
        tma_load_2d(sA + 8192, mdy, &full[stage], jb.dy_col + m_tile * 128 + 64, jb.dy_row0 + p);
        for (int c = 0; c < nb; ++c) tma_load_2d(sB + c * 8192, mx, &full[stage], jb.x_col + c * 64, jb.x_row0 + p);
        if (has2) tma_load_2d(sB + kBBytes, mx2, &full[stage], jb.x2_col, jb.x2_row0 + p);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 1, 1);
    const uint32_t idesc_ones = make_idesc_bf16(128, 16, 1, 1);
    const uint32_t idesc2 = make_idesc_bf16(128, 64, 1, 1);
    const uint64_t odesc = make_smem_desc(smem_u32(smem + kOnesOff), 8192, 1024);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
      const uint32_t b_addr = a_addr + kABytes;
      // MN-major SW128: LBO = 8192 B between 64-wide M/N atoms (separate TMA boxes), SBO = 1024 B per 8 points
      const uint64_t adesc = make_smem_desc(a_addr, 8192, 1024);
      const uint64_t bdesc = make_smem_desc(b_addr, 8192, 1024);
      const uint64_t b2desc = make_smem_desc(b_addr + kBBytes, 8192, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 16 points = 2048 B -> +128 in the (addr >> 4) field
          const uint32_t acc = (kb | k) != 0;
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, acc);
          if (has2) umma_bf16(tmem_base + kCol2, adesc + (uint64_t)(k * 128), b2desc + (uint64_t)(k * 128), idesc2, acc);
          if (want_db) umma_bf16(tmem_base + kColOnes, adesc + (uint64_t)(k * 128), odesc, idesc_ones, acc);
        }
        umma_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(tfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int m = m_tile * 128 + q * 32 + lane;
    // flush: red.global.add.v4.f32 per 4 columns when the row start is 16 B aligned, scalar atomics otherwise (odd
    // leading dimensions 63 / 283 / 319 of the reference's first, dir and skip layers)
    auto flush = [&](uint32_t tcol0, int ncols, float* dW, int ldw, int w_col, int n_valid) {
      float* wrow = dW + (size_t)m * ldw + w_col;
      const bool vec_ok = ((ldw & 3) == 0) && ((w_col & 3) == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0);
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + tcol0 + c0 + ((uint32_t)(q * 32) << 16), r);
        tmem_ld_wait();
        if (m < jb.M) {
          if (vec_ok && c0 + 32 <= n_valid) {
#pragma unroll
            for (int c = 0; c < 32; c += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + c0 + c),
                           "f"(__uint_as_float(r[c])), "f"(__uint_as_float(r[c + 1])), "f"(__uint_as_float(r[c + 2])),
                           "f"(__uint_as_float(r[c + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c0 + c < n_valid) atomicAdd(wrow + c0 + c, __uint_as_float(r[c]));
          }
        }
      }
    };
    flush(0, N, jb.dW, jb.ldw, jb.w_col, jb.n_valid);
    if (has2) flush(kCol2, 64, jb.dW2, jb.ldw2, jb.w2_col, jb.n_valid2);
    if (want_db) {
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + kColOnes + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      if (m < jb.M) atomicAdd(jb.db + m, __uint_as_float(r[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarp0) tmem_dealloc<512>(tmem_base);
}

}  // namespace

namespace nmx {

int launch_wgrad_batch(const WgradBatchDesc& d, cudaStream_t stream) {
  if (d.P <= 0 || d.n_jobs <= 0) return 0;
  if (d.n_jobs > kMaxWgradJobs || d.n_tensors < 1 || d.n_tensors > kMaxWgradTensors) { set_error("wgrad_batch: 1..%d jobs, 1..%d tensors", kMaxWgradJobs, kMaxWgradTensors); return NMX_E_BADARG; }
  BatchMaps maps;
  int rc;
  for (int t = 0; t < kMaxWgradTensors; ++t) {
    const WgradBatchTensor& T = d.t[t < d.n_tensors ? t : 0];
    if (t >= d.n_tensors) { maps.t[t] = maps.t[0]; continue; }
    if (T.rows > 0x7fffffff) { set_error("wgrad_batch: tensor rows exceed the TMA coordinate range"); return NMX_E_BADARG; }
    if ((rc = make_tmap_bf16_2d(&maps.t[t], T.base, (uint64_t)T.rows, (uint64_t)T.cols, (uint64_t)T.cols, 64))) return rc;
  }
  BatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n_jobs = d.n_jobs;
  a.total_kb = (int)((d.P + 63) / 64);
  // CTA shares proportional to the bytes a job reads per 64-point block (all jobs are HBM-read bound)
  std::vector<double> w(d.n_jobs);
  double wsum = 0.0;
  for (int j = 0; j < d.n_jobs; ++j) {
    const WgradBatchJob& g = d.job[j];
    if (g.M % 64 || g.M <= 0 || g.M > 256 || g.N % 64 || g.N <= 0 || g.N > 256) { set_error("wgrad_batch: job %d: M, N multiples of 64 in [64, 256]", j); return NMX_E_BADARG; }
    if (g.x2_t >= 0 && (g.dW2 == nullptr || g.n_valid2 <= 0 || g.n_valid2 > 64)) { set_error("wgrad_batch: job %d: second operand needs dW2, 0 < n_valid2 <= 64", j); return NMX_E_BADARG; }
    JobArgs& o = a.job[j];
    o.dy_t = g.dy_t; o.x_t = g.x_t; o.x2_t = g.x2_t;
    o.dy_row0 = (int)g.dy_row0; o.x_row0 = (int)g.x_row0; o.x2_row0 = (int)g.x2_row0;
    o.dy_col = g.dy_col; o.x_col = g.x_col; o.x2_col = g.x2_col;
    o.M = g.M; o.N = g.N; o.m_tiles = (g.M + 127) / 128;
    o.dW = g.dW; o.ldw = g.ldw; o.w_col = g.w_col; o.n_valid = g.n_valid > 0 ? g.n_valid : g.N;
    o.db = g.db; o.dW2 = g.dW2; o.ldw2 = g.ldw2; o.w2_col = g.w2_col; o.n_valid2 = g.n_valid2;
    w[j] = (double)o.m_tiles * (kABytes / 2 * (g.M >= 128 ? 2 : 1) + g.N * 128 + (g.x2_t >= 0 ? kB2Bytes : 0));
    wsum += w[j];
  }
  int total = 0;
  for (int j = 0; j < d.n_jobs; ++j) {
    JobArgs& o = a.job[j];
    int splits = (int)(kNumSMs * w[j] / wsum / o.m_tiles + 0.5);
    splits = std::max(1, std::min(splits, a.total_kb));
    o.kb_per_cta = (a.total_kb + splits - 1) / splits;
    splits = (a.total_kb + o.kb_per_cta - 1) / o.kb_per_cta;
    o.n_ctas = splits * o.m_tiles;
    o.cta0 = total;
    total += o.n_ctas;
  }
  static bool attr[64] = {};
  if (once_per_device(attr)) { NMX_CUDA(cudaFuncSetAttribute(wgrad_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlloc)); }
  double flops = 0.0;
  for (int j = 0; j < d.n_jobs; ++j) flops += 2.0 * (double)d.P * d.job[j].M * (d.job[j].N + (d.job[j].x2_t >= 0 ? 64 : 0));
  prof_begin(1, flops, stream);
  wgrad_batch_kernel<<<total, kThreads, kAlloc, stream>>>(maps, a);
  prof_end(stream);
  NMX_LAUNCH_CHECK();
  return 0;
}

}  // namespace nmx
