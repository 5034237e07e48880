// Diagnostics (not on the product path): raw tcgen05.mma issue-rate probe used to calibrate the MLP kernels' roofline.
// Every CTA issues `iters` back-to-back M=128 x N x K=16 bf16 MMAs on resident shared-memory operands (no TMA, no
// epilogue) and reports SM clocks and nanoseconds, so the per-SM MMA period and the clock the chip sustains under a
// full-chip tensor load can be read directly.
#include "nmx_common.cuh"
#include "nmx_sm100.cuh"

using namespace nmx;
using namespace nmx::sm100;

namespace {

__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(int N, int iters, int n_slabs, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];  // mode >= 1: commit target every 4 MMAs; mode 2: an always-complete barrier that is polled
  __shared__ uint32_t tmem_ptr;
  // A: 128 x 64 bf16 (16 KB), B slabs: n_slabs x (256 x 64 bf16 = 32 KB); pseudo-random bf16 values in [-1, 1)
  const int total_words = (16384 + n_slabs * 32768) / 4;
  uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
  for (int i = threadIdx.x; i < total_words; i += blockDim.x) {
    x = x * 1664525u + 1013904223u;
    uint32_t lo = 0x3C00u | ((x >> 9) & 0x83FFu), hi = 0x3C00u | ((x >> 20) & 0x83FFu);
    reinterpret_cast<uint32_t*>(smem)[i] = lo | (hi << 16);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2[0], 1);
    mbar_init(&bar2[1], 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  if (threadIdx.x < 32) {  // warp-uniform issue loop, one elected lane issues (as in the product kernels)
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t a_addr = smem_u32(smem);
    const bool alt = (mode & 4) && N <= 128;
    long long c0 = clock64();
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    int slab = 0;
    for (int i = 0; i < iters; i += 4) {
      if (++slab == n_slabs) slab = 0;
      const uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
      const uint64_t bdesc = make_smem_desc(a_addr + 16384 + slab * 32768, 16, 1024);
      const uint32_t d = tmem_base + ((i >> 4) & 1) * 256 + (alt ? ((i >> 2) & 1) * 128 : 0);
      if ((mode & 3) >= 2) {
        mbar_wait(&bar2[1], 1);  // parity of the phase before the current one: completes immediately
        tc_fence_after();
      }
      if (elect_one()) {
        umma_bf16(d, adesc, bdesc, idesc, (i & 15) != 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_bf16(d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, 1u);
        if ((mode & 3) >= 1) umma_commit(&bar2[0]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long c1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (threadIdx.x == 0) {
      out[blockIdx.x * 2 + 0] = c1 - c0;
      out[blockIdx.x * 2 + 1] = (long long)(g1 - g0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem_base);
}

// TMEM -> register read-rate probe: `warps` warps (warp w reads lane quadrant w & 3) issue `iters` back-to-back
// tcgen05.ld.32x32b.x16 (mode 0) or .x32 (mode 1) over the 512 allocated columns, one tcgen05.wait::ld per `batch`
// loads; out[0] = clocks for the whole CTA, out[1] = bytes read.
__global__ void __launch_bounds__(1024, 1)
tmem_ld_rate_kernel(int iters, int mode, int batch, long long* out) {
  __shared__ uint32_t tmem_ptr;
  __shared__ long long t0s, t1s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<512>(&tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  if (threadIdx.x == 0) t0s = clock64();
  __syncthreads();
  const int cols = mode ? 32 : 16;
  for (int i = 0; i < iters; i += batch) {
    for (int b = 0; b < batch; ++b) {
      const uint32_t col = (uint32_t)(((i + b) * cols + (warp >> 2) * 64) & 511) & ~(uint32_t)(cols - 1);
      if (mode) {
        uint32_t r[32];
        tmem_ld_32x32(base + (col & 480u), r);
        if (b == batch - 1) tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= r[k];
      } else {
        uint32_t r[16];
        tmem_ld_32x16(base + (col & 496u), r);
        if (b == batch - 1) tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc ^= r[k];
      }
    }
  }
  tmem_ld_wait();
  __syncthreads();
  if (threadIdx.x == 0) {
    t1s = clock64();
    out[blockIdx.x * 2 + 0] = t1s - t0s;
    out[blockIdx.x * 2 + 1] = (long long)iters * cols * 4 * 32 * (blockDim.x >> 5);
  }
  if (acc == 0x12345u && lane == 0) out[0] = -1;  // keep the loads alive
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_ptr);
}

}  // namespace

extern "C" int nmx_diag_tmem_ld_rate(int iters, int warps, int mode, int batch, int ctas, long long* out, void* stream) {
  NMX_CHECK_ARG(out && iters > 0 && warps >= 1 && warps <= 32 && batch >= 1 && ctas > 0, "iters > 0; 1 <= warps <= 32; batch >= 1");
  tmem_ld_rate_kernel<<<ctas, warps * 32, 0, (cudaStream_t)stream>>>(iters, mode, batch, out);
  NMX_LAUNCH_CHECK();
  return 0;
}

// out: int64 [2 * ctas] device buffer = (SM clocks, nanoseconds) per CTA
extern "C" int nmx_diag_mma_rate(int N, int iters, int n_slabs, int ctas, long long* out, void* stream, int mode) {
  NMX_CHECK_ARG(out && (N == 64 || N == 128 || N == 256) && iters > 0 && n_slabs >= 1 && n_slabs <= 6 && ctas > 0,
                "N in {64,128,256}; iters > 0; 1 <= n_slabs <= 6; ctas > 0");
  const int smem_bytes = 16384 + n_slabs * 32768 + 1024;
  NMX_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  mma_rate_kernel<<<ctas, 128, smem_bytes, (cudaStream_t)stream>>>(N, iters, n_slabs, mode, out);
  NMX_LAUNCH_CHECK();
  return 0;
}
