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


// ---- HBM bandwidth probes (scripts/bw_probe.py -> profiles/r2_bw_probe.txt): what a write-only / read-only stream
// reaches on this part, measured with the store patterns the product kernels use.
// mode 1: vectorised st.global.v4 fill, grid-stride, full occupancy.
__global__ void __launch_bounds__(256)
bw_fill_v4_kernel(uint4* __restrict__ dst, int64_t n16) {
  const uint4 v = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = v;
}
// mode 3: read-only ld.global.v4 stream (xor-reduced so the loads stay alive)
__global__ void __launch_bounds__(256)
bw_read_v4_kernel(const uint4* __restrict__ src, int64_t n16, uint32_t* __restrict__ sink) {
  uint32_t acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(src + i);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x9E3779B9u) *sink = acc;
}
// mode 4: copy (ld.global.v4 -> st.global.v4)
__global__ void __launch_bounds__(256)
bw_copy_v4_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __ldg(src + i);
}
// mode 2: bulk-async stores only.  One CTA per SM; each CTA owns a 64 KB staging tile in shared memory and writes it to
// consecutive 64 KB tiles of global memory as 4 x 16 KB cp.async.bulk.global.shared::cta copies per tile, `depth` bulk
// groups in flight (the fused chain keeps 1 tile-layer in flight per CTA).
template <int DEPTH>
__global__ void __launch_bounds__(128, 1)
bw_bulk_store_kernel(uint8_t* __restrict__ dst, int64_t n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll
      for (int c = 0; c < 4; ++c) bulk_store_1d(dst + t * 65536 + c * 16384, smem + c * 16384, 16384);
      tma_store_commit();
      tma_store_wait_read<DEPTH - 1>();
    }
    tma_store_wait<0>();
  }
}
// mode 5: the fused chain's own store pattern: a [rows, 256] bf16 row-major tensor written as 128-row x 64-column
// SW128 boxes (cp.async.bulk.tensor.2d), 4 boxes = one 64 KB tile-layer, one bulk group per tile-layer.
template <int DEPTH>
__global__ void __launch_bounds__(128, 1)
bw_tensor_store_kernel(const __grid_constant__ CUtensorMap map, int64_t n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_store_2d(&map, smem + c * 16384, c * 64, (int)(t * 128));
      tma_store_commit();
      tma_store_wait_read<DEPTH - 1>();
    }
    tma_store_wait<0>();
  }
}
// mode 6: TMA tile loads only (the wgrad kernels' read pattern): 128 x 64 bf16 boxes of a [rows, 256] tensor into a
// 4-slot shared-memory ring
__global__ void __launch_bounds__(128, 1)
bw_tensor_load_kernel(const __grid_constant__ CUtensorMap map, int64_t n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map);
    int64_t issued = 0, done = 0;
    const int64_t mine = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t k) {
      const int64_t t = blockIdx.x + k * gridDim.x;
      const int slot = (int)(k & 1);
      mbar_arrive_expect_tx(&bar[slot], 65536);
#pragma unroll
      for (int c = 0; c < 4; ++c) tma_load_2d(smem + slot * 65536 + c * 16384, &map, &bar[slot], c * 64, (int)(t * 128));
    };
    for (; issued < 2 && issued < mine; ++issued) issue(issued);
    for (; done < mine; ++done) {
      mbar_wait(&bar[done & 1], (uint32_t)((done >> 1) & 1));
      if (issued < mine) { issue(issued); ++issued; }
    }
  }
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

// HBM bandwidth probes; `bytes` must be a multiple of 64 KB.  mode 0 cudaMemsetAsync, 1 st.global.v4 fill,
// 2 bulk-async (TMA 1-D) stores from shared memory, 3 ld.global.v4 read, 4 ld/st copy (src -> buf), 5 TMA 2-D tensor
// stores in the fused chain's pattern ([rows,256] bf16, 128x64 boxes), 6 TMA 2-D tile loads.  depth = bulk groups in flight.
extern "C" int nmx_diag_bw(int mode, void* buf, const void* src, int64_t bytes, int ctas, int depth, void* stream) {
  NMX_CHECK_ARG(buf && bytes > 0 && bytes % 65536 == 0 && ctas > 0 && depth >= 1 && depth <= 3, "buf non-null; bytes % 64 KiB == 0; 1 <= depth <= 3");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n16 = bytes / 16, n_tiles = bytes / 65536;
  const int smem_bytes = 65536 + 1024;
  if (mode == 0) {
    NMX_CUDA(cudaMemsetAsync(buf, 0x3c, (size_t)bytes, s));
    return 0;
  } else if (mode == 1) {
    bw_fill_v4_kernel<<<ctas, 256, 0, s>>>((uint4*)buf, n16);
  } else if (mode == 3) {
    bw_read_v4_kernel<<<ctas, 256, 0, s>>>((const uint4*)buf, n16, (uint32_t*)buf);
  } else if (mode == 4) {
    NMX_CHECK_ARG(src != nullptr, "copy needs src");
    bw_copy_v4_kernel<<<ctas, 256, 0, s>>>((const uint4*)src, (uint4*)buf, n16);
  } else if (mode == 2) {
#define NMX_BS(D)                                                                                                  \
  NMX_CUDA(cudaFuncSetAttribute(bw_bulk_store_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
  bw_bulk_store_kernel<D><<<ctas, 128, smem_bytes, s>>>((uint8_t*)buf, n_tiles)
    if (depth == 1) { NMX_BS(1); } else if (depth == 2) { NMX_BS(2); } else { NMX_BS(3); }
#undef NMX_BS
  } else if (mode == 5 || mode == 6) {
    CUtensorMap map;
    int rc = make_tmap_bf16_2d(&map, buf, (uint64_t)(bytes / 512), 256, 256, 128);
    if (rc) return rc;
    if (mode == 6) {
      NMX_CUDA(cudaFuncSetAttribute(bw_tensor_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 65536 + 1024));
      bw_tensor_load_kernel<<<ctas, 128, 2 * 65536 + 1024, s>>>(map, n_tiles);
    } else {
#define NMX_TS(D)                                                                                                    \
  NMX_CUDA(cudaFuncSetAttribute(bw_tensor_store_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)); \
  bw_tensor_store_kernel<D><<<ctas, 128, smem_bytes, s>>>(map, n_tiles)
      if (depth == 1) { NMX_TS(1); } else if (depth == 2) { NMX_TS(2); } else { NMX_TS(3); }
#undef NMX_TS
    }
  } else {
    set_error("nmx_diag_bw: mode in [0,6]");
    return NMX_E_BADARG;
  }
  NMX_LAUNCH_CHECK();
  return 0;
}
