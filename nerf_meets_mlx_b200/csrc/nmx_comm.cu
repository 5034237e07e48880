// K6 (multi-GPU): the data-parallel gradient exchange of the training step, fused with the optimiser update.
//
// The reference has no distributed code; data parallelism is the B200 build's own layer: every
// rank computes the gradient of its ray shard, the flat fp32 gradients (2.38 MB per net) are averaged, and identical
// replicated Adam updates follow (__test_nerf.py:128-145 semantics per replica).  NCCL does this as
// all-reduce -> scale -> Adam = three passes over the buffer plus a collective launched from the host, which cannot sit
// inside a captured CUDA graph without the teardown problems seen in round 1.  Here it is ONE kernel over NVLink peer
// memory (CUDA IPC mappings of every rank's gradient buffer):
//
//   entry barrier : block 0 publishes "my gradient of epoch e is complete" into every peer's flag row (system-scope
//                   release store); every CTA polls its LOCAL flag row until all ranks have published e;
//   reduce + Adam : each thread owns float4 elements: loads them from all `world` gradient buffers (its own and the
//                   peers' through NVLink, ld.relaxed.sys), sums IN RANK ORDER (so every rank computes bit-identical
//                   sums), scales by 1/world and applies the MLX-style Adam update to the local replica (p, m, v);
//   exit barrier  : the last CTA to finish publishes "done reading e" to every peer and waits for every peer's "done"
//                   before the kernel ends, so whatever follows in the stream may overwrite the gradient buffer.
//
// The epoch lives in device memory, so the kernel is replayed unchanged from a CUDA graph.  Every spin loop has a
// clock64() timeout that records an error code instead of hanging the GPU when a rank is missing.
#include "nmx_common.cuh"
#include "nmx_optim.cuh"

using namespace nmx;

namespace {

constexpr int kMaxWorld = 8;
constexpr long long kSpinTimeout = 4000000000ll;  // ~2 s of SM clocks

// flags block (one per rank, peer-mapped), all 64-bit: [0, 8) ready[src rank], [8, 16) done[src rank],
// [16] epoch, [17] CTA arrival counter, [18] error code (0 ok, 1 entry timeout, 2 exit timeout)
constexpr int kReady = 0, kDone = 8, kEpoch = 16, kArrive = 17, kError = 18;

struct P2PArgs {
  const float* grads[kMaxWorld];             // gradient buffer of every rank (own + peer mappings), same layout
  unsigned long long* flags[kMaxWorld];      // flags block of every rank
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
allreduce_adam_kernel(const P2PArgs a, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ g_avg, int64_t count, float lr, const float* __restrict__ lr_dev, float b1,
                      float b2, float eps) {
  unsigned long long* my = a.flags[a.rank];
  __shared__ unsigned long long s_epoch;
  __shared__ int s_last;
  if (threadIdx.x == 0) s_epoch = ld_acquire_sys(my + kEpoch) + 1ull;
  __syncthreads();
  const unsigned long long e = s_epoch;
  // ---- entry barrier
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + kReady + a.rank, e);
  }
  if (threadIdx.x < a.world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(my + kReady + threadIdx.x) < e) {
      if (clock64() - t0 > kSpinTimeout) { atomicMax(my + kError, 1ull); break; }
    }
  }
  __syncthreads();
  // ---- reduce (rank order) + scale + Adam
  if (lr_dev != nullptr) lr = __ldg(lr_dev);
  const float inv_world = 1.0f / (float)a.world;
  const int64_t n4 = count >> 2;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nth) {
    float4 s = ld_relaxed_sys_f4(a.grads[0] + 4 * i);
    for (int r = 1; r < a.world; ++r) {
      const float4 t = ld_relaxed_sys_f4(a.grads[r] + 4 * i);
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    s.x *= inv_world; s.y *= inv_world; s.z *= inv_world; s.w *= inv_world;
    float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
    adam_update(pi.x, mi.x, vi.x, s.x, lr, b1, b2, eps, 1.0f, 1.0f);
    adam_update(pi.y, mi.y, vi.y, s.y, lr, b1, b2, eps, 1.0f, 1.0f);
    adam_update(pi.z, mi.z, vi.z, s.z, lr, b1, b2, eps, 1.0f, 1.0f);
    adam_update(pi.w, mi.w, vi.w, s.w, lr, b1, b2, eps, 1.0f, 1.0f);
    reinterpret_cast<float4*>(p)[i] = pi;
    reinterpret_cast<float4*>(m)[i] = mi;
    reinterpret_cast<float4*>(v)[i] = vi;
    if (g_avg != nullptr) reinterpret_cast<float4*>(g_avg)[i] = s;
  }
  for (int64_t i = (n4 << 2) + tid; i < count; i += nth) {  // tail (count % 4)
    float s = ld_relaxed_sys_f1(a.grads[0] + i);
    for (int r = 1; r < a.world; ++r) s += ld_relaxed_sys_f1(a.grads[r] + i);
    s *= inv_world;
    float pi = p[i], mi = m[i], vi = v[i];
    adam_update(pi, mi, vi, s, lr, b1, b2, eps, 1.0f, 1.0f);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (g_avg != nullptr) g_avg[i] = s;
  }
  // ---- exit barrier (last CTA)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(my + kArrive, 1ull) == (unsigned long long)gridDim.x - 1ull) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < a.world) {
    st_release_sys(a.flags[threadIdx.x] + kDone + a.rank, e);
    const long long t0 = clock64();
    while (ld_acquire_sys(my + kDone + threadIdx.x) < e) {
      if (clock64() - t0 > kSpinTimeout) { atomicMax(my + kError, 2ull); break; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    my[kArrive] = 0ull;
    st_release_sys(my + kEpoch, e);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ peer memory
extern "C" int nmx_p2p_alloc(int64_t bytes, void** ptr, void* handle64) {
  NMX_CHECK_ARG(bytes > 0 && ptr && handle64, "bytes > 0; ptr, handle non-null");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* d = nullptr;
  NMX_CUDA(cudaMalloc(&d, (size_t)bytes));
  NMX_CUDA(cudaMemset(d, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, d);
  if (e != cudaSuccess) {
    cudaFree(d);
    set_error("nmx_p2p_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *ptr = d;
  return 0;
}

extern "C" int nmx_p2p_open(const void* handle64, void** ptr) {
  NMX_CHECK_ARG(handle64 && ptr, "handle, ptr non-null");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  NMX_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int nmx_p2p_close(void* ptr) {
  if (ptr) NMX_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int nmx_p2p_free(void* ptr) {
  if (ptr) NMX_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int nmx_allreduce_adam(const void* const* grads, void* const* flags, int rank, int world, float* p, float* m,
                                  float* v, float* g_avg, int64_t count, float lr, const float* lr_dev, float b1,
                                  float b2, float eps, void* stream) {
  NMX_CHECK_ARG(grads && flags && p && m && v && count > 0, "grads, flags, p, m, v non-null; count > 0");
  NMX_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "1 <= world <= 8; 0 <= rank < world");
  P2PArgs a;
  for (int r = 0; r < kMaxWorld; ++r) {
    a.grads[r] = (const float*)grads[r < world ? r : 0];
    a.flags[r] = (unsigned long long*)flags[r < world ? r : 0];
    if (r < world) {
      NMX_CHECK_ARG(a.grads[r] && a.flags[r], "every rank's gradient / flags pointer non-null");
      NMX_CHECK_ARG((reinterpret_cast<uintptr_t>(a.grads[r]) & 15) == 0, "gradient buffers 16-byte aligned");
    }
  }
  a.rank = rank; a.world = world;
  // one float4 per thread and pass at the flat-parameter sizes of the path (595 844 floats -> 146 CTAs of 1024 elements)
  int blocks = (int)((count / 4 + 255) / 256);
  if (blocks > kNumSMs) blocks = kNumSMs;
  if (blocks < 1) blocks = 1;
  allreduce_adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, p, m, v, g_avg, count, lr, lr_dev, b1, b2, eps);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_p2p_flags_bytes(void) { return 32 * 8; }
