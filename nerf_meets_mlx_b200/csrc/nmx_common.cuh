// Shared helpers for libnmx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/nmx.h"

namespace nmx {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// NMX_DEBUG_SYNC=1: synchronise after every launch and report the failing call site (debug builds of a run only)
int debug_sync(const char* func, int line);

#define NMX_CHECK_ARG(cond, msg)                              \
  do {                                                        \
    if (!(cond)) {                                            \
      ::nmx::set_error("%s: bad argument: %s", __func__, msg); \
      return NMX_E_BADARG;                                    \
    }                                                         \
  } while (0)

#define NMX_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::nmx::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(_e));      \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define NMX_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::nmx::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e));  \
      return (int)_e;                                                                     \
    }                                                                                     \
    ::nmx::count_launch();                                                                \
    { int _d = ::nmx::debug_sync(__func__, __LINE__); if (_d) return _d; }                \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// Timing experiments inside the product kernels ("skip the epilogue math", "no weight traffic", event traces ...) exist
// only in builds with -DNMX_EXPERIMENTS (python -m nerf_meets_mlx_b200.build --experiments).  In the default build
// NMX_DBG() is the constant 0, every such branch is compiled out, and no environment variable can reach a kernel.
#ifdef NMX_EXPERIMENTS
#define NMX_DBG(prm, bits) (((prm).dbg & (bits)) != 0)
#else
#define NMX_DBG(prm, bits) (false)
#endif
// host side: value of an experiment environment variable (0 in the default build)
inline int experiment_env(const char* name) {
#ifdef NMX_EXPERIMENTS
  const char* e = getenv(name);
  return e ? atoi(e) : 0;
#else
  (void)name;
  return 0;
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inclusive prefix sum across the warp (lane order)
__device__ __forceinline__ float warp_scan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// inclusive suffix sum across the warp (lane i gets sum over lanes >= i)
__device__ __forceinline__ float warp_rscan_incl(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

__device__ __forceinline__ double warp_scan_incl_f64(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// true exactly once per device and call site: function attributes (dynamic shared-memory size) are per DEVICE, so a
// per-process flag would leave the kernels of a second GPU in the same process unconfigured
inline bool once_per_device(bool (&done)[64]) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
  if (done[d]) return false;
  done[d] = true;
  return true;
}

inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace nmx
