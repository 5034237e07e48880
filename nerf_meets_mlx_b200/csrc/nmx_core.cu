// libnmx core: error state, launch counter, sampling (K1) and stand-alone positional encodings (K2a/K2b).
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>

#include "nmx_common.cuh"

namespace nmx {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int debug_sync(const char* func, int line) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("NMX_DEBUG_SYNC");
    enabled = (e && e[0] == '1') ? 1 : 0;
  }
  if (!enabled) return 0;
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    set_error("%s:%d: kernel failed: %s", func, line, cudaGetErrorString(err));
    fprintf(stderr, "[nmx] %s:%d: kernel failed: %s\n", func, line, cudaGetErrorString(err));
    return (int)err;
  }
  return 0;
}
}  // namespace nmx

using namespace nmx;

extern "C" int nmx_version(void) { return NMX_VERSION; }
extern "C" const char* nmx_last_error_string(void) { return nmx::g_err; }
extern "C" int64_t nmx_launch_count(void) { return nmx::g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// K1: depth sampling.  Explicit _rn intrinsics pin the reference's operation order (no FMA contraction):
//   t = i*step (+0);  z = near*(1-t) + far*t             (sampling/uniform.py:13-16)
//   z = 1/(1/(near*(1-t)) + 1/(far*t))                   (sampling/linear_disparity.py:14-17, as written)
__global__ void sample_z_kernel(const float* __restrict__ near, const float* __restrict__ far, int64_t bound_stride,
                                float* __restrict__ z, int64_t total, int n, float step, int lindisp) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = idx / n;
    int i = (int)(idx - b * n);
    float t = __fmul_rn((float)i, step);
    float nr = near[b * bound_stride], fr = far[b * bound_stride];
    float a = __fmul_rn(nr, __fsub_rn(1.0f, t));
    float c = __fmul_rn(fr, t);
    float v;
    if (!lindisp) {
      v = __fadd_rn(a, c);
    } else {
      v = __fdiv_rn(1.0f, __fadd_rn(__fdiv_rn(1.0f, a), __fdiv_rn(1.0f, c)));
    }
    z[idx] = v;
  }
}

extern "C" int nmx_sample_z_fwd(const float* near, const float* far, float* z, int64_t B, int n, int lindisp,
                                void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 2, "B >= 0, n >= 2");
  if (B == 0) return 0;
  float step = (float)((1.0 - 0.0) / (double)(n - 1));
  int64_t total = B * n;
  sample_z_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(near, far, 1, z, total, n, step, lindisp);
  NMX_LAUNCH_CHECK();
  return 0;
}

// same, with near / far read straight from the assembled ray rows (columns 6 and 7 of [B, ray_stride], render.py:105-106)
extern "C" int nmx_sample_z_rays(const float* rays, int ray_stride, float* z, int64_t B, int n, int lindisp, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 2 && ray_stride >= 8, "B >= 0, n >= 2, ray_stride >= 8");
  if (B == 0) return 0;
  NMX_CHECK_ARG(rays && z, "rays, z non-null");
  float step = (float)((1.0 - 0.0) / (double)(n - 1));
  int64_t total = B * n;
  sample_z_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(rays + 6, rays + 7, ray_stride, z, total, n, step, lindisp);
  NMX_LAUNCH_CHECK();
  return 0;
}

// add_noise_z (sampling/__init__.py:17-29): mids, upper=[mids, z_last], lower=[z_0, mids];
// z = lower + (upper-lower) * (t_rand*strength)
__global__ void add_noise_z_kernel(const float* __restrict__ z, const float* __restrict__ t_rand,
                                   float* __restrict__ out, int64_t total, int n, float strength) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int i = (int)(idx % n);
    float zi = z[idx];
    float upper = (i == n - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, z[idx + 1]));
    float lower = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(z[idx - 1], zi));
    float t = __fmul_rn(t_rand[idx], strength);
    out[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
  }
}

extern "C" int nmx_add_noise_z_fwd(const float* z, const float* t_rand, float* z_out, int64_t B, int n,
                                   float strength, void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 1, "B >= 0, n >= 1");
  if (B == 0) return 0;
  int64_t total = B * n;
  if (strength <= 0.0f) {  // sampling/__init__.py:14-15 returns the input unchanged
    if (z_out != z) NMX_CUDA(cudaMemcpyAsync(z_out, z, total * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
  }
  add_noise_z_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(z, t_rand, z_out, total, n, strength);
  NMX_LAUNCH_CHECK();
  return 0;
}

// pos = o + z*d (rendering/render.py:142)
__global__ void ray_points_kernel(const float* __restrict__ rays, int ray_stride, const float* __restrict__ z,
                                  float* __restrict__ pos, int64_t total, int n) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t pt = idx / 3;
    int c = (int)(idx - pt * 3);
    int64_t b = pt / n;
    const float* r = rays + b * ray_stride;
    pos[idx] = __fadd_rn(r[c], __fmul_rn(z[pt], r[3 + c]));
  }
}

extern "C" int nmx_ray_points_fwd(const float* rays, int ray_stride, const float* z, float* pos, int64_t B, int n,
                                  void* stream) {
  NMX_CHECK_ARG(B >= 0 && n >= 1 && ray_stride >= 6, "B >= 0, n >= 1, ray_stride >= 6");
  if (B == 0) return 0;
  int64_t total = B * n * 3;
  ray_points_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(rays, ray_stride, z, pos, total, n);
  NMX_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K2a: Embedder PE (models/embedding.py:35-71).  Channel c of the output:
//   c < inc           -> x[c]                      (inc = include_input ? in_dim : 0)
//   else k=(c-inc)/(2*in_dim), r=(c-inc)%(2*in_dim), fn = r/in_dim (0 sin, 1 cos), d = r%in_dim,
//        f_k = k^2 (reference quirk) ; out = fn(x[d]*f_k)
// Full-range sinf/cosf (no fast-math): arguments reach |x|*81.
// A block encodes tiles of 32 points: threads take (point, frequency, dim) items and produce sin and cos together
// (sincosf), rows are staged in shared memory and the tile leaves as one contiguous, coalesced span -- several
// independent range reductions in flight per thread, no 64-bit index division per element.
constexpr int kPeTile = 32;
__global__ void __launch_bounds__(256)
pe_embedder_kernel(const float* __restrict__ x, const float* __restrict__ bands, float* __restrict__ out, int64_t P, int in_dim,
                   int n_freqs, int inc) {
  extern __shared__ float s_rows[];
  const int out_dim = inc + 2 * in_dim * n_freqs;
  const int items = in_dim * n_freqs;
  float* s_x = s_rows + kPeTile * out_dim;  // [kPeTile * in_dim]
  for (int64_t p0 = (int64_t)blockIdx.x * kPeTile; p0 < P; p0 += (int64_t)gridDim.x * kPeTile) {
    const int np = (int)((P - p0 < kPeTile) ? (P - p0) : kPeTile);
    for (int i = threadIdx.x; i < np * in_dim; i += 256) s_x[i] = x[p0 * in_dim + i];
    __syncthreads();
    for (int i = threadIdx.x; i < np * inc; i += 256) {
      const int pt = i / inc, c = i - pt * inc;
      s_rows[pt * out_dim + c] = s_x[pt * in_dim + c];
    }
#pragma unroll 4
    for (int i = threadIdx.x; i < np * items; i += 256) {
      const int pt = i / items, q = i - pt * items;
      const int k = q / in_dim, d = q - k * in_dim;
      const float a = __fmul_rn(s_x[pt * in_dim + d], bands != nullptr ? __ldg(bands + k) : (float)(k * k));
      float sv, cv;
      sincosf(a, &sv, &cv);
      float* row = s_rows + pt * out_dim + inc + k * 2 * in_dim;
      row[d] = sv;
      row[in_dim + d] = cv;
    }
    __syncthreads();
    float* o = out + p0 * out_dim;
    for (int i = threadIdx.x; i < np * out_dim; i += 256) o[i] = s_rows[i];
    __syncthreads();
  }
}

extern "C" int nmx_pe_embedder_bands_fwd(const float* x, const float* bands, float* out, int64_t P, int in_dim, int n_freqs,
                                         int include_input, void* stream) {
  NMX_CHECK_ARG(P >= 0 && in_dim >= 1 && n_freqs >= 0, "P >= 0, in_dim >= 1, n_freqs >= 0");
  if (P == 0) return 0;
  int inc = include_input ? in_dim : 0;
  const int out_dim = inc + 2 * in_dim * n_freqs;
  if (out_dim == 0) return 0;
  NMX_CHECK_ARG(out_dim <= 1024, "encoded width <= 1024");
  const size_t smem = (size_t)kPeTile * (out_dim + in_dim) * sizeof(float);
  if (smem > 48 * 1024) NMX_CUDA(cudaFuncSetAttribute(pe_embedder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pe_embedder_kernel<<<grid_for(P, kPeTile, 16), 256, smem, (cudaStream_t)stream>>>(x, bands, out, P, in_dim, n_freqs, inc);
  NMX_LAUNCH_CHECK();
  return 0;
}

extern "C" int nmx_pe_embedder_fwd(const float* x, float* out, int64_t P, int in_dim, int n_freqs, int include_input,
                                   void* stream) {
  return nmx_pe_embedder_bands_fwd(x, nullptr, out, P, in_dim, n_freqs, include_input, stream);
}

// K2b: SinusoidalEncoding (encoding/sinusoidal.py:49-66).  out[c]:
//   c < in_dim*n_freqs      -> sin(x[d]*band[k]),            d = c / n_freqs, k = c % n_freqs
//   c < 2*in_dim*n_freqs    -> sin(x[d]*band[k] + fp32(pi/2))  (the reference's cos)
//   else                    -> x[c - 2*in_dim*n_freqs]
__global__ void __launch_bounds__(256)
pe_sinusoidal_kernel(const float* __restrict__ x, const float* __restrict__ bands, float* __restrict__ out, int64_t P,
                     int in_dim, int n_freqs, int out_dim) {
  extern __shared__ float s_rows[];
  const float half_pi = 1.57079637050628662109375f;  // fp32(pi/2)
  const int half = in_dim * n_freqs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float* row = s_rows + warp * out_dim;
  for (int64_t p = (int64_t)blockIdx.x * nw + warp; p < P; p += (int64_t)gridDim.x * nw) {
    const float* xp = x + p * in_dim;
    for (int q = lane; q < half; q += 32) {
      const int d = q / n_freqs, k = q - d * n_freqs;
      const float sc = __fmul_rn(xp[d], bands[k]);
      row[q] = sinf(sc);
      row[half + q] = sinf(__fadd_rn(sc, half_pi));
    }
    for (int c = 2 * half + lane; c < out_dim; c += 32) row[c] = xp[c - 2 * half];
    __syncwarp();
    float* o = out + p * out_dim;
    for (int c = lane; c < out_dim; c += 32) o[c] = row[c];
    __syncwarp();
  }
}

extern "C" int nmx_pe_sinusoidal_fwd(const float* x, const float* bands, float* out, int64_t P, int in_dim,
                                     int n_freqs, int include_input, void* stream) {
  NMX_CHECK_ARG(P >= 0 && in_dim >= 1 && n_freqs >= 1, "P >= 0, in_dim >= 1, n_freqs >= 1");
  if (P == 0) return 0;
  int out_dim = 2 * in_dim * n_freqs + (include_input ? in_dim : 0);
  NMX_CHECK_ARG(out_dim <= 1024, "encoded width <= 1024");
  pe_sinusoidal_kernel<<<grid_for(P, 8, 16), 256, 8 * out_dim * sizeof(float), (cudaStream_t)stream>>>(x, bands, out, P, in_dim, n_freqs, out_dim);
  NMX_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Spherical-harmonics direction encoding, degree <= 4 (encoding/spherical_harmonics.py:33-94).  One thread per
// direction; every product / sum is a separately rounded fp32 operation in the reference's evaluation order
// (Python scalar * fp32 array stays fp32), so the result is bit-comparable with the restated reference.
__global__ void __launch_bounds__(256)
sh_encode_kernel(const float* __restrict__ dirs, int in_dim, float* __restrict__ out, int64_t B, int level) {
  // one thread per direction; the block's 256 rows are staged in shared memory (row pitch od: odd for every level, so
  // the strided per-thread writes are conflict-free) and written out as one contiguous, coalesced span
  extern __shared__ float s_tile[];
  const int od = (level + 1) * (level + 1);
  for (int64_t p0 = (int64_t)blockIdx.x * 256; p0 < B; p0 += (int64_t)gridDim.x * 256) {
    const int64_t p = p0 + threadIdx.x;
    const int64_t pc = p < B ? p : B - 1;
    const float x = dirs[pc * in_dim + 0], y = dirs[pc * in_dim + 1], z = dirs[pc * in_dim + 2];
#define M_(a, b) __fmul_rn((a), (b))
#define S_(a, b) __fsub_rn((a), (b))
#define A_(a, b) __fadd_rn((a), (b))
    const float xx = M_(x, x), yy = M_(y, y), zz = M_(z, z), xy = M_(x, y), yz = M_(y, z), xz = M_(x, z);
    float* o = s_tile + threadIdx.x * od;
    o[0] = 0.28209479177387814f;
    if (level >= 1) {
      o[1] = M_(0.4886025119029199f, y);
      o[2] = M_(0.4886025119029199f, z);
      o[3] = M_(0.4886025119029199f, x);
    }
    if (level >= 2) {
      o[4] = M_(1.0925484305920792f, xy);
      o[5] = M_(1.0925484305920792f, yz);
      o[6] = S_(M_(0.9461746957575601f, zz), 0.31539156525251999f);
      o[7] = M_(1.0925484305920792f, xz);
      o[8] = M_(0.5462742152960396f, S_(xx, yy));
    }
    if (level >= 3) {
      const float t3xx_yy = S_(M_(3.f, xx), yy), xx_3yy = S_(xx, M_(3.f, yy));
      o[9] = M_(M_(0.5900435899266435f, y), t3xx_yy);
      o[10] = M_(M_(2.890611442640554f, xy), z);
      o[11] = M_(M_(0.4570457994644658f, y), S_(M_(5.f, zz), 1.f));
      o[12] = M_(M_(0.3731763325901154f, z), S_(M_(5.f, zz), 3.f));
      o[13] = M_(M_(0.4570457994644658f, x), S_(M_(5.f, zz), 1.f));
      o[14] = M_(M_(1.445305721320277f, z), S_(xx, yy));
      o[15] = M_(M_(0.5900435899266435f, x), xx_3yy);
      if (level >= 4) {
        o[16] = M_(M_(2.5033429417967046f, xy), S_(xx, yy));
        o[17] = M_(M_(1.7701307697799304f, yz), t3xx_yy);
        o[18] = M_(M_(0.9461746957575601f, xy), S_(M_(7.f, zz), 1.f));
        o[19] = M_(M_(0.6690465435572892f, yz), S_(M_(7.f, zz), 3.f));
        o[20] = M_(0.10578554691520431f, A_(S_(M_(M_(35.f, zz), zz), M_(30.f, zz)), 3.f));
        o[21] = M_(M_(0.6690465435572892f, xz), S_(M_(7.f, zz), 3.f));
        o[22] = M_(M_(0.47308734787878004f, S_(xx, yy)), S_(M_(7.f, zz), 1.f));
        o[23] = M_(M_(1.7701307697799304f, xz), xx_3yy);
        o[24] = M_(0.6258357354491761f, S_(M_(xx, xx_3yy), M_(yy, t3xx_yy)));
      }
    }
#undef M_
#undef S_
#undef A_
    __syncthreads();
    const int64_t nrow = (B - p0 < 256) ? (B - p0) : 256;
    const int n_out = (int)nrow * od;
    float* dst = out + p0 * od;
    for (int i = threadIdx.x; i < n_out; i += 256) dst[i] = s_tile[i];
    __syncthreads();
  }
}

extern "C" int nmx_sh_encode_fwd(const float* dirs, int in_dim, float* out, int64_t B, int n_degrees, void* stream) {
  NMX_CHECK_ARG(B >= 0 && in_dim >= 3 && n_degrees >= 0 && n_degrees <= 4, "B >= 0, in_dim >= 3, 0 <= n_degrees <= 4");
  if (B == 0) return 0;
  NMX_CHECK_ARG(dirs && out, "dirs, out non-null");
  const int od = (n_degrees + 1) * (n_degrees + 1);
  sh_encode_kernel<<<grid_for(B, 256, 16), 256, 256 * od * sizeof(float), (cudaStream_t)stream>>>(dirs, in_dim, out, B, n_degrees);
  NMX_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Ray generation + ray-batch assembly (rendering/ray.py:7-35, rendering/render.py:283-328 with ndc=False,
// __test_nerf.py:208-236): pixel id -> [o(3), d(3), near, far, viewdirs(3)] and, optionally, the target pixel.
// get_rays runs in float64 in the reference when K is a float64 array (NumPy >= 2 promotion) and is cast to fp32
// afterwards: the same here, with separately rounded products and the sequential 3-term sum NumPy uses.
__global__ void __launch_bounds__(256) gen_rays_kernel(const float* __restrict__ c2w, int c2w_ld, double fx, double fy, double cx, double cy,
                                int W, const int32_t* __restrict__ pix, int64_t B, float near, float far,
                                float* __restrict__ rays, int ray_stride, const float* __restrict__ image, int img_ld,
                                float* __restrict__ target) {
  // one thread per ray; rows are staged in shared memory (pitch 6 / 8 / 11 floats) and stored as one coalesced span
  extern __shared__ float s_tile[];
  for (int64_t b0 = (int64_t)blockIdx.x * 256; b0 < B; b0 += (int64_t)gridDim.x * 256) {
    const int64_t b = b0 + threadIdx.x;
    const bool live = b < B;
    const int64_t id = live ? (pix ? (int64_t)pix[b] : b) : 0;
    const int row = (int)(id / W), col = (int)(id - (int64_t)row * W);
    const double d0 = __ddiv_rn(__dsub_rn((double)(float)col, cx), fx);
    const double d1 = -__ddiv_rn(__dsub_rn((double)(float)row, cy), fy);
    const double d2 = -1.0;
    float d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double v = __dadd_rn(__dadd_rn(__dmul_rn(d0, (double)c2w[r * c2w_ld + 0]), __dmul_rn(d1, (double)c2w[r * c2w_ld + 1])),
                                 __dmul_rn(d2, (double)c2w[r * c2w_ld + 2]));
      d[r] = (float)v;
    }
    float* o = s_tile + threadIdx.x * ray_stride;
    o[0] = c2w[0 * c2w_ld + 3];
    o[1] = c2w[1 * c2w_ld + 3];
    o[2] = c2w[2 * c2w_ld + 3];
    o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
    if (ray_stride >= 8) { o[6] = near; o[7] = far; }
    if (ray_stride >= 11) {  // viewdirs = d / ||d|| in fp32 (render.py:307)
      const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2]));
      const float nrm = __fsqrt_rn(n2);
      o[8] = __fdiv_rn(d[0], nrm); o[9] = __fdiv_rn(d[1], nrm); o[10] = __fdiv_rn(d[2], nrm);
    }
    float* t = s_tile + 256 * ray_stride + threadIdx.x * 3;
    if (target && live) {
      const float* px = image + id * img_ld;
      t[0] = px[0]; t[1] = px[1]; t[2] = px[2];
    }
    __syncthreads();
    const int nrow = (int)((B - b0 < 256) ? (B - b0) : 256);
    float* dst = rays + b0 * ray_stride;
    for (int i = threadIdx.x; i < nrow * ray_stride; i += 256) dst[i] = s_tile[i];
    if (target) {
      float* td = target + b0 * 3;
      for (int i = threadIdx.x; i < nrow * 3; i += 256) td[i] = s_tile[256 * ray_stride + i];
    }
    __syncthreads();
  }
}

extern "C" int nmx_gen_rays(const float* c2w, int c2w_ld, double fx, double fy, double cx, double cy, int H, int W,
                            const int32_t* pix, int64_t B, float near, float far, float* rays, int ray_stride,
                            const float* image, int img_ld, float* target, void* stream) {
  NMX_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && c2w_ld >= 4 && (ray_stride == 6 || ray_stride == 8 || ray_stride == 11),
                "B >= 0; H, W >= 1; c2w_ld >= 4; ray_stride in {6, 8, 11}");
  NMX_CHECK_ARG(fx != 0.0 && fy != 0.0, "focal lengths must be non-zero");
  NMX_CHECK_ARG((target == nullptr) || (image != nullptr && img_ld >= 3), "target needs image with img_ld >= 3");
  NMX_CHECK_ARG(pix != nullptr || B <= (int64_t)H * W, "without pixel ids B <= H*W");
  if (B == 0) return 0;
  NMX_CHECK_ARG(c2w && rays, "c2w, rays non-null");
  gen_rays_kernel<<<grid_for(B, 256, 16), 256, 256 * (ray_stride + 3) * sizeof(float), (cudaStream_t)stream>>>(c2w, c2w_ld, fx, fy, cx, cy, W, pix, B, near, far,
                                                                      rays, ray_stride, image, img_ld, target);
  NMX_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- ray-batch assembly from explicit origins/directions
// rays[b] = [o(3), d(3), near, far, d/||d||(3)]: the row mlx_mse_coarse / mlx_mse_fine build before render_rays
// (__test_nerf.py:57-82, 97-104: viewdirs = d / ||d||, near / far columns), one thread per ray.
__global__ void __launch_bounds__(256)
assemble_rays_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t B, float near, float far,
                     float* __restrict__ rays) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float ox = o[3 * b], oy = o[3 * b + 1], oz = o[3 * b + 2];
    const float dx = d[3 * b], dy = d[3 * b + 1], dz = d[3 * b + 2];
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float nrm = __fsqrt_rn(n2);
    float* r = rays + 11 * b;
    r[0] = ox; r[1] = oy; r[2] = oz; r[3] = dx; r[4] = dy; r[5] = dz; r[6] = near; r[7] = far;
    r[8] = __fdiv_rn(dx, nrm); r[9] = __fdiv_rn(dy, nrm); r[10] = __fdiv_rn(dz, nrm);
  }
}

extern "C" int nmx_assemble_rays(const float* rays_o, const float* rays_d, int64_t B, float near, float far, float* rays,
                                 void* stream) {
  NMX_CHECK_ARG(B >= 0, "B >= 0");
  if (B == 0) return 0;
  NMX_CHECK_ARG(rays_o && rays_d && rays, "rays_o, rays_d, rays non-null");
  assemble_rays_kernel<<<grid_for(B, 256, 4), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, B, near, far, rays);
  NMX_LAUNCH_CHECK();
  return 0;
}
