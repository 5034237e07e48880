/*
 * nmx.h -- C ABI of libnmx.so, the B200 (sm_100a) implementation of the volume-learning hot path of
 * piljoong-jeong/nerf_meets_mlx.
 *
 * The reference has no FFI: its boundary is the Python call surface of mlx_nerf/{sampling,encoding,
 * models,rendering}.  Each entry point below names the reference function (file:line, relative to the
 * reference repo root) whose arithmetic it replaces; the Python package `nerf_meets_mlx_b200` mirrors
 * those functions' names/signatures and calls these symbols through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (borrowed for the duration of the stream-ordered call) unless
 *     the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value 0 = success, non-zero = cudaError_t or NMX_E_* (message via nmx_last_error_string);
 *   - no hidden global state except a per-process cache of kernel attributes / driver entry points;
 *   - all tensors are dense row-major; fp32 unless stated; 16-byte aligned base pointers.
 */
#ifndef NMX_H_
#define NMX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NMX_VERSION 100
#define NMX_E_BADARG 10001
#define NMX_E_UNSUPPORTED 10002
#define NMX_E_DRIVER 10003

int nmx_version(void);
const char* nmx_last_error_string(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
int64_t nmx_launch_count(void);

/* ---------------------------------------------------------------- sampling (K1) */
/* uniform.sample_z (sampling/uniform.py:7-18) / linear_disparity.sample_z (linear_disparity.py:8-19).
 * near, far: [B]; z: [B, n].  t_i = fp32(i) * fp32(1/(n-1)); z = near*(1-t) + far*t (two-product form). */
int nmx_sample_z_fwd(const float* near, const float* far, float* z, int64_t B, int n, int lindisp, void* stream);
/* the same with near / far taken from columns 6 / 7 of the assembled ray rows [B, ray_stride] (render.py:105-106) */
int nmx_sample_z_rays(const float* rays, int ray_stride, float* z, int64_t B, int n, int lindisp, void* stream);
/* add_noise_z (sampling/__init__.py:10-31) with the uniform draw t_rand [B, n] explicit. */
int nmx_add_noise_z_fwd(const float* z, const float* t_rand, float* z_out, int64_t B, int n, float strength, void* stream);
/* pos = o + z*d (rendering/render.py:142): rays [B, ray_stride] (o at col 0, d at col 3), z [B, n] -> pos [B, n, 3] */
int nmx_ray_points_fwd(const float* rays, int ray_stride, const float* z, float* pos, int64_t B, int n, void* stream);

/* ---------------------------------------------------------------- ray generation */
/* get_rays (rendering/ray.py:7-35) + the ray-batch assembly of render() (rendering/render.py:283-328, ndc=False)
 * and of the training loop (__test_nerf.py:208-236), per pixel id (= row*W + col; pix == NULL -> ids 0..B-1):
 * rays[b] = [o(3), d(3), near, far, d/||d||(3)] truncated to ray_stride in {6, 8, 11}.  c2w: device [3, >=4] fp32
 * (row stride c2w_ld).  Pinhole K as doubles: the direction runs in float64 and is cast to fp32, as the reference
 * does with a float64 K.  Optional target gather: target[b, 0:3] = image[id, 0:3] (image [H*W, img_ld] fp32). */
int nmx_gen_rays(const float* c2w, int c2w_ld, double fx, double fy, double cx, double cy, int H, int W,
                 const int32_t* pix, int64_t B, float near, float far, float* rays, int ray_stride,
                 const float* image, int img_ld, float* target, void* stream);

/* Ray-batch assembly from explicit origins / directions [B, 3]: rays[b] = [o, d, near, far, d/||d||] ([B, 11]), the row
 * the reference's loss functions build before render_rays (__test_nerf.py:57-82, 97-104). */
int nmx_assemble_rays(const float* rays_o, const float* rays_d, int64_t B, float near, float far, float* rays,
                      void* stream);

/* ---------------------------------------------------------------- positional encodings (K2a/K2b) */
/* Embedder.embed (models/embedding.py:35-71): [x, sin(f0 x), cos(f0 x), ...], f_k = k^2 (reference quirk).
 * x: [P, in_dim] -> out: [P, (include_input?in_dim:0) + 2*in_dim*n_freqs] */
int nmx_pe_embedder_fwd(const float* x, float* out, int64_t P, int in_dim, int n_freqs, int include_input, void* stream);
/* the same with explicit frequency bands [n_freqs] fp32 on device (an Embedder built with max_freq_log2 != num_freqs - 1,
 * models/embedding.py:44-49: bands = linspace(0, max_freq_log2, num_freqs) ** 2); bands == NULL: f_k = k^2 */
int nmx_pe_embedder_bands_fwd(const float* x, const float* bands, float* out, int64_t P, int in_dim, int n_freqs,
                              int include_input, void* stream);
/* SinusoidalEncoding.__call__ (encoding/sinusoidal.py:39-66): sin([s, s + fp32(pi/2)]), s = x[:,d]*bands[k]
 * (dim-major, freq-minor); optional input appended at the end.  bands: [n_freqs] fp32 on device. */
int nmx_pe_sinusoidal_fwd(const float* x, const float* bands, float* out, int64_t P, int in_dim, int n_freqs,
                          int include_input, void* stream);

/* SphericalHarmonicsEncoding.__call__ (encoding/spherical_harmonics.py:33-94), real SH basis up to degree 4 of the
 * first three components of dirs [B, in_dim] -> out [B, (n_degrees+1)^2]; fp32, reference evaluation order. */
int nmx_sh_encode_fwd(const float* dirs, int in_dim, float* out, int64_t B, int n_degrees, void* stream);

/* ---------------------------------------------------------------- multires hash grid (K2c) */
/* MultiHashEncoding.hash (encoding/multi_hash.py:61-77): coords [M, 3] int32 -> idx [M] int32,
 * (x*1 ^ y*2654435761 ^ z*805459861) mod 2^log2_T in uint32 wraparound arithmetic. */
int nmx_hashgrid_hash(const int32_t* coords, int32_t* idx, int64_t M, int log2_T, void* stream);
/* MultiHashEncoding.__call__ (encoding/multi_hash.py:79-137): x [P,3], tables [L, 2^log2_T, F], scaled_res [L]
 * -> out [P, L*F]; optional idx_out [P, L, 8] int32 (corner table indices, reference corner order). F in {1,2,4}. */
int nmx_hashgrid_fwd(const float* x, const float* tables, const float* scaled_res, float* out, int32_t* idx_out,
                     int64_t P, int L, int F, int log2_T, void* stream);
/* gradient w.r.t. tables: d_tables [L, T, F] += scatter(d_out [P, L*F]) (caller zeroes d_tables). */
int nmx_hashgrid_bwd(const float* x, const float* scaled_res, const float* d_out, float* d_tables,
                     int64_t P, int L, int F, int log2_T, void* stream);

/* ---------------------------------------------------------------- compositing (K4) */
/* raw2outputs (rendering/render.py:20-96).  raw [B,n,4] (rgb, sigma), z [B,n], rays_d [B, d_stride] (first 3 used),
 * optional noise [B,n] (N(0,1)) scaled by raw_noise_std.  Outputs: rgb [B,3], disp [B], acc [B], weights [B,n],
 * depth [B].  Any output pointer may be NULL. */
int nmx_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                      float raw_noise_std, int white_bkgd, float* rgb, float* disp, float* acc, float* weights,
                      float* depth, int64_t B, int n, void* stream);
/* backward of raw2outputs w.r.t. raw.  d_rgb [B,3] required; d_disp/d_acc/d_depth [B], d_weights [B,n] optional (NULL).
 * d_raw [B,n,4].  n <= 256. */
int nmx_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride, const float* noise,
                      float raw_noise_std, int white_bkgd, const float* d_rgb, const float* d_disp,
                      const float* d_acc, const float* d_depth, const float* d_weights, float* d_raw,
                      int64_t B, int n, void* stream);

/* The loss functions of the training step (mlx_mse_coarse / mlx_mse_fine, __test_nerf.py:83-88, 111-124) fused:
 * rgb = raw2outputs(raw, z, rays_d)[0]; loss[0] += mean((rgb - target)^2) (caller zeroes loss); d_raw = d loss / d raw.
 * target [B,3]; optional outputs rgb [B,3], weights [B,n] (NULL to skip).  n <= 256. */
int nmx_composite_loss_fwd_bwd(const float* raw, const float* z, const float* rays_d, int d_stride, int white_bkgd,
                               const float* target, float* loss, float* d_raw, float* rgb, float* weights,
                               int64_t B, int n, void* stream);

/* ---------------------------------------------------------------- inverse-CDF resampling (K5) */
/* sample_from_inverse_cdf_torch (sampling/__init__.py:101-178) + sort-merge (rendering/render.py:225).
 * z [B,n], weights [B,n], u [B,N].  cdf_in [B,n+1] optional (NULL -> built in-kernel: fp64 sum / prefix).
 * Outputs (each optional): z_imp [B,N] unsorted (the reference function's result), inds [B,N] int32
 * (searchsorted-right), cdf_out [B,n+1], z_merged [B,n+N] ascending. */
int nmx_sample_pdf_fwd(const float* z, const float* weights, const float* u, const float* cdf_in, float eps,
                       float* z_imp, int32_t* inds, float* cdf_out, float* z_merged,
                       int64_t B, int n, int N, void* stream);
/* stand-alone sort(concat(a [B,na], b [B,nb])) -> out [B, na+nb]; a need not be sorted. */
int nmx_sort_merge_z(const float* a, const float* b, float* out, int64_t B, int na, int nb, void* stream);

/* ---------------------------------------------------------------- loss / optimiser (K6) */
/* mean((pred - target)^2) over count elements (__test_nerf.py:88,124); loss [1] is accumulated (caller zeroes);
 * d_pred = 2 (pred - target) / count * grad_scale (NULL to skip). */
int nmx_mse_fwd_bwd(const float* pred, const float* target, float* loss, float* d_pred, int64_t count,
                    float grad_scale, void* stream);
/* optim.Adam of MLX 0.7.0 (NeRF.py:120), no bias correction: m=b1 m+(1-b1)g; v=b2 v+(1-b2)g^2;
 * p -= lr*m/(sqrt(v)+eps).  With bias_correction!=0 uses the standard corrected form at step t. */
int nmx_adam_step(float* p, const float* g, float* m, float* v, int64_t count, float lr, float b1, float b2,
                  float eps, int bias_correction, int64_t t, void* stream);

/* same update without bias correction, learning rate read from device memory (lets a captured CUDA graph of the
 * training iteration follow the reference's decaying schedule, __test_nerf.py:302-305) */
int nmx_adam_step_lrdev(float* p, const float* g, float* m, float* v, int64_t count, const float* lr_dev, float b1,
                        float b2, float eps, void* stream);

/* ---------------------------------------------------------------- data-parallel gradient exchange */
/* Peer-mapped device memory (CUDA IPC, one process per GPU on one NVSwitch box).  nmx_p2p_alloc: cudaMalloc'd, zeroed
 * block + its 64-byte IPC handle (exchange the handles with any host-side transport, e.g. torch.distributed
 * all_gather_object); nmx_p2p_open maps a PEER's block into this process (peer access enabled lazily). */
int nmx_p2p_alloc(int64_t bytes, void** ptr, void* handle64);
int nmx_p2p_open(const void* handle64, void** ptr);
int nmx_p2p_close(void* ptr);
int nmx_p2p_free(void* ptr);
/* size of one rank's flags block (zero-initialised, peer-mapped): ready[8], done[8], epoch, arrival counter, error */
int nmx_p2p_flags_bytes(void);
/* One kernel = gradient all-reduce (mean) over NVLink peer memory + the MLX-style Adam update of the local replica
 * (the two optimiser steps of __test_nerf.py:128-145 under ray-sharded data parallelism).  grads[r] / flags[r]: HOST
 * arrays of `world` device pointers -- rank r's gradient buffer [count] fp32 and flags block as mapped in THIS process
 * (own allocation for r == rank).  Sums in rank order, so all ranks update bit-identically.  p, m, v: local replica.
 * g_avg: optional local copy of the averaged gradient.  lr_dev: optional device learning rate (CUDA-graph replays).
 * Entry and exit barriers across the ranks are inside the kernel (system-scope flags, device-side epoch): the call is
 * stream-ordered, capturable in a CUDA graph, and the gradient buffer may be overwritten as soon as it completes.
 * A rank that never arrives is reported in flags[18] (1 = entry, 2 = exit barrier timeout) instead of hanging. */
int nmx_allreduce_adam(const void* const* grads, void* const* flags, int rank, int world, float* p, float* m, float* v,
                       float* g_avg, int64_t count, float lr, const float* lr_dev, float b1, float b2, float eps,
                       void* stream);

/* ---------------------------------------------------------------- NeRF MLP (K3), tcgen05/TMEM/TMA */
/* Opaque plan for one NeRF network (models/NeRF.py:160-243) in its reference geometry.
 * Packed parameter layout (fp32, `params`, count = nmx_mlp_param_count): for each Linear in the order
 *   list_linears_pos[0..D-1], then (view-dir head) feature_linear, alpha_linear, list_linears_dir[0], rgb_linear
 *   or (no-view head) output_linear:  weight [out, in] row-major followed by bias [out]. */
typedef struct nmx_mlp_plan nmx_mlp_plan;

typedef struct {
  int n_layers;       /* D (8) */
  int width;          /* W (256); must be a multiple of 64, <= 256 */
  int in_pos;         /* encoded position channels (63 for Embedder N=10, 40 for image PE) */
  int in_dir;         /* encoded view-dir channels (27) or 0 */
  int out_ch;         /* no-view head output channels (channel_output) */
  int skip_layer;     /* index i such that [input_pos, h] is concatenated after layer i; -1 = none */
  int use_viewdirs;   /* 1 = view-dir head (rgb,alpha), 0 = output_linear */
  int n_freqs_pos, n_freqs_dir; /* encoder bands, used by the fused input encoders (enc_kind 1 / 2 below) */
} nmx_mlp_config;

int64_t nmx_mlp_param_count(const nmx_mlp_config* cfg);
int nmx_mlp_plan_create(const nmx_mlp_config* cfg, int64_t max_points, nmx_mlp_plan** out);
void nmx_mlp_plan_destroy(nmx_mlp_plan* plan);
/* bytes of device workspace the plan needs for `max_points` with/without saved activations */
int64_t nmx_mlp_workspace_bytes(const nmx_mlp_plan* plan, int training);
/* refresh the bf16 operand copies of the weights from fp32 params (call after every optimiser step) */
int nmx_mlp_load_params(nmx_mlp_plan* plan, const float* params, void* workspace, void* stream);

/* NeRF.forward via run_model (models/NeRF.py:25-48,201-243).  `params` = packed fp32 parameters (biases and the
 * tiny rgb/alpha/output heads are read from it directly; the bf16 operand copies come from nmx_mlp_load_params).
 * enc_kind selects how the bf16 operand tile is produced:
 *   enc_kind 0: x [P, in_pos+in_dir] fp32 (already encoded); P = B*n.
 *   enc_kind 1: rays [B, ray_stride] (o, d, near, far, viewdirs at the last 3 cols), z [B, n]; P = B*n.
 *   enc_kind 2: x [P, in_dim] raw coordinates, bands [n_freqs_pos]; P = B*n.
 * out [P, out_cols] fp32 (out_cols = 4 (rgb, sigma) for the view-dir head, out_ch otherwise).
 * save_activations != 0 keeps what nmx_mlp_bwd needs in the workspace (sized with training = 1). */
int nmx_mlp_fwd(nmx_mlp_plan* plan, void* workspace, const float* params, int enc_kind, const float* x_or_rays,
                int ray_stride, const float* z, const float* bands, float* out, int64_t B, int n,
                int save_activations, void* stream);
/* gradients of all parameters (packed like `params`, fp32, overwritten) from d_out [P, out_cols], using the
 * activations saved by the preceding nmx_mlp_fwd(save_activations = 1) on the same workspace. */
int nmx_mlp_bwd(nmx_mlp_plan* plan, void* workspace, const float* params, const float* d_out, float* d_params,
                int64_t P, void* stream);
/* same, and additionally the gradient w.r.t. the (already encoded, enc_kind 0) POSITION inputs:
 * d_input fp32 [P, cols], cols = nmx_mlp_input_grad_cols(plan) >= in_pos, columns >= in_pos are zero
 * (= dY_0 W_0[:, :in_pos] + dY_skip W_skip[:, :in_pos]).  This is what lets a learnable encoder in front of the MLP
 * -- the hash grid (encoding/multi_hash.py) -- receive its gradient, as the reference's autograd would provide. */
int nmx_mlp_bwd_input(nmx_mlp_plan* plan, void* workspace, const float* params, const float* d_out, float* d_params,
                      float* d_input, int64_t P, void* stream);
/* row length of the d_input the NEXT nmx_mlp_bwd_input on this plan writes: in_pos rounded up to 64 on the per-layer /
 * chain paths, exactly in_pos after a saving forward through the fused width-64 kernel (width 64, no view-dir head, no
 * skip connection, in_pos 32 or 64, enc_kind 0), whose backward writes the compact gradient directly. */
int nmx_mlp_input_grad_cols(const nmx_mlp_plan* plan);

/* generic bf16 GEMM building block on tcgen05 (exposed for unit tests / profiling):
 * D[M,N] = act(A[M,K] * B[N,K]^T + bias[N]);  A,B bf16 row-major (K-major), D bf16 or fp32. */
int nmx_gemm_bf16(const void* A, const void* Bm, const float* bias, void* D, int64_t M, int N, int K,
                  int relu, int d_is_fp32, void* stream);
/* the same product on CTA pairs (tcgen05 cta_group::2, M = 256 tiles, each CTA stages half of the weight slab):
 * D[M, 256] fp32 = A[M, K] * B[256, K]^T; K % 64 == 0; max_pairs caps the number of clusters (0 = all SM pairs). */
int nmx_gemm_pair_bf16(const void* A, const void* Bm, float* D, int64_t M, int K, int max_pairs, void* stream);
/* dW[M,N] (fp32, accumulated: caller zeroes) += dY[P,M]^T * X[P,N]; bf16 row-major inputs, read as MN-major
 * UMMA operands (no transposes); M % 64 == 0, N % 64 == 0, N <= 256.  Optional db[M] (fp32, accumulated) += column
 * sums of dY (the bias gradient), computed by one extra N=16 MMA against a constant all-ones operand tile. */
int nmx_wgrad_bf16(const void* dY, const void* X, float* dW, float* db, int64_t P, int M, int N, void* stream);
/* out[N] (fp32, accumulated) += column sums of Y[P,N] bf16 (bias gradients). */
int nmx_colsum_bf16(const void* Y, float* out, int64_t P, int N, void* stream);

/* live per-kernel timing for bench.py's roofline: when enabled, every tensor-core GEMM launch is bracketed by CUDA
 * events on its own stream.  kind 0 = layer GEMM (forward + dgrad of nets the fused chain does not cover), 1 = wgrad,
 * 2 = fused chain forward (inference), 3 = fused chain forward (training: saves activations), 4 = fused chain backward
 * (data gradients).  flops are the PADDED flops launched. */
int nmx_profile_enable(int on);
int nmx_profile_report(int kind, double* total_ms, double* total_flops, int64_t* launches);

/* ---- diagnostics (not on the product path; used by scripts/ to calibrate the roofline) ------------------------
 * nmx_diag_mma_rate: every CTA issues `iters` back-to-back M=128 x N x K=16 bf16 tcgen05 MMAs on resident operands;
 * out[2*cta] = SM clocks, out[2*cta+1] = nanoseconds.  mode bits 0-1: 1 = tcgen05.commit after every 4 MMAs,
 * 2 = commit + an mbarrier poll + fence; bit 2: alternate two 128-column accumulator halves.
 * nmx_chain_trace_read: copies the (clock, ns) event trace the fused MLP chain records for CTA 0 when the environment
 * variable NMX_CHAIN_DBG has bit 2 set. */
int nmx_diag_mma_rate(int N, int iters, int n_slabs, int ctas, long long* out, void* stream, int mode);
/* TMEM -> register read-rate probe (tcgen05.ld.32x32b.x16 / .x32): out[2*cta] = clocks, out[2*cta+1] = bytes read */
int nmx_diag_tmem_ld_rate(int iters, int warps, int mode, int batch, int ctas, long long* out, void* stream);
int nmx_chain_trace_read(long long* out, int n);
/* HBM bandwidth probes (scripts/bw_probe.py -> profiles/r2_bw_probe.txt).  mode 0 cudaMemsetAsync, 1 st.global.v4 fill,
 * 2 bulk-async 1-D stores from shared memory (64 KB tiles, one CTA per SM), 3 ld.global.v4 read, 4 ld/st copy
 * (src -> buf), 5 TMA 2-D tensor stores in the fused MLP chain's own pattern ([rows, 256] bf16 written as 128 x 64
 * boxes), 6 TMA 2-D tile loads (the weight-gradient kernels' read pattern).  bytes % 64 KiB == 0; depth = bulk groups
 * in flight per CTA (modes 2, 5). */
int nmx_diag_bw(int mode, void* buf, const void* src, int64_t bytes, int ctas, int depth, void* stream);
/* byte offsets of the training workspace regions: out[12] = {activation base, x0, h0, h stride, feature, hd, g0, g stride,
 * ghd, relu sign bits, capacity (points), x0 columns}; region offsets are relative to the activation base. */
int nmx_mlp_debug_layout(const nmx_mlp_plan* plan, int64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* NMX_H_ */
