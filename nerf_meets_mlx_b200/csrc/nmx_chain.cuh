// Host/device description of the fused MLP forward chain (nmx_chain.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nmx {

constexpr int kMaxChainLayers = 10;
constexpr int kSrcPos = 4;  // A slab = encoded-position chunk of the input tile
constexpr int kSrcDir = 5;  // A slab = encoded-view-dir chunk of the input tile (0..3 = activation chunk index)

struct ChainLayerDesc {
  int n_slabs;     // K / 64
  int src[5];      // A-operand source of every K slab
  uint32_t src_packed;  // the same, 4 bits per slab (filled by launch_chain_fwd)
  // issue schedule of the layer's pieces (filled by launch_chain_fwd; bit layout documented there)
  uint64_t piece_tab[10];
  int n_pieces;
  int N;           // 256 or 128
  int relu;
  int bias_off;    // float offset into params
  int feeds_next;  // its output chunks are the A operand of a later layer (MMA waits on act_ready)
  int save_kind;   // 0 none, 1 = [rows, 256] activation store (h_l / feature), 2 = [P, 128] store (hd)
  int save_row0;   // first row of this layer's region in the activation store
  // ReLU sign bits (1 bit per activation, 32 B per point and layer: bit e / 16+e of word w = columns 32w+2e / 32w+2e+1).
  // forward (training): row offset of this layer's slot in ChainParams::bits, or -1 (layer has no ReLU / not saved);
  // backward: slot of the activation that masks this layer's output (epi >= 1)
  int bits_row0;
  // backward chain only: epilogue kind (0 plain, 1 ReLU mask, 2 mask + alpha rank-1 term)
  int epi;
};

struct ChainParams {
  int n_layers;
  ChainLayerDesc L[kMaxChainLayers];
  int P, save;
  const float* params;
  float* out;
  int out_cols;
  int head7_layer, head7_n, head7_w_off, head7_b_off;  // alpha (n=1) or output_linear (n=out_ch) from h_{D-1}
  int rgb_layer, rgb_w_off, rgb_b_off;                 // rgb head from the dir layer's output, or rgb_layer = -1
  int uses_dir, x0_dir_col;
  int pos_last_layer, pos_prefetch_layer, dir_layer;
  // ReLU sign-bit store [slots * cap rows][8] uint32: written by the training forward, read by the backward chain
  uint32_t* bits;
  int hd_bits_row0;   // backward: slot of the dir-layer activation hd (masks step A)
  // backward chain only: d_raw [P, 4] fp32 (d_rgb, d_sigma)
  const float* d_out;
  // forward, fused input encoding (enc_fused != 0): a dedicated warp evaluates pos = o + z d and the Embedder PE
  // (10 position bands, models/embedding.py:35-71) for the next tile straight into the shared-memory input chunks;
  // the per-ray view-dir PE (dir_pe [rays, 64] bf16) is gathered.  p0 / b0: first point / ray of this launch.
  int enc_fused;
  const float* rays;
  int ray_stride;
  const float* z;
  const void* dir_pe;
  long long p0, b0;
  int n_per_ray;
  int cap;       // rows per slot of the activation store (experiment NMX_CHAIN_DBG bit 5: chunk-major store layout)
  int max_ctas;  // 0 = one CTA per SM; smaller values leave SMs to a kernel running concurrently on another stream
  int dbg;  // experiments only (NMX_CHAIN_DBG): bit 0 = no weight TMA traffic, bit 1 = epilogue math skipped
};

struct ChainMaps {
  CUtensorMap w[kMaxChainLayers];  // bf16 weights [N_l, K_l] (padded K), box 64 x 128
  CUtensorMap x0;                  // encoded input [P, pos_pad + dir_pad], box 64 x 128
  CUtensorMap save;                // activation store [(D+1) * cap, 256], box 64 x 128
  CUtensorMap hd;                  // [P, 128], box 64 x 128
};

int launch_chain_fwd(const ChainMaps& maps, const ChainParams& prm, cudaStream_t stream);
int launch_chain_bwd(const ChainMaps& maps, const ChainParams& prm, cudaStream_t stream);

// Inference forward of the 8 x 256 view-dir net on CTA pairs with two tiles in ping-pong (nmx_chain2.cu).
struct Chain2Launch {
  long long P;                 // points of this launch
  const float* params;         // packed fp32 parameters
  float* out;                  // [P, 4] fp32
  const void* w_ptr[10];       // bf16 weights [N_l, K_l] of the 8 trunk layers, feature layer, dir layer
  int w_k[10];                 // padded K (row length) of each
  int bias_off[10];
  int alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off;
  const float* rays; int ray_stride; const float* z;
  long long p0; int n_per_ray; int n_freqs_dir;
  int dir_w_off, dir_ldw;      // fp32 dir-layer weight [128, dir_ldw]; its columns 256.. multiply PE(dir)
  float* dir_bias;             // scratch [rays of this launch, 128] fp32
};
int launch_chain2(const Chain2Launch& a, cudaStream_t stream);

// Per-launch preparation of the pair chains in one kernel (nmx_chain2.cu `ray_prep_kernel`): the constants block
// (biases, w_alpha, w_rgb: 3328 floats), the per-ray view-dir term of the dir layer [rays, 128] fp32 and, when dir_pe is
// non-null, the bf16 per-ray PE(dir) table [rays, 64].
struct RayPrep {
  const float* params;
  int bias_off[10];
  int alpha_w_off, rgb_w_off;
  const float* rays; int ray_stride;
  long long b0, B;             // first ray / number of rays
  int n_freqs_dir;
  int dir_w_off, dir_ldw;
  float* consts;
  float* dir_bias;
  void* dir_pe;
};
int launch_ray_prep(const RayPrep& rp, cudaStream_t stream);

// Training forward of the same net on CTA pairs (nmx_chain2t.cu): saves h_0 .. h_7, hd, the ReLU sign bits and the
// encoded input tile X0, in the layouts the one-tile training chain (nmx_chain.cu) writes.
struct Chain2TrainLaunch {
  long long P;                 // points (all rays of the pass: p0 = 0)
  const float* params;
  float* out;                  // [P, 4] fp32
  const void* w_ptr[10];
  int w_k[10];
  int bias_off[10];
  int alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off;
  const float* rays; int ray_stride; const float* z;
  int n_per_ray; int in_dir;   // in_dir = encoded view-dir channels (27)
  int dir_w_off, dir_ldw;
  void* dir_pe;                // [rays, 64] bf16 per-ray view-dir PE table (written by the launch's ray_prep_kernel)
  int n_freqs_dir;
  float* scratch;              // chain2_train_scratch_bytes(rays): constants block + per-ray dir-layer term
  void* save_base; long long save_rows; long long cap;  // activation store [(D + 1) * cap, 256] bf16
  void* hd;                    // [P, 128] bf16
  void* x0;                    // [P, 128] bf16
  uint32_t* bits;              // sign-bit store [(D + 1) * cap][8] uint32
};
// Backward data-gradient chain on CTA pairs (nmx_chain2t.cu): layer order d_feature, dY_7, dY_6 .. dY_0.
struct Chain2BwdLaunch {
  long long P;
  const float* params;         // packed fp32 parameters (w_alpha, w_rgb)
  const float* d_out;          // [P, 4] fp32
  const void* w_ptr[9];        // transposed bf16 weights [256, K]: W_dir[:, :W]^T (K = 128), W_feat^T, W_7^T .. W_1^T (h parts)
  int w_k[9];
  int alpha_w_off, rgb_w_off;
  const uint32_t* bits;        // the forward's sign-bit store
  int bits_word_major;         // 1: 4 KB tiles are [8 words][128 rows] (written by the pair forward), 0: [128 rows][8 words]
  void* save_base; long long save_rows; long long cap;  // dY store [(D + 1) * cap, 256]
  void* ghd;                   // d_hd [P, 128]
};
int launch_chain2_bwd(const Chain2BwdLaunch& a, cudaStream_t stream);
int64_t chain2_train_scratch_bytes(int64_t n_rays);
int launch_chain2_train(const Chain2TrainLaunch& a, cudaStream_t stream);

}  // namespace nmx
