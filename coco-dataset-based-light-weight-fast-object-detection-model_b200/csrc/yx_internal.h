// Internal (non-ABI) declarations shared by the translation units of libyolox_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/yolox_b200.h"

namespace yx {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define YX_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return ::yx::cuda_fail(_e, #call, __FILE__, __LINE__);  \
  } while (0)

#define YX_REQUIRE(cond, msg)                                   \
  do {                                                          \
    if (!(cond)) {                                              \
      ::yx::set_error(std::string("invalid argument: ") + msg); \
      return YX_ERR_INVALID;                                    \
    }                                                           \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ------------------------------------------------------------------ conv (yx_conv.cu)
struct ConvParams {
  CUtensorMap tmA[4];  // activation source(s): [0] for stride 1 / halo, [py*2+px] parity views for stride 2
  CUtensorMap tmW;     // weights  (cin_pad, cout_pad, taps) [memory order is cout, tap, cin], box (64, BN, b_taps)
  CUtensorMap tmOut;   // output   (c, W, H, N),             box (64, TW, TH or 16, 1)
  CUtensorMap tmRes;   // residual, same geometry as tmOut
  CUtensorMap tmUp;    // fused upsample source: (c, dup_x, w/2, dup_y, n*h/2) with stride-0 dup dims, box (64, 2, TW/2, 2, TH/2)
  int up_chunks, up_h;  // leading K chunks that come from tmUp; rows per image of the low-res tensor
  const float* bias;
  int ksize, stride, act, has_res;
  int ky, kx, pad_y, pad_x;      // tap grid actually iterated (3x1 for the row-packed stem conv)
  int TH, TW, tiles_h, tiles_w;  // spatial tiling of the output (halo: TH = 16*mh, TW = 8)
  int n_tiles_m, n_tiles_n, BN;
  int step_nt, step_x, step_y, step_img;  // mixed-radix digits of the persistent-tile step (= grid size)
  int cin, cout16;
  int k_chunks;
  int halo, mh;                  // halo variant: mh stacked 128-pixel halves per CTA
  int pair;                      // CTA-pair mode (cluster of 2, cta_group::2 MMAs of M = 256)
  int rowpack;                   // row-packed stem (aux == 1); with halo: vertical halo of 8-pixel rows, 3 taps
  int stages_a, a_stage_bytes, a_box_bytes;  // A ring
  int b_slots, b_stage_bytes, b_resident;    // B ring (or the whole weight tile, loaded once)
  int b_taps;                    // filter taps per B ring stage (halo + streamed weights: 3 = one filter row per stage, one TMA
                                 // load, one wait and one commit per three taps; otherwise 1)
  int shared_ring;               // generic + streamed weights: A and B share one full/empty barrier pair per stage
  int stage_bufs, out_box_bytes, bias_bytes; // output staging buffers (1 or 2), bytes per 64-ch residual box
  int epi_groups;                // epilogue warpgroups (1 or 2)
  int epi_alt;                   // two groups: 0 = split every tile's columns, 1 = alternate tiles (group g <-> accumulator g)
  int w3_role;                   // warp 3: 0 idle, 1 second A producer, 2 second B producer
  int w2_role;                   // warp 2 (B producer): 1 = joins the A producers once its resident weights are loaded
  int tmem_cols, acc_stride;     // TMEM columns allocated (power of two) and columns per accumulator
  // 2:4 sparse tensor-core variant (SP): the weights are the sparse A operand (M = 128 couts per MMA, compressed rows of
  // 64 bytes per 64-channel chunk, SWIZZLE_64B), the pixels the B operand (N = 128), the accumulator holds couts on the
  // TMEM lanes and pixels on the columns; metadata columns sit behind the two accumulators
  int sp;                        // 1: sparse variant
  const uint32_t* sp_meta;       // [n_tiles_n][sp_cols_per_tile][128] metadata words (device)
  int sp_cols_per_tile;          // taps * (cin / 32): one column per (tap, K = 32 step)
  int sp_meta_col0;              // first TMEM column of the metadata
  // image-fed row-packed stem (IMG): the A operand is built by the kernel from the caller's NCHW image (set per launch)
  CUtensorMap tmImg;             // raw image (W, H, 3, B), box (32 px, 2(TH+2) rows, 3, 1): encoded per launch
  int img_fused;                 // 1: IMG variant
  const void* img;               // [B,3,img_h,img_w] of img_dtype
  int img_dtype, img_h, img_w, img_order, img_affine;   // order 1 = pixel_unshuffle, 0 = Focus; affine: x*scale+shift on
  float img_scale, img_shift;
  int diag;                      // experiments (YX_CONV_DIAG): 1 no epilogue work, 2 no A loads, 4 no MMAs, 16 general MMA-issue loops only (valid results)
  long long* trace;              // diagnostics: per-tile timeline of CTA 0 (nullptr = off)
};

// Launch-shape knobs of one conv (chosen by default_tune() or by yx_engine_tune()).
struct ConvTune {
  int variant;      // 1 generic (one A box per tap), 2 halo (3x3/s1: one halo box per chunk, 9 descriptors)
  int bn;           // N tile
  int ctas;         // CTAs per SM (1 or 2)
  int mh;           // halo: stacked halves (1 or 2)
  int epi_groups;   // 1 or 2
  int stage_bufs;   // 1 or 2
  int w3;           // 0: no second producer warp, otherwise pick automatically
  int no_resident;  // 1: never keep the weights resident (experiments)
  int pair;         // 1: CTA-pair mode (cta_group::2)
  int sparse;       // 1: 2:4 sparse tensor-core variant (weights = sparse A operand of tcgen05.mma.sp)
  int epi_alt;      // 1: the two epilogue groups alternate tiles instead of splitting each tile's columns
};

struct ConvPlan {
  ConvParams p;
  ConvTune tune;
  int grid, threads;
  int store_only;  // launch an in-place residual conv (p.has_res == 2) with plain stores (tuning / profiling only)
  int no_pdl;      // launch without the programmatic-dependent-launch attribute (an op that waits on another stream's event)
  int smem_bytes;
  double flops;  // algorithmic: 2*N*Hout*Wout*Cout*Cin*k*k (real channel counts)
  double bytes;  // algorithmic: fp16 in + out (+ residual) + weights
  char desc[160];
};

// A conv's 2:4-packed weights (yx_sparse.cu), when its mask is compliant: device pointers owned by the engine.
struct SparseWeights {
  const void* wc = nullptr;        // compressed [round_up(cout_pad,128)][taps][cin_pad/2] fp16
  const uint32_t* meta = nullptr;  // [n_mt][taps * cin_pad / 32][128]
};
int sparse_pack(const void* weights_krsc, int cout_pad, int taps, int cin_pad, void* wc_out, void* meta_out, int* scratch_dev,
                int* compliant, cudaStream_t stream);
bool sparse_shape_ok(const yx_op& op);   // geometry the sparse variant supports (before looking at the weights)

int conv_plan(const yx_op& op, void* base, const void* weights, const void* biases, int num_sms, const ConvTune* tune,
              ConvPlan* out, const SparseWeights* sp = nullptr);
void conv_candidates(const yx_op& op, std::vector<ConvTune>* out, bool sparse_ok = false);
int conv_launch(const ConvPlan& plan, cudaStream_t stream);
int conv_bind_image(ConvPlan* plan, const void* image, int image_dtype, float scale, float shift);   // image-fed stem: per launch

// ------------------------------------------------------------------ aux ops (yx_aux.cu)
int s2d_launch(const void* image, int image_dtype, int aux, int B, int H, int W, float scale, float shift,
               void* base, const yx_view& dst, cudaStream_t stream);
int spp_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream);
int upsample_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream);
int dwconv_launch(void* base, const yx_op& op, const void* weights, const void* biases, cudaStream_t stream);
int view_gather(void* base, const yx_view& v, void* out_contiguous, cudaStream_t stream);
// max |view - ref| in fp16 rounding steps at the larger operand's magnitude (>= floor_mag), as the bits of a float
int view_max_diff(void* base, const yx_view& v, const void* ref_contiguous, float floor_mag, unsigned int* out_bits,
                  cudaStream_t stream);

}  // namespace yx
