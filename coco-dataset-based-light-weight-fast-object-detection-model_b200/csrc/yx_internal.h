// Internal (non-ABI) declarations shared by the translation units of libyolox_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

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
  CUtensorMap tmA[4];  // activation source(s): [0] for stride 1, [py*2+px] parity views for stride 2
  CUtensorMap tmW;     // weights  (cin_pad, k*k, cout_pad), box (64, 1, BN)
  CUtensorMap tmOut;   // output   (c, W, H, N),             box (64, TW, TH, 1)
  CUtensorMap tmRes;   // residual, same geometry as tmOut
  const float* bias;
  int ksize, stride, act, has_res;
  int ky, kx, pad_y, pad_x;     // tap grid actually iterated (3x1 for the row-packed stem conv)
  int TH, TW, tiles_h, tiles_w;  // spatial tiling of the output
  int n_tiles_m, n_tiles_n, BN;
  int cin, cout16;
  int k_chunks;
  int stages, b_stage_bytes, a_box_bytes;
  int halo, mh, stages_a, a_stage_bytes;  // conv3x3_halo_kernel: stacked 128-pixel halves, halo ring
  int tmem_cols, acc_stride;
  int noload;                 // diagnostics: skip TMA loads after the ring is primed (results are garbage)
  long long* trace;           // diagnostics: per-tile timeline of CTA 0 (nullptr = off)  // TMEM columns allocated (power of two) and offset of accumulator 1
};

struct ConvPlan {
  ConvParams p;
  int grid;
  int smem_bytes;
  double flops;  // algorithmic: 2*N*Hout*Wout*Cout*Cin*k*k (real channel counts)
  double bytes;  // algorithmic: fp16 in + out (+ residual) + weights
};

int conv_plan(const yx_op& op, void* base, const void* weights, const void* biases, int num_sms, ConvPlan* out);
int conv_launch(const ConvPlan& plan, cudaStream_t stream);

// ------------------------------------------------------------------ aux ops (yx_aux.cu)
int s2d_launch(const void* image, int image_dtype, int aux, int B, int H, int W, float scale, float shift,
               void* base, const yx_view& dst, cudaStream_t stream);
int spp_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream);
int upsample_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream);
int dwconv_launch(void* base, const yx_op& op, const void* weights, const void* biases, cudaStream_t stream);

}  // namespace yx
