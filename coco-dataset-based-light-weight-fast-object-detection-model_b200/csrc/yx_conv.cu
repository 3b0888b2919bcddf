// K1 — NHWC fp16 implicit-GEMM convolution for sm_100a (tcgen05 + TMEM + TMA), with bias,
// activation and residual-add fused into the epilogue and concat-slice reads/writes expressed
// through the tensor maps.  Replaces BaseConv / Bottleneck of the reference
// (yolox/models/network_blocks.py:73-84,199-205; choijhanyangackr/yolox_infer/models/blocks.py:21-49).
//
// GEMM view:  D[M = pixels, N = Cout] = sum over (tap, cin) A[pixel shifted by tap, cin] * W[cout, tap, cin]
//   * M tile  = TH x TW output pixels of one image (<= 128 rows -> one UMMA M=128 tile; TMEM lane = row)
//   * K chunk = 64 input channels of one filter tap = one 128-byte swizzled smem row per pixel
//   * A operand: one 4-D TMA box (64 ch, TW, TH, 1) per (tap, chunk), shifted by the tap offset;
//     TMA zero-fills outside the image (= conv zero padding) and beyond the channel extent.
//     Stride 2 uses four parity views (even/odd rows x even/odd cols) of the input, so every tap
//     is again a dense box.
//   * B operand: 3-D TMA box (64 ch, 1 tap, BN couts) of the KRSC weight tensor.
//   * accumulators: fp32 in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1.
// Warp roles (192 threads, persistent CTA, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue (TMEM -> regs -> bias/act/residual -> swizzled smem -> TMA store).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "yx_internal.h"
#include "yx_ptx.cuh"

namespace yx {

constexpr int kThreads = 192;
constexpr int kAStageBytes = 128 * 128;  // 128 rows x 64 fp16
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kBarBytes = 256;
constexpr int kSmemTwoCtas = 112 * 1024;  // per-CTA budget that lets two CTAs share an SM

// arithmetic intensity (FLOP / byte of fp16 activation traffic) below which a layer is launched in the
// two-CTAs-per-SM "streaming" shape.  B200 ridge = 1414.9 TF/s / 6527 GB/s = 217 FLOP/B.
// YX_MEM_AI overrides it for experiments (0 = never, 1e9 = always).
static double mem_bound_ai() {
  static double v = -1.0;
  if (v < 0) {
    const char* e = getenv("YX_MEM_AI");
    v = e ? atof(e) : 300.0;
  }
  return v;
}

// Activation on the fp16-rounded conv output, evaluated in fp32 like torch's half kernels (opmath = float).
// Compile-time ACT keeps the epilogue straight-line: the epilogue has ONE warp per scheduler, so it lives on
// instruction-level parallelism across the 16 columns of a TMEM chunk (a runtime switch per element
// serialised it to ~150 cycles/element in the first version — see profiles/r01_conv_tile_trace.txt).
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == YX_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
  if (ACT == YX_ACT_HSWISH) return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
  if (ACT == YX_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == YX_ACT_LRELU) return x > 0.0f ? x : 0.1f * x;
  return x;
}

// optional per-tile timeline (diagnostics): trace[tile_local * 8 + event] = clock64(), CTA 0 only
#define YX_TRACE(ev, tl)                                                                   \
  do {                                                                                     \
    if (p.trace != nullptr && blockIdx.x == 0 && (tl) < 32) p.trace[(tl) * 8 + (ev)] = clock64(); \
  } while (0)

template <int ACT, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 2) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int groups = (p.BN + 63) >> 6;  // 64-channel output groups per tile
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + p.stages * kAStageBytes;
  const uint32_t sStage = sB + p.stages * p.b_stage_bytes;
  const uint32_t sBar = sStage + groups * kAStageBytes;
  // barrier layout (8 bytes each): full[8], empty[8], tmem_full[2], tmem_empty[2], res_full, then tmem ptr
  const uint32_t bar_full = sBar, bar_empty = sBar + 64, bar_tfull = sBar + 128, bar_tempty = sBar + 144;
  const uint32_t bar_res = sBar + 160, tmem_slot = sBar + 168;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4);  // one arrive per epilogue warp
    }
    mbar_init(bar_res, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int n_tiles = p.n_tiles_m * p.n_tiles_n;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int taps = p.ky * p.kx;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int y0 = (r / p.tiles_w) * p.TH, x0 = (r % p.tiles_w) * p.TW;
        const int n0 = nt * p.BN;
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = tap / p.kx, dx = tap % p.kx;
          int mi = 0, cx, cy;
          if (p.stride == 1) {
            cx = x0 + dx - p.pad_x;
            cy = y0 + dy - p.pad_y;
          } else {
            const int pad = p.pad_y;
            // input row 2*y + dy - pad.  For k=3,pad=1: dy=0 -> odd row of cell y-1; dy=1 -> even row
            // of cell y; dy=2 -> odd row of cell y.  k=1 (pad 0): even row/col of cell y.
            const int oy = dy - pad, ox = dx - pad;
            const int py = oy & 1, px = ox & 1;
            mi = py * 2 + px;
            cy = y0 + ((oy - py) >> 1);
            cx = x0 + ((ox - px) >> 1);
          }
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            mbar_expect_tx(bar_full + 8 * s, p.a_box_bytes + p.b_stage_bytes);
            tma_load_4d(sA + s * kAStageBytes, &p.tmA[mi], bar_full + 8 * s, kc * 64, cx, cy, img);
            tma_load_3d(sB + s * p.b_stage_bytes, &p.tmW, bar_full + 8 * s, kc * 64, tap, n0);
          }
        }
        YX_TRACE(0, (tile - (int)blockIdx.x) / (int)gridDim.x);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      uint32_t it = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int nt = tile % p.n_tiles_n;
        const int n0 = nt * p.BN;
        const int bn_cur = min(p.BN, p.cout16 - n0);
        const uint32_t idesc = make_idesc_f16(bn_cur);
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        YX_TRACE(1, t);
        const uint32_t d_tmem = tmem_base + acc * p.acc_stride;
        uint32_t accum = 0;
        for (int tap = 0; tap < taps; ++tap) {
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            const uint32_t s = it % p.stages, ph = (it / p.stages) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            tc_fence_after();
            const int ksteps = min(64, p.cin - kc * 64) >> 4;
            const uint64_t adesc = make_sdesc_sw128(sA + s * kAStageBytes);
            const uint64_t bdesc = make_sdesc_sw128(sB + s * p.b_stage_bytes);
            for (int ks = 0; ks < ksteps; ++ks) {
              // advance 16 fp16 = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr >> 4)
              umma_f16_ss(d_tmem, adesc + 2 * ks, bdesc + 2 * ks, idesc, accum);
              accum = 1;
            }
            umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
          }
        }
        umma_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
        YX_TRACE(2, t);
      }
    }
  } else {
    // ================================ epilogue (warps 2..5) ================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const bool row_valid = row < p.TH * p.TW;
    const bool leader = (warp == 2 && lane == 0);
    uint32_t t = 0, res_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y0 = (r / p.tiles_w) * p.TH, x0 = (r % p.tiles_w) * p.TW;
      const int n0 = nt * p.BN;
      const int bn_cur = min(p.BN, p.cout16 - n0);
      const int groups_cur = (bn_cur + 63) >> 6;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;

      if (HAS_RES && leader) {
        mbar_expect_tx(bar_res, groups_cur * p.a_box_bytes);
        for (int g = 0; g < groups_cur; ++g)
          tma_load_4d(sStage + g * kAStageBytes, &p.tmRes, bar_res, n0 + g * 64, x0, y0, img);
      }
      mbar_wait(bar_tfull + 8 * acc, acc_ph);
      tc_fence_after();
      if (leader) YX_TRACE(3, t);
      if (HAS_RES) {
        mbar_wait(bar_res, res_cnt & 1);
        ++res_cnt;
      }
      const uint32_t taddr = tmem_base + acc * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
      for (int c0 = 0; c0 < bn_cur; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(taddr + c0, v);
        tmem_ld_wait();
        if (row_valid) {
          float bb[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0) + j);
            bb[4 * j] = b4.x; bb[4 * j + 1] = b4.y; bb[4 * j + 2] = b4.z; bb[4 * j + 3] = b4.w;
          }
          // two 16-byte chunks (8 channels each) of this row's 128-byte swizzled staging line
          const uint32_t line = sStage + (c0 >> 6) * kAStageBytes + row * 128;
          const uint32_t a0 = line + ((((c0 & 63) >> 3) ^ (row & 7)) << 4);
          const uint32_t a1 = line + (((((c0 & 63) >> 3) + 1) ^ (row & 7)) << 4);
          uint32_t rr[8];
          if (HAS_RES) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]) : "r"(a0));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]) : "r"(a1));
          }
          uint32_t out[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            // the reference rounds the conv output to fp16 before its (separate) activation kernel
            const __half2 pre = __floats2half2_rn(__uint_as_float(v[2 * i]) + bb[2 * i],
                                                  __uint_as_float(v[2 * i + 1]) + bb[2 * i + 1]);
            const float2 pf = __half22float2(pre);
            float f0 = apply_act<ACT>(pf.x), f1 = apply_act<ACT>(pf.y);
            if (HAS_RES) {  // half + half as torch computes it: exact fp32 sum of the two halves, rounded once
              const float2 af = __half22float2(__floats2half2_rn(f0, f1));
              const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rr[i]));
              f0 = af.x + rf.x;
              f1 = af.y + rf.y;
            }
            const __half2 o = __floats2half2_rn(f0, f1);
            out[i] = *reinterpret_cast<const uint32_t*>(&o);
          }
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(out[0]), "r"(out[1]), "r"(out[2]),
                       "r"(out[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(out[4]), "r"(out[5]), "r"(out[6]),
                       "r"(out[7]) : "memory");
        }
      }
      // accumulator drained -> MMA warp may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      // publish the staged tile to the async proxy and store it
      fence_proxy_async_smem();
      if (leader) YX_TRACE(4, t);
      named_bar_sync(1, 128);
      if (leader) {
        YX_TRACE(5, t);
        for (int g = 0; g < groups_cur; ++g)
          tma_store_4d(&p.tmOut, sStage + g * kAStageBytes, n0 + g * 64, x0, y0, img);
        tma_store_commit();
        tma_store_wait_read0();
        YX_TRACE(6, t);
      }
      named_bar_sync(1, 128);  // staging buffer reusable
    }
    if (leader) tma_store_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// rank-r fp16 tensor map; dims/strides innermost first; strides[0] implied (2 bytes)
static int encode_map(CUtensorMap* m, void* addr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool swizzle128, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return YX_ERR_CUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i];
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, addr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[512];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled(%s) failed: %d addr=%p rank=%d dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] "
             "box=[%u,%u,%u,%u]",
             what, (int)r, addr, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)dims[2], (unsigned long long)(rank > 3 ? dims[3] : 0),
             (unsigned long long)strides_bytes[1], (unsigned long long)strides_bytes[2],
             (unsigned long long)(rank > 3 ? strides_bytes[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0);
    set_error(buf);
    return YX_ERR_CUDA;
  }
  return YX_OK;
}

// NHWC view -> (c, w, h, n) map with box (64, tw, th, 1)
static int encode_view(CUtensorMap* m, void* base, const yx_view& v, int tw, int th, const char* what) {
  uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  uint64_t st[4] = {2, (uint64_t)v.pitch * 2, (uint64_t)v.pitch * 2 * v.w, (uint64_t)v.nstride * 2};
  uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
  return encode_map(m, static_cast<uint8_t*>(base) + v.offset, 4, dims, st, box, true, what);
}

static void choose_tile(int H, int W, int* th, int* tw) {
  // minimise padded MMA rows; ties -> squarer tile (fewer halo re-reads from L2)
  double best = 1e30;
  int bh = 1, bw = 1;
  for (int w = 1; w <= std::min(W, 128); ++w) {
    int h = std::min(H, 128 / w);
    if (h < 1) continue;
    if (w > 256 || h > 256) continue;
    double tiles = (double)ceil_div(H, h) * ceil_div(W, w);
    double cost = tiles * 1000.0 + std::abs(h - w) * 0.01;
    if (cost < best) { best = cost; bh = h; bw = w; }
  }
  *th = bh;
  *tw = bw;
}

int conv_plan(const yx_op& op, void* base, const void* weights, const void* biases, int num_sms, ConvPlan* out) {
  const yx_view& s = op.src;
  const yx_view& d = op.dst;
  YX_REQUIRE(op.ksize == 1 || op.ksize == 3, "conv ksize must be 1 or 3");
  YX_REQUIRE(op.stride == 1 || op.stride == 2, "conv stride must be 1 or 2");
  YX_REQUIRE(op.cin_pad % 16 == 0 && op.cout_pad % 16 == 0, "cin_pad/cout_pad must be multiples of 16");
  YX_REQUIRE(s.c <= op.cin_pad && s.c % 8 == 0, "src channels must be a multiple of 8 and <= cin_pad");
  YX_REQUIRE(op.aux == 0 || op.aux == 1, "conv aux must be 0 or 1 (row-packed)");
  YX_REQUIRE(d.c % 8 == 0 && d.c <= op.cout_pad, "dst channels must be a multiple of 8 and <= cout_pad");
  YX_REQUIRE(s.pitch % 8 == 0 && d.pitch % 8 == 0 && s.offset % 16 == 0 && d.offset % 16 == 0 && s.nstride % 8 == 0 &&
                 d.nstride % 8 == 0,
             "views must be 16-byte aligned");
  // aux == 1: "row-packed" 3x3 conv over a 16-channel tensor stored with 1 zero column on the left and 3 on
  // the right (the s2d output).  TMA reads it through an OVERLAPPING view (64 channels per pixel, pixel pitch 16),
  // so the 128-byte smem row of pixel x holds pixels x-1..x+2 = the three horizontal taps (+1 ignored): the
  // conv becomes 3 vertical taps with K = 48 instead of 9 taps with K = 16, and every TMA row is a full line.
  const bool rowpack = op.aux == 1;
  if (rowpack)
    YX_REQUIRE(op.ksize == 3 && op.stride == 1 && s.c == 16 && s.pitch == 16 && op.cin_pad == 48 && s.w > 4 &&
                   op.res.c == 0,
               "row-packed conv needs k=3, s=1, a padded 16-channel source and cin_pad = 48");
  const int pad = op.ksize / 2;
  const int Hout = rowpack ? s.h : (s.h + 2 * pad - op.ksize) / op.stride + 1;
  const int Wout = rowpack ? s.w - 4 : (s.w + 2 * pad - op.ksize) / op.stride + 1;
  YX_REQUIRE(d.h == Hout && d.w == Wout && d.n == s.n, "dst spatial dims do not match the conv geometry");
  const bool has_res = op.res.c > 0;
  if (has_res)
    YX_REQUIRE(op.res.h == d.h && op.res.w == d.w && op.res.c == d.c && op.res.n == d.n && op.res.pitch % 8 == 0 &&
                   op.res.offset % 16 == 0 && op.res.nstride % 8 == 0,
               "residual view must match dst");

  ConvPlan pl;
  memset(&pl, 0, sizeof pl);
  ConvParams& p = pl.p;
  p.ksize = op.ksize; p.stride = op.stride; p.act = op.act; p.has_res = has_res;
  p.ky = op.ksize; p.kx = rowpack ? 1 : op.ksize;
  p.pad_y = pad; p.pad_x = rowpack ? 0 : pad;
  p.cin = op.cin_pad;
  p.cout16 = op.cout_pad;
  p.k_chunks = ceil_div(p.cin, 64);
  choose_tile(Hout, Wout, &p.TH, &p.TW);
  p.tiles_h = ceil_div(Hout, p.TH);
  p.tiles_w = ceil_div(Wout, p.TW);
  p.n_tiles_m = d.n * p.tiles_h * p.tiles_w;
  // Algorithmic work decides the launch shape.  Layers below the ridge (HBM-bound: all 1x1 convs at
  // these channel counts, the 48/96-channel 3x3 convs) run TWO CTAs per SM with narrow N tiles so one
  // CTA's epilogue / TMA-store latency hides behind the other's loads; tensor-bound layers keep one CTA
  // per SM with the widest N tile (fewest re-reads of A) and the deepest smem pipeline.
  const double px_out_ = (double)d.n * Hout * Wout;
  const int cin_real = rowpack ? 12 : s.c;
  const double flops_ = 2.0 * px_out_ * d.c * cin_real * op.ksize * op.ksize;
  const double bytes_ = 2.0 * ((double)s.n * s.h * s.w * s.c + px_out_ * d.c * (has_res ? 2 : 1));
  const bool mem_bound = flops_ / bytes_ < mem_bound_ai();
  int ctas_per_sm = 1;
  if (mem_bound) {
    if (p.cout16 <= 128) {
      p.BN = p.cout16;
      p.n_tiles_n = 1;
    } else {
      p.BN = 128;
      p.n_tiles_n = ceil_div(p.cout16, 128);
    }
    ctas_per_sm = 2;
  } else if (p.cout16 <= 256) {
    p.BN = p.cout16;
    p.n_tiles_n = 1;
  } else {
    int nt = ceil_div(p.cout16, 256);
    p.BN = round_up(ceil_div(p.cout16, nt), 64);
    p.n_tiles_n = ceil_div(p.cout16, p.BN);
  }
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.BN) p.tmem_cols <<= 1;
  p.acc_stride = p.tmem_cols / 2;
  p.b_stage_bytes = p.BN * 128;
  p.a_box_bytes = p.TH * p.TW * 128;
  const int groups = ceil_div(p.BN, 64);
  const int fixed = groups * kAStageBytes + kBarBytes + 1024;  // staging + barriers + alignment slack
  const int budget = ctas_per_sm == 2 ? kSmemTwoCtas : kSmemLimit;
  p.stages = std::min(kMaxStages, (budget - fixed) / (kAStageBytes + p.b_stage_bytes));
  YX_REQUIRE(p.stages >= 2, "not enough shared memory for a 2-stage pipeline");
  pl.smem_bytes = fixed + p.stages * (kAStageBytes + p.b_stage_bytes);
  // never let a third CTA (which would stall in tcgen05.alloc) fit on an SM
  if (ctas_per_sm == 2) pl.smem_bytes = std::max(pl.smem_bytes, 80 * 1024);
  pl.grid = std::min(p.n_tiles_m * p.n_tiles_n, ctas_per_sm * num_sms);
  p.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(biases) + op.b_offset);

  int rc;
  if (rowpack) {
    uint64_t dims[4] = {64, (uint64_t)Wout, (uint64_t)s.h, (uint64_t)s.n};
    uint64_t st[4] = {2, 32, (uint64_t)s.w * 32, (uint64_t)s.nstride * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if ((rc = encode_map(&p.tmA[0], static_cast<uint8_t*>(base) + s.offset, 4, dims, st, box, true, "A-rowpack")) != YX_OK)
      return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else if (op.stride == 1) {
    yx_view sv = s;
    if ((rc = encode_view(&p.tmA[0], base, sv, p.TW, p.TH, "A")) != YX_OK) return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else {
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        // parity view: rows py, py+2, ... and cols px, px+2, ...
        uint64_t dims[4] = {(uint64_t)s.c, (uint64_t)((s.w - px + 1) / 2), (uint64_t)((s.h - py + 1) / 2), (uint64_t)s.n};
        uint64_t st[4] = {2, (uint64_t)s.pitch * 4, (uint64_t)s.pitch * 4 * s.w, (uint64_t)s.nstride * 2};
        uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
        uint8_t* addr = static_cast<uint8_t*>(base) + s.offset + ((int64_t)py * s.w + px) * s.pitch * 2;
        YX_REQUIRE(dims[1] > 0 && dims[2] > 0, "stride-2 conv needs H,W >= 2");
        if ((rc = encode_map(&p.tmA[py * 2 + px], addr, 4, dims, st, box, true, "A-parity")) != YX_OK) return rc;
      }
  }
  {
    const int taps = p.ky * p.kx;
    uint64_t dims[3] = {(uint64_t)op.cin_pad, (uint64_t)taps, (uint64_t)op.cout_pad};
    uint64_t st[3] = {2, (uint64_t)op.cin_pad * 2, (uint64_t)op.cin_pad * 2 * taps};
    uint32_t box[3] = {64, 1, (uint32_t)p.BN};
    uint8_t* addr = const_cast<uint8_t*>(static_cast<const uint8_t*>(weights)) + op.w_offset;
    YX_REQUIRE(op.w_offset % 16 == 0, "weight offset must be 16-byte aligned");
    if ((rc = encode_map(&p.tmW, addr, 3, dims, st, box, true, "W")) != YX_OK) return rc;
  }
  if ((rc = encode_view(&p.tmOut, base, d, p.TW, p.TH, "out")) != YX_OK) return rc;
  if (has_res) {
    if ((rc = encode_view(&p.tmRes, base, op.res, p.TW, p.TH, "res")) != YX_OK) return rc;
  } else {
    p.tmRes = p.tmOut;
  }
  const double px_out = (double)d.n * Hout * Wout;
  pl.flops = 2.0 * px_out * d.c * cin_real * op.ksize * op.ksize;
  pl.bytes = 2.0 * ((double)s.n * s.h * s.w * s.c + px_out * d.c * (has_res ? 2 : 1)) +
             2.0 * (double)d.c * cin_real * op.ksize * op.ksize;
  *out = pl;
  return YX_OK;
}

template <int ACT, bool HAS_RES>
static int launch_variant(const ConvPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    YX_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<ACT, HAS_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  conv_igemm_kernel<ACT, HAS_RES><<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.p);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

template <int ACT>
static int launch_act(const ConvPlan& plan, cudaStream_t stream) {
  return plan.p.has_res ? launch_variant<ACT, true>(plan, stream) : launch_variant<ACT, false>(plan, stream);
}

int conv_launch(const ConvPlan& plan, cudaStream_t stream) {
  switch (plan.p.act) {
    case YX_ACT_NONE: return launch_act<YX_ACT_NONE>(plan, stream);
    case YX_ACT_SILU: return launch_act<YX_ACT_SILU>(plan, stream);
    case YX_ACT_HSWISH: return launch_act<YX_ACT_HSWISH>(plan, stream);
    case YX_ACT_RELU: return launch_act<YX_ACT_RELU>(plan, stream);
    case YX_ACT_LRELU: return launch_act<YX_ACT_LRELU>(plan, stream);
    default: set_error("unknown activation code"); return YX_ERR_INVALID;
  }
}

}  // namespace yx
