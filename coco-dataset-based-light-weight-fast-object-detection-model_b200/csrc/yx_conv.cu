// K1 — NHWC fp16 implicit-GEMM convolution for sm_100a (tcgen05 + TMEM + TMA), with bias,
// activation and residual-add fused into the epilogue and concat-slice reads/writes expressed
// through the tensor maps.  Replaces BaseConv / Bottleneck of the reference
// (yolox/models/network_blocks.py:73-84,199-205; choijhanyangackr/yolox_infer/models/blocks.py:21-49).
//
// GEMM view:  D[M = pixels, N = Cout] = sum over (tap, cin) A[pixel shifted by tap, cin] * W[cout, tap, cin]
//   * M tile  = 128 output pixels of one image (TMEM lane = pixel); MH such tiles may be stacked per CTA
//   * K chunk = 64 input channels of one filter tap = one 128-byte swizzled smem row per pixel
//   * accumulators: fp32 in TMEM, double buffered, so the epilogue of tile i overlaps the MMAs of tile i+1
//
// Two operand-A strategies share one kernel (template HALO):
//   generic  one 4-D TMA box (64 ch, TW, TH, 1) per (tap, chunk), shifted by the tap offset; TMA zero-fills
//            outside the image (= conv padding) and beyond the channel extent.  Stride 2 uses four parity views.
//   halo     3x3 / stride 1: ONE (TH+2)x(8+2) halo box per chunk, the nine taps are nine smem DESCRIPTORS into
//            it (start = base + (dy*10+dx)*128 B, SBO = 1280 B).  9x fewer activation bytes through L2->smem.
//
// What the round-1 probes showed (profiles/r01_tma_issue_probe.txt): ONE thread can issue only about one tiled TMA
// load per 400-600 cycles (mbarrier wait + expect_tx + UTMALDG are each long-latency and serialise in a single
// thread), independent of the box size; more producer WARPS scale linearly up to ~72 B/clk/SM.  Hence:
//   * A and B (weights) have their own rings and their own producer warps (warp 0: A, warp 2: B, warp 3: a second
//     producer for whichever operand has more loads);
//   * B is kept RESIDENT in shared memory for the CTA's lifetime whenever the layer's whole weight tile fits
//     (1x1 convs and small 3x3 convs), which removes half of the loads of the HBM-bound layers;
//   * the output staging buffer is double buffered so the TMA store of tile i drains while tile i+1 is converted,
//     and the conversion can be spread over two epilogue warpgroups.
// Warp roles: 0 = A producer, 1 = MMA issuer, 2 = B producer (+TMEM alloc), 3 = second producer,
//             4..7 = epilogue group 0, 8..11 = epilogue group 1 (optional).
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "yx_internal.h"
#include "yx_ptx.cuh"

namespace yx {

constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 fp16: one generic A stage / one staging group
constexpr int kSmemLimit = 227 * 1024;
constexpr int kBarBytes = 1024;
constexpr int kSmemTwoCtas = 113 * 1024;  // per-CTA budget that lets two CTAs share an SM
constexpr int kMaxARing = 8, kMaxBRing = 32;
constexpr int kHaloW = 10;  // halo row width in pixels (TW = 8)

// arithmetic intensity (FLOP / byte of fp16 activation traffic) below which the DEFAULT (untuned) launch shape is
// the two-CTAs-per-SM streaming one.  B200 ridge = 1414.9 TF/s / 6527 GB/s = 217 FLOP/B.
static double mem_bound_ai() {
  static double v = -1.0;
  if (v < 0) {
    const char* e = getenv("YX_MEM_AI");
    v = e ? atof(e) : 300.0;
  }
  return v;
}

// Activation on the fp16-rounded conv output, evaluated in fp32 like torch's half kernels (opmath = float).
// Compile-time ACT keeps the epilogue straight-line (profiles/r01_conv_tile_trace.txt).
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == YX_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
  // x * relu6(x + 3) / 6 == x * sat(x / 6 + 0.5): FFMA.SAT + FMUL instead of five instructions.  The epilogue is
  // instruction-issue bound (~7 SASS instructions per output element made every small-channel HBM-bound layer wait for
  // it); the two forms differ by an fp32 ulp before the result is rounded to fp16.
  if (ACT == YX_ACT_HSWISH) return x * __saturatef(fmaf(x, 1.0f / 6.0f, 0.5f));
  if (ACT == YX_ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == YX_ACT_LRELU) return x > 0.0f ? x : 0.1f * x;
  return x;
}

// 16 accumulator columns of one row: +bias -> activation (fp32) -> (+ residual already sitting
// in the staging line) -> fp16 -> two 16-byte chunks of the row's 128-byte-swizzled staging line.
template <int ACT, bool HAS_RES>
__device__ __forceinline__ void convert16(const uint32_t (&v)[16], int c0, int row, uint32_t sStage, uint32_t sBiasTile) {
  float bb[16];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(bb[4 * j]), "=f"(bb[4 * j + 1]), "=f"(bb[4 * j + 2]), "=f"(bb[4 * j + 3])
                 : "r"(sBiasTile + (c0 + 4 * j) * 4));
  const uint32_t line = sStage + (c0 >> 6) * kTileBytes + row * 128;
  const uint32_t a0 = line + ((((c0 & 63) >> 3) ^ (row & 7)) << 4);
  const uint32_t a1 = line + (((((c0 & 63) >> 3) + 1) ^ (row & 7)) << 4);
  uint32_t rr[8];
  if (HAS_RES) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]) : "r"(a0));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]) : "r"(a1));
  }
  uint32_t out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    // The reference rounds the conv output to fp16 before its (separate) activation kernel; here the activation is applied
    // to the fp32 sum and the result is rounded ONCE: closer to the exact value than the reference's two roundings (the two
    // can differ by one fp16 ulp), and three instructions fewer per pair in an instruction-issue-bound epilogue.
    float f0 = apply_act<ACT>(__uint_as_float(v[2 * i]) + bb[2 * i]), f1 = apply_act<ACT>(__uint_as_float(v[2 * i + 1]) + bb[2 * i + 1]);
    if (HAS_RES) {  // half + half as torch computes it: exact fp32 sum of the two halves, rounded once
      const float2 af = __half22float2(__floats2half2_rn(f0, f1));
      const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rr[i]));
      f0 = af.x + rf.x;
      f1 = af.y + rf.y;
    }
    const __half2 o = __floats2half2_rn(f0, f1);
    out[i] = *reinterpret_cast<const uint32_t*>(&o);
  }
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(out[0]), "r"(out[1]), "r"(out[2]), "r"(out[3]) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(out[4]), "r"(out[5]), "r"(out[6]), "r"(out[7]) : "memory");
}

// Drains this warp's share of one 128-lane accumulator.  The 16-column chunks of the tile are dealt round-robin to
// the `groups` epilogue warpgroups; each warp keeps two TMEM loads in flight per wait.
template <int ACT, bool HAS_RES>
__device__ __forceinline__ void epilogue_convert(uint32_t taddr, int bn_cur, int row, bool row_valid, uint32_t sStage,
                                                 uint32_t sBiasTile, int group, int groups) {
  const int nch = bn_cur >> 4;
  for (int j = group; j < nch; j += 2 * groups) {
    const int j2 = j + groups;
    const bool two = j2 < nch;  // warp-uniform
    uint32_t v0[16], v1[16];
    tmem_ld_32x32b_x16(taddr + j * 16, v0);
    if (two) tmem_ld_32x32b_x16(taddr + j2 * 16, v1);
    tmem_ld_wait();
    if (row_valid) {
      convert16<ACT, HAS_RES>(v0, j * 16, row, sStage, sBiasTile);
      if (two) convert16<ACT, HAS_RES>(v1, j2 * 16, row, sStage, sBiasTile);
    }
  }
}

// ---- sparse (transposed) epilogue: the accumulator holds output CHANNELS on the TMEM lanes and the tile's PIXELS on the
// columns, so a thread owns one channel: bias is one register, and 16 loaded columns are 16 pixels of that channel.  The
// staging tile keeps the dense layout (one 128-byte swizzled row per pixel and 64-channel group), written with 2-byte
// stores: the 32 lanes of a warp are 32 consecutive channels of one pixel = 64 contiguous bytes, conflict-free.
template <int ACT, bool HAS_RES>
__device__ __forceinline__ void convert16_sp(const uint32_t (&v)[16], int p0, float bias, uint32_t cbase, uint32_t cchunk) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t addr = cbase + (p0 + i) * 128 + ((cchunk ^ (uint32_t)(i & 7)) << 4);   // p0 is a multiple of 16
    float f = apply_act<ACT>(__uint_as_float(v[i]) + bias);
    if (HAS_RES) {  // half + half as torch computes it (see convert16)
      unsigned short r;
      asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(addr));
      f = __half2float(__float2half_rn(f)) + __half2float(__ushort_as_half(r));
    }
    const unsigned short o = __half_as_ushort(__float2half_rn(f));
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(o) : "memory");
  }
}
template <int ACT, bool HAS_RES>
__device__ __forceinline__ void epilogue_convert_sp(uint32_t taddr, int npix, int ch, bool ch_valid, float bias, uint32_t sStage,
                                                    int group, int groups) {
  const uint32_t cbase = sStage + (ch >> 6) * kTileBytes + (ch & 7) * 2;
  const uint32_t cchunk = (uint32_t)(ch & 63) >> 3;
  const int nch = (npix + 15) >> 4;
  for (int j = group; j < nch; j += 2 * groups) {
    const int j2 = j + groups;
    const bool two = j2 < nch;  // warp-uniform
    uint32_t v0[16], v1[16];
    tmem_ld_32x32b_x16(taddr + j * 16, v0);
    if (two) tmem_ld_32x32b_x16(taddr + j2 * 16, v1);
    tmem_ld_wait();
    if (ch_valid) {
      convert16_sp<ACT, HAS_RES>(v0, j * 16, bias, cbase, cchunk);
      if (two) convert16_sp<ACT, HAS_RES>(v1, j2 * 16, bias, cbase, cchunk);
    }
  }
}

// optional per-tile timeline (diagnostics): trace[tile_local * 8 + event] = clock64(), CTA 0 only
#define YX_TRACE(ev, tl)                                                                          \
  do {                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && (tl) < 32) p.trace[(tl) * 8 + (ev)] = clock64(); \
  } while (0)

// mbarrier wait that, in trace mode, also accumulates the cycles this role spent blocked (who is the bottleneck?)
__device__ __forceinline__ void mbar_wait_acc(uint32_t bar, uint32_t parity, bool tracing, long long& acc) {
  if (tracing) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}
// trace[256 + slot]: blocked cycles per role of CTA 0
enum { TW_APROD_EMPTY = 0, TW_BPROD_EMPTY = 1, TW_MMA_FULLA = 2, TW_MMA_FULLB = 3, TW_MMA_TEMPTY = 4, TW_EPI_TFULL = 5,
       TW_EPI_STAGE = 6, TW_TOTAL = 7 };
#define YX_TRACE_SUM(slot, v)                                                              \
  do {                                                                                     \
    if (tracing && blockIdx.x == 0 && lane == 0) p.trace[256 + (slot)] = (v);              \
  } while (0)

// Tile index = ((img * tiles_h + ty) * tiles_w + tx) * n_tiles_n + nt.  Each role walks tiles blockIdx.x, +gridDim.x, ...;
// the coordinates are kept as mixed-radix digits and advanced by the (host-decomposed) step with carries, because the
// three runtime integer divisions of a plain decode cost ~450 cycles of dependent latency per tile in EVERY role.
struct TileIter {
  int nt, tx, ty, img;
  __device__ __forceinline__ void init(const ConvParams& p, int tile) {
    nt = tile % p.n_tiles_n;
    int r = tile / p.n_tiles_n;
    tx = r % p.tiles_w; r /= p.tiles_w;
    ty = r % p.tiles_h;
    img = r / p.tiles_h;
  }
  __device__ __forceinline__ void next(int n_tiles_n, int tiles_w, int tiles_h, int s_nt, int s_x, int s_y, int s_img) {
    nt += s_nt;
    int c = nt >= n_tiles_n ? 1 : 0;
    nt -= c ? n_tiles_n : 0;
    tx += s_x + c;
    c = tx >= tiles_w ? 1 : 0;
    tx -= c ? tiles_w : 0;
    ty += s_y + c;
    c = ty >= tiles_h ? 1 : 0;
    ty -= c ? tiles_h : 0;
    img += s_img + c;
  }
};
#define YX_TILE_NEXT(it) (it).next(n_tiles_n, tiles_w, tiles_h, p.step_nt, p.step_x, p.step_y, p.step_img)

// ---- image-fed stem (IMG): the space-to-depth of the NCHW input image (Focus / FocusCustom, network_blocks.py:330-361,
// blocks.py:286-304) is done by the conv kernel's own producer warps, straight into the swizzled operand tile, so the
// 16-channel s2d tensor (1.3 GB written and read back per bs64 step, a separate write-bound kernel) never exists.
// Input affine exactly as the stand-alone s2d kernel applies it (yx_aux.cu affine_in: the predict loop's
// img.mul_(s).add_(b) in the image dtype, main.py:164).
template <typename T> __device__ __forceinline__ float img_affine(T v, float scale, float shift, bool on);
template <> __device__ __forceinline__ float img_affine<__half>(__half v, float scale, float shift, bool on) {
  float x = __half2float(v);
  if (on) {
    x = __half2float(__float2half_rn(x * scale));
    x = __half2float(__float2half_rn(x + shift));
  }
  return x;
}
template <> __device__ __forceinline__ float img_affine<uint8_t>(uint8_t v, float scale, float shift, bool on) {
  return img_affine<__half>(__ushort2half_rn(v), scale, shift, on);
}
template <> __device__ __forceinline__ float img_affine<float>(float v, float scale, float shift, bool on) {
  return on ? (v * scale) + shift : v;
}

constexpr int kImgProdThreads = 224;  // warps 0, 2, 3 + one extra warpgroup behind the epilogue warps (the tile builder is
                                      // latency-bound per task: more warps, not fewer instructions, is what speeds it up)

// One tile of the stem's halo operand from the raw image.  Tile = TH x 8 s2d pixels at (y0, x0); the operand is the
// (TH + 2) x 10 halo of s2d pixels, one 128-byte swizzled smem row each (16 channels = 32 bytes used: 12 real + 4 zero;
// the nine filter taps are nine descriptors into it, K = 16 per tap), out-of-image pixels zero (the conv's padding acts on
// the s2d tensor).  The raw patch (3 planes x 2(TH+2) rows x 32 pixels) arrives in a scratch buffer by ONE TMA load (issued
// a tile ahead into the other scratch buffer); one task per s2d pixel then converts its 2x2x3 values and writes 32 bytes.
// Plain C++ shared-memory accesses through generic pointers (not asm volatile): the compiler may overlap them.
template <typename T>
__device__ __forceinline__ void img_build_tile(const ConvParams& p, uint8_t* stageA, const uint8_t* scratch, uint32_t scratch_u32,
                                               const unsigned short* lut, int tid, int x0, int y0, int TH) {
  const int H2 = p.img_h >> 1, W2 = p.img_w >> 1;
  const int prow = 2 * (TH + 2);
  const bool affine = p.img_affine != 0;
  const float scale = p.img_scale, shift = p.img_shift;
  const int order = p.img_order;
  const int tasks = (TH + 2) * kHaloW;
#pragma unroll 2
  for (int t = tid; t < tasks; t += kImgProdThreads) {
    const int pp = t % kHaloW, yy = t / kHaloW;
    const int sx = x0 - 1 + pp, sy = y0 - 1 + yy;
    unsigned short v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0;
    if (sx >= 0 && sx < W2 && sy >= 0 && sy < H2) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        unsigned short h[2][2];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          // the patch arrives as 128-byte rows in the TMA's 128-byte swizzle (16-byte chunk index ^ (row & 7)); the pixel
          // pair of s2d column pp sits at byte 2 * pp * sizeof(T) of patch row R = plane * prow + image row
          // (the row starts kImgLead<T> pixels left of the tile's first needed pixel 2*x0 - 2: see img_producer_loop)
          const uint32_t R = (uint32_t)(c * prow + 2 * yy + dy), byte = (uint32_t)(16 / (int)sizeof(T) - 2 + 2 * pp) * (uint32_t)sizeof(T);
          // (the swizzle XORs the chunk index with bits [7,10) of the shared-memory ADDRESS of the row: the patch buffers are
          // 128- but not 1024-byte aligned, so the row's position inside its 1024-byte window comes from the address)
          const uint8_t* src = scratch + R * 128u + ((((byte >> 4) ^ ((scratch_u32 >> 7) + R)) & 7u) << 4) + (byte & 15u);
          if (sizeof(T) == 1) {
            // uint8 pixels: the affine + fp16 roundings come from a 256-entry table built once per CTA (exactly the
            // values img_affine<uint8_t> gives): two table reads instead of ~16 conversion instructions per pixel pair
            const unsigned short w = *reinterpret_cast<const unsigned short*>(src);
            h[dy][0] = lut[w & 0xffu];
            h[dy][1] = lut[w >> 8];
          } else if (sizeof(T) == 2) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(src);
            h[dy][0] = __half_as_ushort(__float2half_rn(img_affine<__half>(__ushort_as_half((unsigned short)(w & 0xffff)), scale, shift, affine)));
            h[dy][1] = __half_as_ushort(__float2half_rn(img_affine<__half>(__ushort_as_half((unsigned short)(w >> 16)), scale, shift, affine)));
          } else {
            const float2 w = *reinterpret_cast<const float2*>(src);
            h[dy][0] = __half_as_ushort(__float2half_rn(img_affine<float>(w.x, scale, shift, affine)));
            h[dy][1] = __half_as_ushort(__float2half_rn(img_affine<float>(w.y, scale, shift, affine)));
          }
        }
        if (order == 1) {   // pixel_unshuffle: channel = c*4 + dy*2 + dx
          v[c * 4 + 0] = h[0][0]; v[c * 4 + 1] = h[0][1]; v[c * 4 + 2] = h[1][0]; v[c * 4 + 3] = h[1][1];
        } else {            // Focus: [TL, BL, TR, BR] patch-major
          v[0 + c] = h[0][0]; v[3 + c] = h[1][0]; v[6 + c] = h[0][1]; v[9 + c] = h[1][1];
        }
      }
    }
    uint4 lo, hi;
    lo.x = v[0] | ((uint32_t)v[1] << 16); lo.y = v[2] | ((uint32_t)v[3] << 16); lo.z = v[4] | ((uint32_t)v[5] << 16); lo.w = v[6] | ((uint32_t)v[7] << 16);
    hi.x = v[8] | ((uint32_t)v[9] << 16); hi.y = v[10] | ((uint32_t)v[11] << 16); hi.z = 0u; hi.w = 0u;
    const uint32_t r = (uint32_t)t;   // halo row = yy * 10 + pp
    uint8_t* line = stageA + r * 128;
    *reinterpret_cast<uint4*>(line + ((0u ^ (r & 7u)) << 4)) = lo;
    *reinterpret_cast<uint4*>(line + ((1u ^ (r & 7u)) << 4)) = hi;
  }
  fence_proxy_async_smem();   // generic-proxy writes of this thread -> visible to the tensor core's async-proxy reads
}

// the producers' loop over this CTA's tiles for one image dtype (tm_img = &p.tmImg taken in the kernel body, so that it is
// certainly a param-space address)
template <typename T>
__device__ __forceinline__ void img_producer_loop(const ConvParams& p, const CUtensorMap* tm_img, uint8_t* smem_gen, uint32_t smem_gen_u32, uint32_t sA,
                                                  uint32_t sScratch, uint32_t bar_fa, uint32_t bar_ea, uint32_t bar_patch, int tid,
                                                  int tile_first, int n_tiles, int tile_step, uint32_t stages_a) {
  const int n_tiles_n = p.n_tiles_n, tiles_w = p.tiles_w, tiles_h = p.tiles_h, TH = p.TH;
  // the image is described to the TMA as fp16 PAIRS OF BYTES (any pixel type): inner box = 64 such elements = 128 bytes
  // per patch row, SWIZZLE_128B -- the one tensor-map shape class this kernel family is known to run (a plain unswizzled
  // 4-D byte / fp16 map of the same tensor raised "illegal instruction" on this driver, whatever its parameters)
  const uint32_t patch_bytes = (uint32_t)(3 * 2 * (TH + 2)) * 128u;
  const uint32_t buf_bytes = patch_bytes;
  // The innermost TMA coordinate must address a 16-BYTE ALIGNED element (a start 4 bytes left of an aligned pixel is an
  // "illegal instruction"): rows start lead = 16 / sizeof(T) pixels (exactly 16 bytes) left of pixel 2*x0 = 16*tx, and the
  // builder skips lead - 2 pixels.  x coordinate in fp16-sized elements: (16*tx - lead) * sizeof(T) / 2 = 8*tx*sizeof(T) - 8.
  const int xs = (int)sizeof(T);
  // uint8 images: value table behind the two patch buffers (the scratch region is sized for fp32 pixels)
  const uint32_t sLut = sScratch + 2u * buf_bytes;
  if (sizeof(T) == 1) {
    for (int i = tid; i < 256; i += kImgProdThreads) {
      const unsigned short h = __half_as_ushort(__float2half_rn(img_affine<uint8_t>((uint8_t)i, p.img_scale, p.img_shift, p.img_affine != 0)));
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(sLut + 2u * i), "h"(h) : "memory");
    }
  }
  named_bar_sync(5, kImgProdThreads);
  uint32_t s = 0, ph = 0, i = 0;
  TileIter ti;
  ti.init(p, tile_first);
  if (tid == 0 && tile_first < n_tiles) {   // the first tile's patch
    mbar_expect_tx(bar_patch, patch_bytes);
    tma_load_4d(sScratch, tm_img, bar_patch, ti.tx * p.TW * xs - 8, 2 * (ti.ty * TH - 1), 0, ti.img);
  }
  for (int tile = tile_first; tile < n_tiles; tile += tile_step, ++i) {
    const int x0 = ti.tx * p.TW, y0 = ti.ty * TH;
    YX_TILE_NEXT(ti);
    const uint32_t b = i & 1u;
    if (tid == 0 && tile + tile_step < n_tiles) {   // next tile's patch into the other buffer: its last reader (tile i - 1) passed
      mbar_expect_tx(bar_patch + 8 * (b ^ 1u), patch_bytes);                                  // barrier 6 before this point
      tma_load_4d(sScratch + (b ^ 1u) * buf_bytes, tm_img, bar_patch + 8 * (b ^ 1u), ti.tx * p.TW * xs - 8, 2 * (ti.ty * TH - 1), 0, ti.img);
    }
    mbar_wait(bar_ea + 8 * s, ph ^ 1);
    mbar_wait(bar_patch + 8 * b, (i >> 1) & 1u);
    img_build_tile<T>(p, smem_gen + (sA + s * p.a_stage_bytes - smem_gen_u32), smem_gen + (sScratch + b * buf_bytes - smem_gen_u32),
                      sScratch + b * buf_bytes, reinterpret_cast<const unsigned short*>(smem_gen + (sLut - smem_gen_u32)), tid, x0, y0, TH);
    named_bar_sync(6, kImgProdThreads);
    if (tid == 0) mbar_arrive(bar_fa + 8 * s);   // after the producers' barrier: every thread's fenced writes are in
    if (++s == stages_a) { s = 0; ph ^= 1; }
  }
}

// ---- MMA issue helpers.  Everything here runs in the single MMA warp, warp-uniformly; the tensor pipe can only be as
// busy as this warp is fast (ncu: with ~85 SASS instructions per tap the warp, not the tensor core, was the limiter
// of every layer with N <= 192), so the tap sequence is fully unrolled and all loop state lives in locals.
template <bool PAIR>
__device__ __forceinline__ void umma_x(uint32_t d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                       uint32_t accumulate) {
  if (PAIR) umma_f16_ss_lohi_2sm(d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  else umma_f16_ss_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
}
template <bool PAIR>
__device__ __forceinline__ void commit_x(uint32_t bar) {
  if (PAIR) umma_commit_2sm(bar);
  else umma_commit(bar);
}

// RP: the row-packed stem (3 vertical taps over rows of TW = 8 pixels whose 128-byte smem row already holds the three
// horizontal neighbours): the halo box is (TH+2) x 8 rows, tap dy starts dy*8 rows (= one swizzle atom) later, SBO = 1024.
// SP (2:4 sparse variant): the weights are the sparse A operand and the pixels the B operand of tcgen05.mma.sp; ksteps counts
// K = 32 steps (at most two per 64-channel chunk), the weight rows advance 32 bytes and the pixel rows 64 bytes per step,
// e_addr is the TMEM address of the metadata column of (tap 0, this chunk, step 0), consecutive taps are e_step columns apart.
template <int MH, bool RING, bool PAIR, bool RP, bool SP>
__device__ __forceinline__ void halo_chunk_mma(uint32_t d0, uint32_t d1, uint32_t a0, uint32_t a_hi, uint32_t& b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t& accum, int ksteps, uint32_t bar_fb, uint32_t bar_eb,
                                               uint32_t& sb, uint32_t& phb, uint32_t b_slots, uint32_t b_lo0, uint32_t b_step,
                                               bool wait_b, bool tracing, long long& w_acc, bool skip_mma, uint32_t e_addr,
                                               uint32_t e_step) {
  constexpr int TAPS = RP ? 3 : 9;
  constexpr int HW = RP ? 8 : kHaloW;  // halo row width in pixels
  if (SP) {
#pragma unroll
    for (int tap = 0; tap < TAPS; ++tap) {
      const uint32_t a_tap = a0 + ((tap / 3) * HW + (tap % 3)) * 8;
      if (RING || wait_b) {
        mbar_wait_acc(bar_fb + 8 * sb, phb, tracing, w_acc);
        tc_fence_after();
      }
      if (elect_one()) {
        if (!skip_mma) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            if (ks < ksteps)
            {
              const uint32_t col = e_addr + tap * e_step + ks;   // even-aligned column address + 1-bit selector
              umma_f16_sp_ss_lohi(d0, b_lo + 2 * ks, b_hi, a_tap + 4 * ks, a_hi, col & ~1u, idesc | (col & 1u), ks == 0 ? accum : 1u);
            }
        }
        if (RING) commit_x<false>(bar_eb + 8 * sb);
      }
      accum = 1;
      b_lo += b_step;
      if (RING) {
        if (++sb == b_slots) { sb = 0; phb ^= 1; b_lo = b_lo0; }
      } else {
        ++sb;
      }
    }
    return;
  }
  if (!RING && !wait_b) {
    // resident weights already in shared memory: nothing to wait for inside the chunk, so the whole 9-tap sequence
    // is ONE elected straight-line block (no per-tap elect / branch / reconvergence)
    if (elect_one() && !skip_mma) {
#pragma unroll
      for (int tap = 0; tap < TAPS; ++tap) {
        const uint32_t a_tap = a0 + (RP ? tap * HW : (tap / 3) * HW + (tap % 3)) * 8;
        const uint32_t b_tap = b_lo + tap * b_step;
        if (ksteps == 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_tap + 2 * ks, b_hi, idesc, (tap | ks) == 0 ? accum : 1u);
            if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * HW * 8 + 2 * ks, a_hi, b_tap + 2 * ks, b_hi, idesc, (tap | ks) == 0 ? accum : 1u);
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            if (ks < ksteps) {
              umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_tap + 2 * ks, b_hi, idesc, (tap | ks) == 0 ? accum : 1u);
              if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * HW * 8 + 2 * ks, a_hi, b_tap + 2 * ks, b_hi, idesc, (tap | ks) == 0 ? accum : 1u);
            }
        }
      }
    }
    accum = 1;
    b_lo += TAPS * b_step;
    return;
  }
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    // tap (dy,dx) of half h starts (16h + dy) halo rows down and dx pixels right: rows are 128 B = 8 descriptor units
    const uint32_t a_tap = a0 + (RP ? tap * HW : (tap / 3) * HW + (tap % 3)) * 8;
    if (wait_b) {
      mbar_wait_acc(bar_fb + 8 * sb, phb, tracing, w_acc);
      tc_fence_after();
    }
    if (elect_one()) {
      if (skip_mma) {
      } else if (ksteps == 4) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
          if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * HW * 8 + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
        }
      } else {
        for (int ks = 0; ks < ksteps; ++ks) {
          umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
          if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * HW * 8 + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
        }
      }
      if (RING) commit_x<PAIR>(bar_eb + 8 * sb);
    }
    accum = 1;
    b_lo += b_step;
    if (RING) {
      if (++sb == b_slots) { sb = 0; phb ^= 1; b_lo = b_lo0; }
    } else {
      ++sb;  // resident weights, first tile: every tap waits on ITS OWN slot's barrier (the loads arrive one by one)
    }
  }
}

// Streamed weights with ONE RING STAGE PER FILTER ROW (p.b_taps == 3): per stage one wait, one elected straight-line block of
// 3 taps x ksteps MMAs, one commit.  b_tap = descriptor units between the taps of a stage.  DIAG: tracing / role-ablation
// build of the same loop (blocked-cycle accounting, MMAs optionally skipped); the product path is DIAG = false.
template <int MH, bool PAIR, bool DIAG>
__device__ __forceinline__ void halo_chunk_mma_rows(uint32_t d0, uint32_t d1, uint32_t a0, uint32_t a_hi, uint32_t& b_lo, uint32_t b_hi,
                                                    uint32_t idesc, uint32_t& accum, int ksteps, uint32_t bar_fb, uint32_t bar_eb,
                                                    uint32_t& sb, uint32_t& phb, uint32_t b_slots, uint32_t b_lo0, uint32_t b_step,
                                                    uint32_t b_tap, bool tracing, long long& w_acc, bool skip_mma) {
#pragma unroll
  for (int row = 0; row < 3; ++row) {
    if (DIAG) mbar_wait_acc(bar_fb + 8 * sb, phb, tracing, w_acc);
    else mbar_wait(bar_fb + 8 * sb, phb);
    tc_fence_after();
    if (elect_one()) {
      if (!DIAG || !skip_mma) {
        if (ksteps == 4) {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t a_tap = a0 + (row * kHaloW + dx) * 8, b_t = b_lo + dx * b_tap;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_t + 2 * ks, b_hi, idesc, (dx | ks) == 0 ? accum : 1u);
              if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * kHaloW * 8 + 2 * ks, a_hi, b_t + 2 * ks, b_hi, idesc, (dx | ks) == 0 ? accum : 1u);
            }
          }
        } else {
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t a_tap = a0 + (row * kHaloW + dx) * 8, b_t = b_lo + dx * b_tap;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks)
              if (ks < ksteps) {
                umma_x<PAIR>(d0, a_tap + 2 * ks, a_hi, b_t + 2 * ks, b_hi, idesc, (dx | ks) == 0 ? accum : 1u);
                if (MH == 2) umma_x<PAIR>(d1, a_tap + 16 * kHaloW * 8 + 2 * ks, a_hi, b_t + 2 * ks, b_hi, idesc, (dx | ks) == 0 ? accum : 1u);
              }
          }
        }
      }
      commit_x<PAIR>(bar_eb + 8 * sb);
    }
    accum = 1;
    b_lo += b_step;
    if (++sb == b_slots) { sb = 0; phb ^= 1; b_lo = b_lo0; }
  }
}

// MODE: 0 = generic (one A box per tap), 1 = halo with one 128-pixel half per CTA, 2 = halo with two stacked halves,
//       3 = generic with a 256-pixel tile (two 128-row halves per A box): halves every per-tile fixed cost of the
//           HBM-bound small-channel layers (N <= 128)
//       4/5 = row-packed stem as a VERTICAL halo (one / two halves): one (TH+2)-row box per tile instead of three
//           shifted boxes -> 2.7x fewer bytes through L2->smem for the layer that was bound by exactly that
// PAIR: two CTAs of a cluster share every MMA (cta_group::2, M = 256): half of the weight tile per CTA
// RES:  0 = no residual; 1 = residual tile TMA-loaded into the staging buffer and added in registers;
//       2 = the residual IS the destination (Bottleneck y = x + f(x) computed in place): the tile is stored with a TMA
//           reduce-add, so the residual never passes through shared memory and the epilogue never waits for it
// SP:   2:4 sparse tensor-core variant (MODE 0 / 1, no pair): weights = sparse A operand of tcgen05.mma.sp (M = 128 output
//       channels per MMA), pixels = B operand (N = 128), accumulator transposed (channels on lanes, pixels on columns)
// IMG:  image-fed stem (MODE 1 / 2): the halo operand tile is built from the raw NCHW image by warps 0, 2, 3
template <int ACT, int RES, int MODE, bool PAIR, bool SP = false, bool IMG = false>
__global__ void __launch_bounds__(IMG ? 512 : 384, 1) conv_gemm_kernel(const __grid_constant__ ConvParams p) {
  static_assert(!SP || (!PAIR && (MODE == 0 || MODE == 1)), "sparse variant: generic or single-half halo tiles, no CTA pair");
  static_assert(!IMG || (!PAIR && !SP && RES == 0 && (MODE == 1 || MODE == 2)), "image-fed variant: single-CTA halo tiles, no residual");
  constexpr bool HAS_RES = RES == 1;
  constexpr bool HALO = MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5;
  constexpr bool RP = MODE >= 4;
  constexpr int MH = (MODE == 2 || MODE == 3 || MODE == 5) ? 2 : 1;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // position in the CTA pair; rank 0 = leader (issues the MMAs)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // tells ptxas the warp index is warp-uniform
  const int lane = threadIdx.x & 31;

  const int groups64 = (p.BN + 63) >> 6;  // 64-channel output groups per N tile
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + p.stages_a * p.a_stage_bytes;
  const uint32_t sStage0 = sB + p.b_slots * p.b_stage_bytes;
  const uint32_t stage_buf_bytes = groups64 * kTileBytes;
  const uint32_t sBias = sStage0 + p.stage_bufs * stage_buf_bytes;
  const uint32_t sBar = sBias + p.bias_bytes;
  const uint32_t sScratch = sBar + kBarBytes;   // IMG: raw image patch of the tile being built
  // barriers (8 bytes each): fullA[8] emptyA[8] fullB[32] emptyB[32] tfull[2] tempty[2] res[2], then the TMEM slot.
  // shared_ring (generic, streamed weights): A and B of a k-iteration share fullA/emptyA (two producer arrivals), so
  // the MMA warp waits and commits once per k-iteration.
  const uint32_t bar_fa = sBar, bar_ea = sBar + 64;
  const uint32_t bar_fb = p.shared_ring ? bar_fa : sBar + 128, bar_eb = p.shared_ring ? bar_ea : sBar + 384;
  const uint32_t bar_tfull = sBar + 640, bar_tempty = sBar + 656, bar_res = sBar + 672, tmem_slot = sBar + 688;

  if (warp == 0 && lane == 0) {
    if (IMG) tma_prefetch_desc(&p.tmImg);
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages_a; ++s) { mbar_init(bar_fa + 8 * s, p.shared_ring ? 2 : 1); mbar_init(bar_ea + 8 * s, 1); }
    if (!p.shared_ring)
      for (int s = 0; s < p.b_slots; ++s) { mbar_init(bar_fb + 8 * s, 1); mbar_init(bar_eb + 8 * s, 1); }
    if (IMG) { mbar_init(sBar + 768, 1); mbar_init(sBar + 776, 1); }   // raw-patch buffers (TMA completion)
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      // one arrive per epilogue warp that drains this accumulator (of both CTAs); alternating groups: one group each
      mbar_init(bar_tempty + 8 * a, (PAIR ? 8 : 4) * (p.epi_alt ? 1 : p.epi_groups));
      mbar_init(bar_res + 8 * a, 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2sm(tmem_slot, p.tmem_cols); else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  for (int i = threadIdx.x; i < p.cout16; i += blockDim.x) {
    const float b = __ldg(p.bias + i);
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(sBias + i * 4), "f"(b) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers must be initialised before anything is signalled across the pair
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (SP) {
    // sparsity metadata of every output-channel tile -> tensor memory, once per CTA (weights are constants: this also
    // overlaps the previous layer's tail).  Warp w of the first epilogue group owns TMEM lanes 32 * (w % 4) .. + 31.
    if (warp >= 4 && warp < 8) {
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (uint32_t)p.sp_meta_col0;
      const uint32_t* src = p.sp_meta + (warp & 3) * 32 + lane;
      const int ncol = p.n_tiles_n * p.sp_cols_per_tile;
      for (int c = 0; c < ncol; ++c) tmem_st_32x32b_x1(lane_addr + c, __ldg(src + c * 128));
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // PAIR: the tile space is walked in units of CTA pairs; CTA `rank` owns the x-tile 2*tx + rank of every pair tile
  const int n_tiles = p.n_tiles_m * p.n_tiles_n;
  const int tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int tile_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int n_tiles_n = p.n_tiles_n, tiles_w = p.tiles_w, tiles_h = p.tiles_h;
  const int taps = p.ky * p.kx;
  const int k_chunks = p.k_chunks;
  const uint32_t stages_a = p.stages_a, b_slots = p.b_slots;
  const bool resident = p.b_resident != 0;

  const bool tracing = p.trace != nullptr;
  long long w_acc0 = 0, w_acc1 = 0, w_acc2 = 0;
  const long long t_begin = tracing ? clock64() : 0;
  const bool is_a_prod = !IMG && ((warp == 0) || (warp == 3 && p.w3_role == 1));
  const bool is_b_prod = (warp == 2) || (warp == 3 && p.w3_role == 2);
  const bool is_a_prod_late = !IMG && warp == 2 && p.w2_role == 1;   // after its (resident) weight loads
  const int img_extra0 = 4 + 4 * p.epi_groups;   // IMG: first warp of the extra builder warpgroup
  const bool is_img_prod = IMG && (warp == 0 || warp == 2 || warp == 3 || warp >= img_extra0);

  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, bias copy: none of it touches an
  // activation) overlapped the previous layer's tail.  The weight producers go on without waiting -- weights are constants,
  // so resident weights are already streaming in while the previous layer finishes; every other role reads or writes
  // activations and waits for the previous grid to complete first.
  griddep_launch_dependents();
  if (!is_b_prod) griddep_wait();

  // Producer and MMA warps run their loops WARP-UNIFORMLY (all 32 lanes, uniform values) and elect one lane only
  // around the issue itself: under `if (lane == 0)` ptxas cannot prove uniformity and wraps every UTMALDG / UTCHMMA
  // in an ELECT + 8x R2UR.BROADCAST loop (~200 cycles per TMA instruction, profiles/r01_tma_issue_probe.txt).
  if (is_b_prod) {
    // ================================ B (weight) producer(s) ================================
    const uint32_t nprod = p.w3_role == 2 ? 2u : 1u, mine = warp == 2 ? 0u : 1u;
    const uint32_t b_stage_bytes = p.b_stage_bytes;
    const int BN = p.BN, cout16 = p.cout16;
    uint32_t s = 0, ph = 0, turn = 0;
    int nt = tile_first % n_tiles_n;
    for (int tile = tile_first; tile < n_tiles; tile += tile_step) {
      int n0 = nt * BN;
      if (PAIR) n0 += (int)rank * (min(BN, cout16 - n0) >> 1);  // this CTA's half of the N tile
      nt += p.step_nt;
      if (nt >= n_tiles_n) nt -= n_tiles_n;
      // ring order must match the MMA issuer: halo = (chunk, tap), generic = (tap, chunk)
      const int b_taps = p.b_taps;   // halo ring: taps per stage (3 = one filter row per load)
      const int outer = HALO ? k_chunks : taps, inner = HALO ? taps / b_taps : k_chunks;
      for (int o = 0; o < outer; ++o)
        for (int i = 0; i < inner; ++i) {
          const int kc = HALO ? o : i, tap = HALO ? i * b_taps : o;
          if (turn == mine) {
            if (!resident) mbar_wait_acc(bar_eb + 8 * s, ph ^ 1, tracing, w_acc0);
            if (elect_one()) {
              if (PAIR) {
                if (rank == 0) mbar_expect_tx(bar_fb + 8 * s, 2 * b_stage_bytes);
                tma_load_3d_2sm(sB + s * b_stage_bytes, &p.tmW, bar_fb + 8 * s, kc * 64, n0, tap);
              } else {
                mbar_expect_tx(bar_fb + 8 * s, b_stage_bytes);
                if (SP) tma_load_3d(sB + s * b_stage_bytes, &p.tmW, bar_fb + 8 * s, kc * 32, tap, n0);  // SP: compressed rows, (cin/2, tap, cout)
                else tma_load_3d(sB + s * b_stage_bytes, &p.tmW, bar_fb + 8 * s, kc * 64, n0, tap);
              }
            }
          }
          if (++turn == nprod) turn = 0;
          if (++s == b_slots) { s = 0; ph ^= 1; }
        }
      if (resident) break;  // one N tile per layer: the weights stay in smem for every later tile
    }
    if (warp == 2) YX_TRACE_SUM(TW_BPROD_EMPTY, w_acc0);
    w_acc0 = 0;
    // resident weights: this warp has nothing left to do for the rest of the kernel -> it becomes one more producer of
    // the ACTIVATION operand (w2_role).  One issuing warp sustains roughly one TMA load per 400-1000 cycles (strided
    // stride-2 boxes are the slow end), so the layers that load a box per filter tap are bound by how many warps issue.
    if (is_a_prod_late || (IMG && warp == 2)) griddep_wait();
  }
  if (is_img_prod) {
    // ================================ image producers (IMG): warps 0, 2, 3 build every tile together ================
    const int tid = (warp >= img_extra0 ? 3 + (warp - img_extra0) : (warp == 0 ? 0 : warp - 1)) * 32 + lane;
    if (p.img_dtype == YX_U8) img_producer_loop<uint8_t>(p, &p.tmImg, smem_raw, smem_u32(smem_raw), sA, sScratch, bar_fa, bar_ea, sBar + 768, tid, tile_first, n_tiles, tile_step, stages_a);
    else if (p.img_dtype == YX_F16) img_producer_loop<__half>(p, &p.tmImg, smem_raw, smem_u32(smem_raw), sA, sScratch, bar_fa, bar_ea, sBar + 768, tid, tile_first, n_tiles, tile_step, stages_a);
    else img_producer_loop<float>(p, &p.tmImg, smem_raw, smem_u32(smem_raw), sA, sScratch, bar_fa, bar_ea, sBar + 768, tid, tile_first, n_tiles, tile_step, stages_a);
  } else if (is_a_prod || is_a_prod_late) {
    // ================================ A producer(s) ================================
    const uint32_t nprod = 1u + (p.w3_role == 1 ? 1u : 0u) + (p.w2_role == 1 ? 1u : 0u);
    const uint32_t mine = warp == 0 ? 0u : (warp == 3 ? 1u : nprod - 1u);
    const uint32_t a_stage_bytes = p.a_stage_bytes, a_box_bytes = p.a_box_bytes;
    uint32_t s = 0, ph = 0, turn = 0;  // ring slot / phase / whose turn, all incremental
    const int TH = p.TH, TW = p.TW;
    TileIter ti;
    ti.init(p, tile_first);
    for (int tile = tile_first; tile < n_tiles; tile += tile_step) {
      const int x0 = (PAIR ? 2 * ti.tx + (int)rank : ti.tx) * TW, y0 = ti.ty * TH, img = ti.img;
      YX_TILE_NEXT(ti);
      if (HALO) {
        for (int kc = 0; kc < k_chunks; ++kc) {
          if (turn == mine) {
            mbar_wait_acc(bar_ea + 8 * s, ph ^ 1, tracing, w_acc0);
            if (elect_one()) {
              if (p.diag & 2) {
                mbar_arrive(bar_fa + 8 * s);
              } else if (PAIR) {  // the leader's barrier collects the bytes of both CTAs' tiles
                if (rank == 0) mbar_expect_tx(bar_fa + 8 * s, 2 * a_box_bytes);
                tma_load_4d_2sm(sA + s * a_stage_bytes, &p.tmA[0], bar_fa + 8 * s, kc * 64, x0 - (RP ? 0 : 1), y0 - 1, img);
              } else {
                mbar_expect_tx(bar_fa + 8 * s, a_box_bytes);
                tma_load_4d(sA + s * a_stage_bytes, &p.tmA[0], bar_fa + 8 * s, kc * 64, x0 - (RP ? 0 : 1), y0 - 1, img);
              }
            }
          }
          if (++turn == nprod) turn = 0;
          if (++s == stages_a) { s = 0; ph ^= 1; }
        }
      } else {
        const int stride = p.stride, kx = p.kx, pad_x = p.pad_x, pad_y = p.pad_y, up_chunks = p.up_chunks, up_h = p.up_h;
        int dy = 0, dx = 0;
        for (int tap = 0; tap < taps; ++tap) {
          int mi = 0, cx, cy;
          if (stride == 1) {
            cx = x0 + dx - pad_x;
            cy = y0 + dy - pad_y;
          } else {
            // input row 2*y + dy - pad.  For k=3,pad=1: dy=0 -> odd row of cell y-1; dy=1 -> even row
            // of cell y; dy=2 -> odd row of cell y.  k=1 (pad 0): even row/col of cell y.
            const int oy = dy - pad_y, ox = dx - pad_y;
            const int py = oy & 1, px = ox & 1;
            mi = py * 2 + px;
            cy = y0 + ((oy - py) >> 1);
            cx = x0 + ((ox - px) >> 1);
          }
          for (int kc = 0; kc < k_chunks; ++kc) {
            if (turn == mine) {
              mbar_wait_acc(bar_ea + 8 * s, ph ^ 1, tracing, w_acc0);
              if (elect_one()) {
                if (p.diag & 2) {
                  mbar_arrive(bar_fa + 8 * s);
                } else if (kc < up_chunks) {
                  // fused nearest x2 upsample + concat: these K chunks come from the low-resolution tensor, each
                  // pixel repeated 2x2 by the tensor map's stride-0 dimensions
                  if (PAIR) {
                    if (rank == 0) mbar_expect_tx(bar_fa + 8 * s, 2 * a_box_bytes);
                    tma_load_5d_2sm(sA + s * a_stage_bytes, &p.tmUp, bar_fa + 8 * s, kc * 64, 0, cx >> 1, 0, img * up_h + (cy >> 1));
                  } else {
                    mbar_expect_tx(bar_fa + 8 * s, a_box_bytes);
                    tma_load_5d(sA + s * a_stage_bytes, &p.tmUp, bar_fa + 8 * s, kc * 64, 0, cx >> 1, 0, img * up_h + (cy >> 1));
                  }
                } else if (PAIR) {
                  if (rank == 0) mbar_expect_tx(bar_fa + 8 * s, 2 * a_box_bytes);
                  tma_load_4d_2sm(sA + s * a_stage_bytes, &p.tmA[mi], bar_fa + 8 * s, (kc - up_chunks) * 64, cx, cy, img);
                } else {
                  mbar_expect_tx(bar_fa + 8 * s, a_box_bytes);
                  tma_load_4d(sA + s * a_stage_bytes, &p.tmA[mi], bar_fa + 8 * s, (kc - up_chunks) * 64, cx, cy, img);
                }
              }
            }
            if (++turn == nprod) turn = 0;
            if (++s == stages_a) { s = 0; ph ^= 1; }
          }
          if (++dx == kx) { dx = 0; ++dy; }
        }
      }
      if (warp == 0 && lane == 0) YX_TRACE(0, (tile - tile_first) / tile_step);
    }
    if (warp == 0) YX_TRACE_SUM(TW_APROD_EMPTY, w_acc0);
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader CTA only in PAIR mode) ================================
    uint32_t t = 0, sa = 0, pha = 0, sb = 0, phb = 0;
    const uint32_t a_hi = sdesc_hi(HALO && !RP ? kHaloW * 128 : 1024), b_hi = SP ? sdesc_hi_sw64(512) : sdesc_hi(1024);
    const uint32_t a_lo0 = sdesc_lo(sA), a_step = p.a_stage_bytes >> 4;
    const uint32_t b_lo0 = sdesc_lo(sB), b_step = p.b_stage_bytes >> 4;
    uint32_t a_lo = a_lo0, b_lo = b_lo0;
    const int ks_last = SP ? (p.cin - (k_chunks - 1) * 64) >> 5 : (p.cin - (k_chunks - 1) * 64) >> 4;  // SP: K = 32 steps
    const int BN = p.BN, cout16 = p.cout16;
    const uint32_t acc_stride = p.acc_stride;
    const bool shared_ring = p.shared_ring != 0;
    const bool diag_path = tracing || p.diag != 0;
    const bool b_rows = p.b_taps == 3;                       // halo ring: one stage per filter row
    const uint32_t b_tap_units = (p.b_stage_bytes / 3) >> 4;   // descriptor units between the taps of such a stage
    const bool fast_ok = !tracing && p.diag == 0;   // (YX_CONV_DIAG=16: the general loop, results unchanged)
    bool b_ready = false;  // resident weights: wait for them during the first tile only
    int nt = tile_first % n_tiles_n;
    for (int tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
      const int n0 = nt * BN;
      // SP: first metadata column of this output-channel tile (one column per (tap, K = 32 step); tmem_base has column 0)
      const uint32_t e_tile = SP ? tmem_base + (uint32_t)(p.sp_meta_col0 + nt * p.sp_cols_per_tile) : 0u;
      nt += p.step_nt;
      if (nt >= n_tiles_n) nt -= n_tiles_n;
      const int bn_cur = min(BN, cout16 - n0);
      const uint32_t idesc = SP ? make_idesc_f16_sp(128, 0) : (PAIR ? make_idesc_f16_m256(bn_cur) : make_idesc_f16(bn_cur));
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      mbar_wait_acc(bar_tempty + 8 * acc, acc_ph ^ 1, tracing, w_acc2);
      tc_fence_after();
      if (lane == 0) YX_TRACE(1, t);
      const uint32_t d0 = tmem_base + (acc * MH) * acc_stride, d1 = d0 + acc_stride;
      uint32_t accum = 0;
      if (resident) { sb = 0; b_lo = b_lo0; }
      if (HALO) {
        for (int kc = 0; kc < k_chunks; ++kc) {
          mbar_wait_acc(bar_fa + 8 * sa, pha, tracing, w_acc0);
          tc_fence_after();
          const int ksteps = (kc == k_chunks - 1) ? ks_last : (SP ? 2 : 4);
          const uint32_t e_chunk = e_tile + 2u * (uint32_t)kc, e_step = (uint32_t)p.cin >> 5;
          if (!SP && !RP && !IMG && b_rows) {   // (the image-fed stem always keeps its weights resident)
            if (diag_path)
              halo_chunk_mma_rows<MH, PAIR, true>(d0, d1, a_lo, a_hi, b_lo, b_hi, idesc, accum, ksteps, bar_fb, bar_eb, sb, phb, b_slots,
                                                  b_lo0, b_step, b_tap_units, tracing, w_acc1, (p.diag & 4) != 0);
            else
              halo_chunk_mma_rows<MH, PAIR, false>(d0, d1, a_lo, a_hi, b_lo, b_hi, idesc, accum, ksteps, bar_fb, bar_eb, sb, phb, b_slots,
                                                   b_lo0, b_step, b_tap_units, false, w_acc1, false);
          } else if (resident)
            halo_chunk_mma<MH, false, PAIR, RP, SP>(d0, d1, a_lo, a_hi, b_lo, b_hi, idesc, accum, ksteps, bar_fb, bar_eb, sb, phb, b_slots,
                                                    b_lo0, b_step, !b_ready, tracing, w_acc1, (p.diag & 4) != 0, e_chunk, e_step);
          else
            halo_chunk_mma<MH, true, PAIR, RP, SP>(d0, d1, a_lo, a_hi, b_lo, b_hi, idesc, accum, ksteps, bar_fb, bar_eb, sb, phb, b_slots,
                                                   b_lo0, b_step, true, tracing, w_acc1, (p.diag & 4) != 0, e_chunk, e_step);
          if (elect_one()) commit_x<PAIR>(bar_ea + 8 * sa);
          a_lo += a_step;
          if (++sa == stages_a) { sa = 0; pha ^= 1; a_lo = a_lo0; }
        }
      } else {
        const int k_iters = taps * k_chunks;
        const uint32_t half_units = (uint32_t)(p.TH / 2 * p.TW) * 8;  // MODE 3: second half starts (TH/2)*TW rows of 128 B later
        int kc = 0;
        uint32_t sp_col = 0;   // SP: metadata column of the current (tap, chunk), relative to the tile's first column
        // The common case (whole 64-channel chunks, weights streamed or already resident, no diagnostics) gets a loop with
        // nothing in it but wait -> 4 MMAs -> commit: every instruction here is serial latency of the tensor pipe's only
        // feeder (conv_trace: a k-iteration of the general loop below costs ~420 cycles of issue, more than the 384 cycles
        // its four N = 192 MMAs take, so every streamed layer with N <= 192 ran at the issue rate, not the MMA rate).
        if (!SP && fast_ok && (!resident || b_ready)) {
          // one k-iteration: FULL = four K = 16 steps, otherwise the ks_last (1..3) steps of a partial last chunk
#define YX_FAST_ITER(FULL)                                                                                              \
          {                                                                                                             \
            mbar_wait(bar_fa + 8 * sa, pha);                                                                            \
            tc_fence_after();                                                                                           \
            if (elect_one()) {                                                                                          \
              _Pragma("unroll") for (int ks = 0; ks < (FULL ? 4 : 3); ++ks)                                             \
                if (FULL || ks < ks_last) {                                                                             \
                  umma_x<PAIR>(d0, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);              \
                  if (MH == 2) umma_x<PAIR>(d1, a_lo + half_units + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u); \
                }                                                                                                       \
              commit_x<PAIR>(bar_ea + 8 * sa);                                                                          \
            }                                                                                                           \
            accum = 1;                                                                                                  \
            a_lo += a_step;                                                                                             \
            b_lo += b_step;                                                                                             \
            if (++sa == stages_a) { sa = 0; pha ^= 1; a_lo = a_lo0; if (shared_ring) b_lo = b_lo0; }                    \
          }
          if (ks_last == 4) {
            for (int i = 0; i < k_iters; ++i) YX_FAST_ITER(true)
          } else {
            for (int tap = 0; tap < taps; ++tap) {
              for (int c = 1; c < k_chunks; ++c) YX_FAST_ITER(true)
              YX_FAST_ITER(false)
            }
          }
#undef YX_FAST_ITER
        } else
        for (int i = 0; i < k_iters; ++i) {
          mbar_wait_acc(bar_fa + 8 * sa, pha, tracing, w_acc0);   // shared ring: covers the weights of this k-iteration too
          if (resident && !b_ready) mbar_wait_acc(bar_fb + 8 * sb, phb, tracing, w_acc1);
          tc_fence_after();
          const bool last = ++kc == k_chunks;
          if (last) kc = 0;
          if (elect_one()) {
            if (p.diag & 4) {
            } else if (SP) {  // metadata column of (tap, chunk, step) = tap * (cin / 32) + 2 * chunk + step
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                if (ks == 0 || !last || ks_last == 2) {
                  const uint32_t col = e_tile + sp_col + ks;
                  umma_f16_sp_ss_lohi(d0, b_lo + 2 * ks, b_hi, a_lo + 4 * ks, a_hi, col & ~1u, idesc | (col & 1u), ks == 0 ? accum : 1u);
                }
            } else if (!last || ks_last == 4) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                umma_x<PAIR>(d0, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
                if (MH == 2) umma_x<PAIR>(d1, a_lo + half_units + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < 3; ++ks)
                if (ks < ks_last) {
                  umma_x<PAIR>(d0, a_lo + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
                  if (MH == 2) umma_x<PAIR>(d1, a_lo + half_units + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, ks == 0 ? accum : 1u);
                }
            }
            commit_x<PAIR>(bar_ea + 8 * sa);  // frees the stage (A, and B when the ring is shared) when these MMAs retire
          }
          accum = 1;
          if (SP) sp_col += last ? (uint32_t)ks_last : 2u;
          a_lo += a_step;
          b_lo += b_step;
          if (++sa == stages_a) { sa = 0; pha ^= 1; a_lo = a_lo0; if (shared_ring) b_lo = b_lo0; }
          if (resident) ++sb;
        }
      }
      if (resident) b_ready = true;
      if (elect_one()) commit_x<PAIR>(bar_tfull + 8 * acc);  // accumulator complete -> epilogue (of both CTAs)
      if (lane == 0) YX_TRACE(2, t);
    }
    YX_TRACE_SUM(TW_MMA_FULLA, w_acc0);
    YX_TRACE_SUM(TW_MMA_FULLB, w_acc1);
    YX_TRACE_SUM(TW_MMA_TEMPTY, w_acc2);
    YX_TRACE_SUM(TW_TOTAL, clock64() - t_begin);
  } else if (warp >= 4 && (!IMG || warp < img_extra0)) {
    // ================================ epilogue (warps 4..7 [, 8..11]) ================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int group = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const bool row_valid = HALO ? true : (row < p.TH / MH * p.TW);
    // Two ways to use two epilogue warpgroups: split the COLUMNS of every tile between them (both walk every tile in
    // lockstep), or ALTERNATE TILES (epi_alt): group g owns accumulator g, staging buffer g, its own named barriers and its
    // own store-issuing warp, and drains tiles t = g (mod 2) on its own.  The per-tile chain (wait accumulator -> tcgen05.ld
    // -> convert -> fence -> barrier -> TMA store) is latency-, not throughput-bound on the small-channel layers; with
    // alternating groups two such chains overlap.
    const bool alt = p.epi_alt != 0;
    const bool lead_warp = alt ? (warp & 3) == 0 : warp == 4;  // issues the residual loads and the stores (one elected lane)
    const int epi_groups = p.epi_groups, stage_bufs = p.stage_bufs;
    const uint32_t n_epi = alt ? 128u : 128u * epi_groups;
    const uint32_t bar_id1 = alt ? 1u + 2u * group : 1u, bar_id2 = alt ? 2u + 2u * group : 2u;
    const int cgroup = alt ? 0 : group, cgroups = alt ? 1 : epi_groups;   // column split inside a tile
    const int store_th = HALO ? 16 : p.TH / MH;
    const int BN = p.BN, cout16 = p.cout16;
    const uint32_t acc_stride = p.acc_stride;
    uint32_t t = 0, u = 0;  // tile counter, staging-unit counter (MH units per tile)
    const int TH = p.TH, TW = p.TW;
    TileIter ti;
    ti.init(p, tile_first);
    for (int tile = tile_first; tile < n_tiles; tile += tile_step, ++t) {
      struct { int x0, y0, img, n0; } tc = {(PAIR ? 2 * ti.tx + (int)rank : ti.tx) * TW, ti.ty * TH, ti.img, ti.nt * BN};
      YX_TILE_NEXT(ti);
      const int bn_cur = min(BN, cout16 - tc.n0);
      const int groups_cur = (bn_cur + 63) >> 6;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      if (alt && (int)acc != group) continue;   // the other group's tile
#pragma unroll
      for (int h = 0; h < MH; ++h, ++u) {
        const int yh = tc.y0 + store_th * h;
        const uint32_t buf = alt ? (uint32_t)group : (stage_bufs == 2 ? (u & 1) : 0);
        const uint32_t sStage = sStage0 + buf * stage_buf_bytes;
        // staging buffer `buf` is free once the store issued stage_bufs units ago has read it
        // (elect.sync picks the same lane for the same mask every time, so the bulk-group state stays with one thread)
        const long long ts0 = tracing ? clock64() : 0;
        if (lead_warp) {
          if (elect_one()) {
            if (stage_bufs == 2 && !alt) tma_store_wait_read1(); else tma_store_wait_read0();
          }
        }
        named_bar_sync(bar_id1, n_epi);
        if (tracing) w_acc1 += clock64() - ts0;
        if (HAS_RES && lead_warp) {
          if (elect_one()) {
            mbar_expect_tx(bar_res + 8 * buf, groups_cur * p.out_box_bytes);
            for (int g = 0; g < groups_cur; ++g)
              tma_load_4d(sStage + g * kTileBytes, &p.tmRes, bar_res + 8 * buf, tc.n0 + g * 64, tc.x0, yh, tc.img);
          }
        }
        if (h == 0) {
          mbar_wait_acc(bar_tfull + 8 * acc, acc_ph, tracing, w_acc0);
          tc_fence_after();
          if (lead_warp && lane == 0) YX_TRACE(3, t);
        }
        if (HAS_RES) mbar_wait(bar_res + 8 * buf, ((stage_bufs == 2 && !alt) ? (u >> 1) : u) & 1);
        const uint32_t taddr = tmem_base + (acc * MH + h) * acc_stride + (static_cast<uint32_t>(q * 32) << 16);
        if (SP) {
          if (!(p.diag & 1)) {
            float bias = 0.0f;
            if (row < bn_cur) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bias) : "r"(sBias + (tc.n0 + row) * 4));
            epilogue_convert_sp<ACT, HAS_RES>(taddr, HALO ? 128 : p.TH * p.TW, row, row < bn_cur, bias, sStage, cgroup, cgroups);
          }
        } else if (!(p.diag & 1)) {
          epilogue_convert<ACT, HAS_RES>(taddr, bn_cur, row, row_valid, sStage, sBias + tc.n0 * 4, cgroup, cgroups);
        }
        if (h == MH - 1) {  // accumulator drained -> MMA warp may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) {
              if (p.diag & 8) mbar_arrive_cluster(mapa_shared(bar_tempty + 8 * acc, 0));
              else mbar_arrive_cluster_relaxed(mapa_shared(bar_tempty + 8 * acc, 0));
            } else {
              mbar_arrive(bar_tempty + 8 * acc);
            }
          }
        }
        // publish the staged tile to the async proxy and store it
        fence_proxy_async_smem();
        if (lead_warp && lane == 0 && h == MH - 1) YX_TRACE(4, t);
        named_bar_sync(bar_id2, n_epi);
        if (lead_warp && !(p.diag & 1)) {
          if (elect_one()) {
            for (int g = 0; g < groups_cur; ++g) {
              if (RES == 2) tma_reduce_add_4d(&p.tmOut, sStage + g * kTileBytes, tc.n0 + g * 64, tc.x0, yh, tc.img);
              else tma_store_4d(&p.tmOut, sStage + g * kTileBytes, tc.n0 + g * 64, tc.x0, yh, tc.img);
            }
            tma_store_commit();
            if (h == MH - 1) YX_TRACE(5, t);
          }
        }
      }
    }
    if (lead_warp) {
      if (elect_one()) tma_store_wait_all0();
      YX_TRACE_SUM(TW_EPI_TFULL, w_acc0);
      YX_TRACE_SUM(TW_EPI_STAGE, w_acc1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still touch it
  if (warp == 2) {
    if (PAIR) tmem_dealloc_2sm(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
// Build layout: this file is compiled once as it is (planning + dispatch) and once per activation with
// -DYX_CONV_ACT_SLICE=<yx_act value> (the ~20 kernel instantiations of that activation and their launcher), so the
// instantiations compile in parallel (_build.py).
#ifndef YX_CONV_ACT_SLICE
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
                      const uint32_t* box, bool swizzle128, const char* what, bool swizzle64 = false) {
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
                   swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : (swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
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

static void choose_tile(int H, int W, int* th, int* tw, bool even = false, int px = 128) {
  // minimise padded MMA rows; ties -> squarer tile (fewer halo re-reads from L2)
  double best = 1e30;
  int bh = even ? 2 : 1, bw = even ? 2 : 1;
  for (int w = 1; w <= std::min(W, px); ++w) {
    int h = std::min(H, px / w);
    if (h > 256 || w > 256) continue;
    if (even) h &= ~1;
    if (even && (w & 1)) continue;
    if (h < 1) continue;
    double tiles = (double)ceil_div(H, h) * ceil_div(W, w);
    double cost = tiles * 1000.0 + std::abs(h - w) * 0.01;
    if (cost < best) { best = cost; bh = h; bw = w; }
  }
  *th = bh;
  *tw = bw;
}

struct ConvGeom {
  bool rowpack, has_res, has_up, img;
  int Hout, Wout, taps, cin_real;
  double flops, act_bytes;
};

static int conv_geom(const yx_op& op, ConvGeom* g) {
  const yx_view& s = op.src;
  const yx_view& d = op.dst;
  YX_REQUIRE(op.ksize == 1 || op.ksize == 3 || (op.ksize == 4 && op.stride == 2), "conv ksize must be 1, 3, or 4 with stride 2");
  YX_REQUIRE(op.stride == 1 || op.stride == 2, "conv stride must be 1 or 2");
  YX_REQUIRE(op.cin_pad % 16 == 0 && op.cout_pad % 16 == 0, "cin_pad/cout_pad must be multiples of 16");
  YX_REQUIRE(s.c % 8 == 0, "src channels must be a multiple of 8");
  YX_REQUIRE(op.aux == 0 || op.aux == 1 || op.aux == 4 || op.aux == 12,
             "conv aux must be 0, 1 (row-packed), or 4 / 12 (stem fed from the image: Focus / pixel_unshuffle order)");
  YX_REQUIRE(d.c % 8 == 0 && d.c <= op.cout_pad, "dst channels must be a multiple of 8 and <= cout_pad");
  YX_REQUIRE(s.pitch % 8 == 0 && d.pitch % 8 == 0 && s.offset % 16 == 0 && d.offset % 16 == 0 && s.nstride % 8 == 0 &&
                 d.nstride % 8 == 0,
             "views must be 16-byte aligned");
  // aux == 1: "row-packed" 3x3 conv over a 16-channel tensor stored with 1 zero column on the left and 3 on
  // the right (the s2d output).  TMA reads it through an OVERLAPPING view (64 channels per pixel, pixel pitch 16),
  // so the 128-byte smem row of pixel x holds pixels x-1..x+2 = the three horizontal taps (+1 ignored): the
  // conv becomes 3 vertical taps with K = 48 instead of 9 taps with K = 16, and every TMA row is a full line.
  g->rowpack = (op.aux & 1) != 0;
  // aux & 4: a 3x3 stem conv FED FROM THE IMAGE -- the kernel's producer warps do the space-to-depth themselves, there is
  // no s2d tensor (op.src is not read; the image pointer arrives with every launch); aux & 8: pixel_unshuffle order
  g->img = (op.aux & 4) != 0;
  if (g->rowpack)
    YX_REQUIRE(op.ksize == 3 && op.stride == 1 && s.c == 16 && s.pitch == 16 && op.cin_pad == 48 && s.w > 4 && op.res.c == 0,
               "row-packed conv needs k=3, s=1, a padded 16-channel source and cin_pad = 48");
  if (g->img)
    YX_REQUIRE(op.ksize == 3 && op.stride == 1 && op.cin_pad == 16 && op.res.c == 0 && op.up.c == 0 && d.w % 8 == 0 && d.h >= 16 && d.w >= 8,
               "image-fed stem needs k=3, s=1, cin_pad = 16, no residual, an image width that is a multiple of 16 and a map of >= 16x8");
  const int pad = (op.ksize - 1) / 2;  // BaseConv: pad = (ksize - 1) // 2 (network_blocks.py:54); 4x4/s2 (P6-v2) pads 1
  g->Hout = g->img ? d.h : (g->rowpack ? s.h : (s.h + 2 * pad - op.ksize) / op.stride + 1);
  g->Wout = g->img ? d.w : (g->rowpack ? s.w - 4 : (s.w + 2 * pad - op.ksize) / op.stride + 1);
  YX_REQUIRE(d.h == g->Hout && d.w == g->Wout && d.n == s.n, "dst spatial dims do not match the conv geometry");
  g->has_res = op.res.c > 0;
  if (g->has_res)
    YX_REQUIRE(op.res.h == d.h && op.res.w == d.w && op.res.c == d.c && op.res.n == d.n && op.res.pitch % 8 == 0 &&
                   op.res.offset % 16 == 0 && op.res.nstride % 8 == 0,
               "residual view must match dst");
  g->has_up = op.up.c > 0;
  if (g->has_up) {
    YX_REQUIRE(op.ksize == 1 && op.stride == 1 && !g->rowpack, "fused upsample needs a 1x1 stride-1 conv");
    YX_REQUIRE(op.up.c % 64 == 0 && op.up.n == s.n && op.up.h * 2 == s.h && op.up.w * 2 == s.w && s.h % 2 == 0 && s.w % 2 == 0,
               "fused upsample: up must be [n, h/2, w/2, 64*k]");
    YX_REQUIRE(op.up.pitch % 8 == 0 && op.up.offset % 16 == 0 && op.up.nstride == (int64_t)op.up.h * op.up.w * op.up.pitch,
               "fused upsample: the low-resolution tensor must be 16-byte aligned with contiguous images");
    YX_REQUIRE(op.cin_pad == op.up.c + round_up(s.c, 16), "fused upsample: cin_pad must be up.c + padded src.c");
  } else {
    YX_REQUIRE(s.c <= op.cin_pad, "src channels must be <= cin_pad");
  }
  g->taps = g->rowpack ? 3 : op.ksize * op.ksize;
  g->cin_real = (g->rowpack || g->img) ? 12 : s.c + (g->has_up ? op.up.c : 0);
  const double px_out = (double)d.n * g->Hout * g->Wout;
  g->flops = 2.0 * px_out * d.c * g->cin_real * op.ksize * op.ksize;
  g->act_bytes = 2.0 * ((double)s.n * s.h * s.w * (g->rowpack ? 12 : s.c) + px_out * d.c * (g->has_res ? 2 : 1));
  if (g->img) g->act_bytes = px_out * 12.0 * 2.0 + 2.0 * px_out * d.c;   // the image once (counted as fp16, like the s2d op) + the output
  if (g->has_up) g->act_bytes += 2.0 * (double)op.up.n * op.up.h * op.up.w * op.up.c;  // read once at LOW resolution
  return YX_OK;
}

static bool halo_ok(const yx_op& op, const ConvGeom& g) {
  if (!(op.ksize == 3 && op.stride == 1 && g.Hout >= 16 && g.Wout >= 8)) return false;  // (row-packed stem: vertical halo)
  const double eff16 = (double)(ceil_div(g.Hout, 16) * 16) * (ceil_div(g.Wout, 8) * 8) / ((double)g.Hout * g.Wout);
  return eff16 <= 1.25;
}

constexpr int kSpMetaCol0 = 256, kSpMetaColsMax = 256;  // TMEM: two 128-column accumulators, then the metadata

// Geometry the 2:4 sparse variant supports: a plain conv (no row-packed stem, no fused upsample) whose input channels are
// whole K = 32 steps and whose metadata for all output-channel tiles fits the 256 TMEM columns behind the accumulators.
bool sparse_shape_ok(const yx_op& op) {
  if (op.kind != YX_OP_CONV || op.aux != 0 || op.up.c > 0) return false;
  if (op.cin_pad % 32 != 0) return false;
  const int taps = op.ksize * op.ksize, n_mt = ceil_div(op.cout_pad, 128);
  return n_mt * taps * (op.cin_pad / 32) <= kSpMetaColsMax;
}

// YX_SPARSE: "0" never use the sparse variant, "force" use it wherever the weights allow (tests, experiments); otherwise
// the tuner decides per layer by measurement
static int sparse_env() {   // read on every call: tests switch it between engines
  const char* e = getenv("YX_SPARSE");
  return !e ? 1 : (strcmp(e, "0") == 0 ? 0 : (strcmp(e, "force") == 0 ? 2 : 1));
}

static ConvTune sparse_tune(const yx_op& op, const ConvGeom& g, int epi_groups, int stage_bufs);

// The heuristic (untuned) launch shape.
static ConvTune default_tune(const yx_op& op, const ConvGeom& g, bool sparse_ok = false) {
  static const int halo_env = getenv("YX_HALO") ? atoi(getenv("YX_HALO")) : 1;
  if (sparse_ok && sparse_env() == 2) return sparse_tune(op, g, 1, 1);
  ConvTune t;
  memset(&t, 0, sizeof t);
  const int cout16 = op.cout_pad;
  const bool mem_bound = g.flops / g.act_bytes < mem_bound_ai();
  t.epi_groups = 1;
  t.stage_bufs = 2;
  if (g.img) {   // the image-fed stem only exists as the vertical-halo shape
    // two stacked halves while the nine resident weight taps, two 43 KB halo stages and the two raw patches fit (cout <= 48)
    t.variant = 2; t.bn = cout16; t.mh = cout16 <= 48 ? 2 : 1; t.ctas = 1; t.w3 = 1; t.epi_groups = cout16 <= 48 ? 2 : 1;
    t.stage_bufs = cout16 <= 64 ? 2 : 1;
    return t;
  }
  if (halo_env && halo_ok(op, g) && op.src.c <= 96) {
    t.variant = 2;
    t.bn = cout16 <= 256 ? cout16 : round_up(ceil_div(cout16, ceil_div(cout16, 256)), 64);
    t.mh = (cout16 <= 128 && ceil_div(g.Hout, 32) * 32 <= 1.2 * ceil_div(g.Hout, 16) * 16) ? 2 : 1;
    t.ctas = 1;
    t.w3 = 2;
    return t;
  }
  t.variant = 1;
  if (mem_bound) {
    t.bn = cout16 <= 128 ? cout16 : 128;
    t.ctas = 2;
  } else {
    t.bn = cout16 <= 256 ? cout16 : round_up(ceil_div(cout16, ceil_div(cout16, 256)), 64);
    t.ctas = 1;
  }
  t.w3 = 1;
  return t;
}

static ConvTune sparse_tune(const yx_op& op, const ConvGeom& g, int epi_groups, int stage_bufs) {
  ConvTune t;
  memset(&t, 0, sizeof t);
  t.variant = halo_ok(op, g) ? 2 : 1;
  t.bn = 128; t.ctas = 1; t.mh = 1; t.epi_groups = epi_groups; t.stage_bufs = stage_bufs; t.w3 = t.variant == 2 ? 2 : 1;
  t.sparse = 1;
  return t;
}

void conv_candidates(const yx_op& op, std::vector<ConvTune>* out, bool sparse_ok) {
  out->clear();
  ConvGeom g;
  if (conv_geom(op, &g) != YX_OK) return;
  out->push_back(default_tune(op, g, sparse_ok));
  if (g.img) {
    for (int mh = 1; mh <= 2; ++mh)
      for (int eg = 1; eg <= 2; ++eg)
        for (int sb = 2; sb >= 1; --sb)
          for (int alt = 0; alt <= ((eg == 2 && sb == 2) ? 1 : 0); ++alt) {
            ConvTune t = (*out)[0];
            t.mh = mh; t.epi_groups = eg; t.stage_bufs = sb; t.epi_alt = alt;
            bool dup = false;
            for (const ConvTune& o : *out) dup = dup || memcmp(&o, &t, sizeof t) == 0;
            if (!dup) out->push_back(t);
          }
    return;
  }
  if (sparse_ok && sparse_env() == 2) return;   // forced: the sparse shape is the only candidate
  if (sparse_ok && sparse_env() == 1)
    for (int eg = 1; eg <= 2; ++eg)
      for (int sb = 1; sb <= 2; ++sb) {
        out->push_back(sparse_tune(op, g, eg, sb));
        if (halo_ok(op, g)) {   // also the generic (one box per tap) sparse shape
          ConvTune t = sparse_tune(op, g, eg, sb);
          t.variant = 1; t.w3 = 1;
          out->push_back(t);
        }
      }
  const int cout16 = op.cout_pad;
  std::vector<int> bns;
  auto add_bn = [&](int b) {
    if (b >= 16 && b <= 256 && b % 16 == 0 && b <= cout16 && (b % 64 == 0 || b == cout16) &&
        std::find(bns.begin(), bns.end(), b) == bns.end())
      bns.push_back(b);
  };
  add_bn(std::min(cout16, 256));
  if (cout16 > 256) add_bn(round_up(ceil_div(cout16, ceil_div(cout16, 256)), 64));
  if (cout16 > 128) add_bn(128);
  if (cout16 > 192) add_bn(192);
  if (cout16 > 64 && cout16 % 64 == 0 && cout16 <= 128) add_bn(64);
  {  // latency regime (bs1, deep 20x20 / 40x40 maps): too few tiles to fill the SMs -> narrower N tiles spread the layer
    const int64_t m_tiles = (int64_t)op.dst.n * ceil_div(op.dst.h * op.dst.w, 128);
    if (cout16 > 64 && m_tiles * ceil_div(cout16, 128) < 148) add_bn(64);
  }
  auto push1 = [&](const ConvTune& t) {
    for (const ConvTune& o : *out)
      if (memcmp(&o, &t, sizeof t) == 0) return;
    out->push_back(t);
  };
  auto push = [&](ConvTune t) {
    push1(t);
    if (t.epi_groups == 2 && t.stage_bufs == 2) {   // the same shape with the two groups alternating tiles
      t.epi_alt = 1;
      push1(t);
    }
  };
  for (int bn : bns) {
    for (int ctas = 1; ctas <= 2; ++ctas)
      for (int eg = 1; eg <= 2; ++eg)
        for (int sb = 2; sb >= 1; --sb) {
          if (ctas == 2 && (bn > 128 || sb == 1 || eg == 2)) continue;  // 2 x 384 threads x ~90 regs exceed the register file
          if (sb == 1 && bn < 96) continue;  // small tiles: the second staging buffer is cheap, keep it
          ConvTune t;
          memset(&t, 0, sizeof t);
          t.variant = 1; t.bn = bn; t.ctas = ctas; t.epi_groups = eg; t.stage_bufs = sb; t.w3 = 1;
          push(t);
          if (bn <= 128 && ctas == 1) {  // 256-pixel tiles (also with a fused upsample: the 5-D box simply covers TH x TW)
            t.mh = 2;
            push(t);
          }
        }
    if (bn >= 64 && cout16 >= 96 && !g.rowpack)  // CTA-pair shapes: half of the weight tile per CTA
      for (int v = 1; v <= (halo_ok(op, g) ? 2 : 1); ++v)
        for (int eg = 1; eg <= (bn <= 128 ? 2 : 1); ++eg)  // short MMA phases (N <= 128) can be epilogue-bound
          for (int sb = 2; sb >= 1; --sb) {
            ConvTune t;
            memset(&t, 0, sizeof t);
            t.variant = v; t.bn = bn; t.ctas = 1; t.mh = 1; t.epi_groups = eg; t.stage_bufs = sb; t.w3 = v == 2 ? 2 : 1; t.pair = 1;
            push(t);
          }
    if (halo_ok(op, g))
      for (int mh = 1; mh <= 2; ++mh)
        for (int eg = 1; eg <= 2; ++eg)
          for (int sb = 2; sb >= 1; --sb) {
            if (sb == 1 && bn < 128) continue;
            if (mh == 2 && bn > 128) continue;  // 2 halves x 2 accumulators x BN columns must fit 512
            ConvTune t;
            memset(&t, 0, sizeof t);
            t.variant = 2; t.bn = bn; t.ctas = 1; t.mh = mh; t.epi_groups = eg; t.stage_bufs = sb; t.w3 = 2;
            push(t);
          }
  }
}

int conv_plan(const yx_op& op, void* base, const void* weights, const void* biases, int num_sms, const ConvTune* tune,
              ConvPlan* out, const SparseWeights* spw) {
  const yx_view& s = op.src;
  const yx_view& d = op.dst;
  ConvGeom g;
  int rc = conv_geom(op, &g);
  if (rc != YX_OK) return rc;
  const bool sparse_ok = spw != nullptr && spw->wc != nullptr && spw->meta != nullptr && sparse_shape_ok(op);
  const ConvTune t = tune ? *tune : default_tune(op, g, sparse_ok);
  const bool halo = t.variant == 2;
  const bool pair = t.pair != 0;
  const bool sp = t.sparse != 0;
  YX_REQUIRE(!sp || sparse_ok, "conv tune: the sparse variant needs 2:4-compliant packed weights and a supported geometry");
  YX_REQUIRE(!sp || (!pair && t.mh != 2 && t.ctas == 1 && t.bn == 128),
             "conv tune: the sparse variant runs one CTA per SM with 128-channel M tiles and one 128-pixel tile");
  YX_REQUIRE(t.variant == 1 || t.variant == 2, "conv tune: variant must be 1 (generic) or 2 (halo)");
  YX_REQUIRE(!pair || (t.ctas == 1 && (!halo || t.mh != 2) && !g.rowpack),
             "conv tune: CTA-pair mode runs one CTA per SM and one 128-pixel half per CTA (not for the row-packed stem)");
  YX_REQUIRE(!halo || halo_ok(op, g), "conv tune: halo variant needs a 3x3 stride-1 conv on a map of at least 16x8");
  YX_REQUIRE(!g.img || (halo && !pair && !sp && t.ctas == 1), "conv tune: the image-fed stem runs as the vertical-halo shape, one CTA per SM");
  YX_REQUIRE(!g.has_up || !halo, "fused upsample is a 1x1 conv: generic variant only");
  YX_REQUIRE(t.bn >= 16 && t.bn <= 256 && t.bn % 16 == 0 && (t.bn % 64 == 0 || t.bn >= op.cout_pad),
             "conv tune: N tile must be a multiple of 64 (or the whole padded Cout), at most 256");
  YX_REQUIRE(t.ctas == 1 || t.ctas == 2, "conv tune: ctas per SM must be 1 or 2");
  YX_REQUIRE(t.epi_groups == 1 || t.epi_groups == 2, "conv tune: epilogue groups must be 1 or 2");
  YX_REQUIRE(!t.epi_alt || (t.epi_groups == 2 && t.stage_bufs != 1),
             "conv tune: alternating epilogue groups need two groups and two staging buffers");
  YX_REQUIRE(op.cout_pad <= 4096, "cout too large");

  ConvPlan pl;
  memset(&pl, 0, sizeof pl);
  pl.tune = t;
  ConvParams& p = pl.p;
  const int pad = (op.ksize - 1) / 2;
  p.ksize = op.ksize; p.stride = op.stride; p.act = op.act;
  const bool inplace = g.has_res && op.res.offset == d.offset && op.res.nstride == d.nstride && op.res.pitch == d.pitch;
  p.has_res = !g.has_res ? 0 : (inplace ? 2 : 1);
  p.ky = op.ksize; p.kx = g.rowpack ? 1 : op.ksize;
  p.pad_y = pad; p.pad_x = g.rowpack ? 0 : pad;
  p.cin = op.cin_pad;
  p.cout16 = op.cout_pad;
  p.up_chunks = g.has_up ? op.up.c / 64 : 0;
  p.up_h = g.has_up ? op.up.h : 0;
  p.k_chunks = p.up_chunks + ceil_div(p.cin - p.up_chunks * 64, 64);
  p.BN = sp ? 128 : std::min(t.bn, p.cout16);   // sparse: the M = 128 rows of the sparse operand (a partial last tile is padded)
  p.n_tiles_n = ceil_div(p.cout16, p.BN);
  p.halo = halo ? 1 : 0;
  p.rowpack = g.rowpack ? 1 : 0;
  p.pair = pair ? 1 : 0;
  p.img_fused = g.img ? 1 : 0;
  p.img_h = 2 * d.h; p.img_w = 2 * d.w; p.img_order = (op.aux & 8) ? 1 : 0;
  p.sp = sp ? 1 : 0;
  p.sp_meta = sp ? spw->meta : nullptr;
  p.sp_cols_per_tile = g.taps * (op.cin_pad / 32);
  p.sp_meta_col0 = kSpMetaCol0;
  p.epi_groups = t.epi_groups;
  p.epi_alt = t.epi_alt ? 1 : 0;
  p.bias_bytes = round_up(p.cout16 * 4, 128);
  p.b_stage_bytes = sp ? 128 * 64 : (pair ? p.BN / 2 : p.BN) * 128;  // pair: each CTA holds half of the N tile's weight rows;
                                                                      // sparse: 128 rows of 32 stored fp16 (one 64-channel chunk)
  p.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(biases) + op.b_offset);
  static const int diag_env = getenv("YX_CONV_DIAG") ? atoi(getenv("YX_CONV_DIAG")) : 0;  // experiments only (results are garbage)
  p.diag = diag_env;
  const int taps = g.taps;
  const int groups64 = ceil_div(p.BN, 64);
  int stride_cols = 32;
  while (stride_cols < p.BN) stride_cols <<= 1;   // (sparse: BN = 128 = the pixel columns of one transposed accumulator)
  int budget = t.ctas == 2 ? kSmemTwoCtas : kSmemLimit;

  if (halo) {
    int mh = t.mh == 2 ? 2 : 1;
    if (2 * mh * stride_cols > 512) mh = 1;
    p.mh = mh;
    p.TH = 16 * mh; p.TW = 8;
    p.a_box_bytes = (p.TH + 2) * (g.rowpack ? 8 : kHaloW) * 128;
    p.a_stage_bytes = round_up(p.a_box_bytes, 1024);
    p.out_box_bytes = kTileBytes;
    p.acc_stride = stride_cols;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * mh * stride_cols) p.tmem_cols <<= 1;
  } else if (t.mh == 2) {
    // generic with a 256-pixel tile: two halves of (TH/2) x TW pixels stacked in one A box
    YX_REQUIRE(!pair && 4 * stride_cols <= 512 && t.ctas == 1,
               "conv tune: the 256-pixel generic tile needs N <= 128, one CTA per SM and no CTA pair");
    p.mh = 2;
    choose_tile(g.Hout, g.Wout, &p.TH, &p.TW, true, 256);
    YX_REQUIRE(p.TH % 2 == 0 && p.TH / 2 * p.TW <= 128 && (p.TH / 2 * p.TW) % 8 == 0 && p.TH / 2 * p.TW >= 64,
               "conv tune: the map does not tile into two aligned 128-row halves");
    p.a_box_bytes = p.TH * p.TW * 128;
    p.a_stage_bytes = 2 * kTileBytes;
    p.out_box_bytes = p.a_box_bytes / 2;
    p.acc_stride = stride_cols;
    p.tmem_cols = 32;
    while (p.tmem_cols < 4 * stride_cols) p.tmem_cols <<= 1;
  } else {
    p.mh = 1;
    choose_tile(g.Hout, g.Wout, &p.TH, &p.TW, g.has_up);
    p.a_box_bytes = p.TH * p.TW * 128;
    p.a_stage_bytes = kTileBytes;
    p.out_box_bytes = p.a_box_bytes;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * stride_cols) p.tmem_cols <<= 1;
    p.acc_stride = p.tmem_cols / 2;
  }
  if (sp) {   // two 128-column accumulators at columns 0 / 128, metadata from column 256 on
    p.acc_stride = 128;
    p.tmem_cols = 512;
  }
  YX_REQUIRE(t.ctas == 1 || p.tmem_cols <= 256, "conv tune: two CTAs per SM need <= 256 TMEM columns each");
  p.tiles_h = ceil_div(g.Hout, p.TH);
  p.tiles_w = ceil_div(g.Wout, p.TW);
  if (pair) p.tiles_w = ceil_div(p.tiles_w, 2);  // iteration space in pair tiles: CTA `rank` owns x-tile 2*tx + rank
  p.n_tiles_m = d.n * p.tiles_h * p.tiles_w;

  // ---- shared-memory budget: [A ring][B ring or resident B][staging x stage_bufs][bias][barriers] ----
  p.stage_bufs = t.stage_bufs == 1 ? 1 : 2;
  const int k_loads_b = taps * p.k_chunks;  // B tiles per output tile
  const int tap_stage_bytes = p.b_stage_bytes;   // one tap's weight tile (of this CTA)
  p.b_taps = 1;
  for (;;) {
    p.b_stage_bytes = tap_stage_bytes;
    p.b_taps = 1;
    // (image-fed stem: + two raw image patches (this tile's and the next one's): 3 planes x 2(TH+2) rows x 128 bytes each,
    // + the 512-byte uint8 value table)
    const int fixed = 1024 + kBarBytes + p.bias_bytes + p.stage_bufs * groups64 * kTileBytes +
                      (g.img ? 2 * (3 * 2 * (p.TH + 2) * 128) + 512 : 0);
    const int avail = budget - fixed;
    const int min_a = (halo ? 2 : 2) * p.a_stage_bytes;
    // (pair: each CTA keeps ITS half of the weight rows resident -> the 96-channel 3x3 layers, whose 166 KB of weights
    // never fit one CTA, run with resident weights at the full N/2 MMA rate)
    const bool can_res = p.n_tiles_n == 1 && k_loads_b <= kMaxBRing && t.no_resident == 0 &&
                         k_loads_b * p.b_stage_bytes + min_a <= avail;
    if (can_res) {
      p.b_resident = 1;
      p.b_slots = k_loads_b;
      p.stages_a = std::min(halo ? 3 : kMaxARing, (avail - k_loads_b * p.b_stage_bytes) / p.a_stage_bytes);
    } else if (halo) {
      p.b_resident = 0;
      // Streamed weights of a 3x3 halo conv: ONE ring stage per filter ROW (three taps, one TMA load, one wait and one
      // commit in the MMA warp) whenever two such stages fit beside two halo stages: the MMA warp's per-stage bookkeeping
      // (~420 cycles: wait, fence, elect, commit, ring advance) exceeded the 384 cycles four N = 192 MMAs take, so with
      // one tap per stage every streamed halo layer ran at the issue rate of that warp (tools/conv_trace.py).
      static const bool btaps_env = !(getenv("YX_BTAPS") && atoi(getenv("YX_BTAPS")) == 1);
      p.b_taps = 1;
      if (btaps_env && !sp && !g.rowpack && taps == 9) {
        const int b3 = 3 * tap_stage_bytes;
        const int try_ab[4][2] = {{3, 3}, {2, 3}, {3, 2}, {2, 2}};
        for (int i = 0; i < 4 && p.b_taps == 1; ++i)
          if (try_ab[i][0] * p.a_stage_bytes + try_ab[i][1] * b3 <= avail) {
            p.b_taps = 3;
            p.stages_a = try_ab[i][0];
            p.b_slots = try_ab[i][1];
          }
      }
      if (p.b_taps == 3) {
        p.b_stage_bytes = 3 * tap_stage_bytes;
      } else {
        p.b_stage_bytes = tap_stage_bytes;
        p.stages_a = (avail - 3 * p.a_stage_bytes >= 4 * p.b_stage_bytes) ? 3 : 2;
        p.b_slots = std::min(12, (avail - p.stages_a * p.a_stage_bytes) / p.b_stage_bytes);
      }
    } else {
      p.b_resident = 0;
      const int st = std::min(kMaxARing, avail / (p.a_stage_bytes + p.b_stage_bytes));
      p.stages_a = st;
      p.b_slots = st;
    }
    p.shared_ring = (!halo && !p.b_resident) ? 1 : 0;
    const bool fits = p.stages_a >= 2 && p.b_slots >= (p.b_resident ? 1 : ((halo && p.b_taps == 1) ? 3 : 2));
    if (fits) {
      pl.smem_bytes = fixed + p.stages_a * p.a_stage_bytes + p.b_slots * p.b_stage_bytes;
      break;
    }
    if (p.stage_bufs == 2 && !p.epi_alt) { p.stage_bufs = 1; continue; }
    set_error("conv plan: launch shape does not fit in shared memory");
    return YX_ERR_INVALID;
  }
  // never let an extra CTA (which would stall in tcgen05.alloc) fit on an SM
  if (t.ctas == 2) pl.smem_bytes = std::max(pl.smem_bytes, 80 * 1024);
  else if (p.tmem_cols > 256) pl.smem_bytes = std::max(pl.smem_bytes, 120 * 1024);
  else pl.smem_bytes = std::max(pl.smem_bytes, 80 * 1024);
  pl.grid = std::min(p.n_tiles_m * p.n_tiles_n, t.ctas * num_sms);
  if (pair) pl.grid = 2 * std::min(p.n_tiles_m * p.n_tiles_n, num_sms / 2);
  {
    // Unequal N tiles (Cout = 288 as 192 + 96): with the N tile fastest and a static stride of `units` tiles, a stride that
    // shares a factor with n_tiles_n pins every CTA to the SAME N tile for the whole launch -- half of the SMs would only
    // ever see the 192-wide tiles and the other half finish in half the time (the 288-channel 3x3 layers ran at 0.66 of the
    // tensor peak for exactly this reason).  A stride coprime to n_tiles_n rotates every CTA through all N tiles.
    static const bool rot_env = !(getenv("YX_NROT") && atoi(getenv("YX_NROT")) == 0);
    int units = pair ? pl.grid / 2 : pl.grid;
    auto gcd = [](int a, int b) { while (b) { const int r = a % b; a = b; b = r; } return a; };
    if (rot_env && p.n_tiles_n > 1 && p.cout16 % p.BN != 0 && p.n_tiles_m * p.n_tiles_n > units) {
      while (units > 1 && gcd(units, p.n_tiles_n) != 1) --units;
      pl.grid = pair ? 2 * units : units;
    }
  }
  {  // mixed-radix digits of the persistent-tile step (see TileIter)
    int st = pair ? pl.grid / 2 : pl.grid;
    p.step_nt = st % p.n_tiles_n; st /= p.n_tiles_n;
    p.step_x = st % p.tiles_w; st /= p.tiles_w;
    p.step_y = st % p.tiles_h;
    p.step_img = st / p.tiles_h;
  }
  pl.threads = 128 + 128 * p.epi_groups + (g.img ? 128 : 0);   // image-fed stem: + one warpgroup of tile builders
  // warp 3: second producer for the operand with more loads per tile (none when B is resident and A is one load)
  const int a_loads = halo ? p.k_chunks : k_loads_b;
  const int b_loads = p.b_resident ? 0 : k_loads_b / p.b_taps;
  p.w3_role = t.w3 == 0 ? 0 : (b_loads > a_loads ? 2 : 1);
  if (p.w3_role == 1 && p.stages_a < 2) p.w3_role = 0;
  if (p.w3_role == 2 && p.b_slots < 2) p.w3_role = 0;
  // warp 2 (weight producer) joins the activation producers once its resident weights are in (YX_W2A=0: off, for A/B runs)
  if (g.img) p.w3_role = 1;   // warps 0, 2, 3 build the operand tiles together; warp 2 first issues the (resident) weight loads
  static const bool w2a_env = !(getenv("YX_W2A") && atoi(getenv("YX_W2A")) == 0);
  p.w2_role = (w2a_env && p.b_resident && t.w3 != 0 && a_loads >= 3 && p.stages_a >= 3 && !g.img) ? 1 : 0;
  YX_REQUIRE(!g.img || p.b_resident, "image-fed stem: the weights must stay resident");

  if (g.img) {
    p.tmA[0] = p.tmA[1] = p.tmA[2] = p.tmA[3] = CUtensorMap{};   // no tensor map: the producers read the image themselves
  } else if (halo && g.rowpack) {
    uint64_t dims[4] = {64, (uint64_t)g.Wout, (uint64_t)s.h, (uint64_t)s.n};
    uint64_t st[4] = {2, 32, (uint64_t)s.w * 32, (uint64_t)s.nstride * 2};
    uint32_t box[4] = {64, 8, (uint32_t)(p.TH + 2), 1};
    if ((rc = encode_map(&p.tmA[0], static_cast<uint8_t*>(base) + s.offset, 4, dims, st, box, true, "A-rowpack-halo")) != YX_OK)
      return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else if (halo) {
    uint64_t dims[4] = {(uint64_t)s.c, (uint64_t)s.w, (uint64_t)s.h, (uint64_t)s.n};
    uint64_t st[4] = {2, (uint64_t)s.pitch * 2, (uint64_t)s.pitch * 2 * s.w, (uint64_t)s.nstride * 2};
    uint32_t box[4] = {64, (uint32_t)kHaloW, (uint32_t)(p.TH + 2), 1};
    if ((rc = encode_map(&p.tmA[0], static_cast<uint8_t*>(base) + s.offset, 4, dims, st, box, true, "A-halo")) != YX_OK)
      return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else if (g.rowpack) {
    uint64_t dims[4] = {64, (uint64_t)g.Wout, (uint64_t)s.h, (uint64_t)s.n};
    uint64_t st[4] = {2, 32, (uint64_t)s.w * 32, (uint64_t)s.nstride * 2};
    uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    if ((rc = encode_map(&p.tmA[0], static_cast<uint8_t*>(base) + s.offset, 4, dims, st, box, true, "A-rowpack")) != YX_OK)
      return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else if (op.stride == 1) {
    if ((rc = encode_view(&p.tmA[0], base, s, p.TW, p.TH, "A")) != YX_OK) return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
    if (g.has_up) {
      // (c, dup_x, w/2, dup_y, n*h/2): the two stride-0 dimensions repeat every low-resolution pixel 2x2, so the box
      // (64, 2, TW/2, 2, TH/2) lands in shared memory in exactly the high-resolution pixel order of the tile
      const yx_view& u = op.up;
      YX_REQUIRE(p.TW % 2 == 0 && p.TH % 2 == 0, "fused upsample needs an even tile");
      uint64_t dims[5] = {(uint64_t)u.c, 2, (uint64_t)u.w, 2, (uint64_t)u.n * u.h};
      uint64_t st[5] = {2, 0, (uint64_t)u.pitch * 2, 0, (uint64_t)u.pitch * 2 * u.w};
      uint32_t box[5] = {64, 2, (uint32_t)p.TW / 2, 2, (uint32_t)p.TH / 2};
      if ((rc = encode_map(&p.tmUp, static_cast<uint8_t*>(base) + u.offset, 5, dims, st, box, true, "A-upsample")) != YX_OK) return rc;
    }
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
  if (sp) {   // compressed weights [round_up(cout,128)][taps][cin/2]: rows of 32 stored fp16 per chunk, SWIZZLE_64B
    const int half = op.cin_pad / 2;
    uint64_t dims[3] = {(uint64_t)half, (uint64_t)taps, (uint64_t)round_up(op.cout_pad, 128)};
    uint64_t st[3] = {2, (uint64_t)half * 2, (uint64_t)half * 2 * taps};
    uint32_t box[3] = {32, 1, 128};
    if ((rc = encode_map(&p.tmW, const_cast<void*>(spw->wc), 3, dims, st, box, false, "W-sparse", true)) != YX_OK) return rc;
  } else {
    // memory order [cout][tap][cin]; the map puts the taps OUTERMOST so that one box can cover several taps of the same
    // rows and lands in shared memory tap by tap, each tap a complete K-major weight tile
    uint64_t dims[3] = {(uint64_t)op.cin_pad, (uint64_t)op.cout_pad, (uint64_t)taps};
    uint64_t st[3] = {2, (uint64_t)op.cin_pad * 2 * taps, (uint64_t)op.cin_pad * 2};
    uint32_t box[3] = {64, (uint32_t)(pair ? p.BN / 2 : p.BN), (uint32_t)p.b_taps};
    uint8_t* addr = const_cast<uint8_t*>(static_cast<const uint8_t*>(weights)) + op.w_offset;
    YX_REQUIRE(op.w_offset % 16 == 0, "weight offset must be 16-byte aligned");
    if ((rc = encode_map(&p.tmW, addr, 3, dims, st, box, true, "W")) != YX_OK) return rc;
  }
  const int store_th = halo ? 16 : p.TH / p.mh;  // stacked halves are stored one at a time
  if ((rc = encode_view(&p.tmOut, base, d, p.TW, store_th, "out")) != YX_OK) return rc;
  if (g.has_res) {
    if ((rc = encode_view(&p.tmRes, base, op.res, p.TW, store_th, "res")) != YX_OK) return rc;
  } else {
    p.tmRes = p.tmOut;
  }
  pl.flops = g.flops;
  pl.bytes = g.act_bytes + 2.0 * (double)d.c * g.cin_real * op.ksize * op.ksize;
  snprintf(pl.desc, sizeof pl.desc, "%s%s%s%s%s BN%d%s mh%d ctas%d epi%d%s sbuf%d A%dx%dK B%d%s%s w3:%d%s grid%d smem%dK", g.img ? "image-fed-" : "", sp ? "sparse24-" : "", p.has_res == 2 ? "inplace-" : "", pair ? "pair-" : "", halo ? "halo" : "generic",
           p.BN, p.n_tiles_n > 1 ? "*" : "", p.mh, t.ctas, p.epi_groups, p.epi_alt ? "alt" : "", p.stage_bufs, p.stages_a, p.a_stage_bytes >> 10, p.b_slots,
           p.b_resident ? "res" : "", p.b_taps == 3 ? "x3" : "", p.w3_role, p.w2_role ? "+w2" : "", pl.grid, pl.smem_bytes >> 10);
  *out = pl;
  return YX_OK;
}

// Binds the caller's image to an image-fed stem plan: the raw-patch tensor map (W, H, 3, B) of the image dtype with box
// (128 bytes, 2(TH+2) rows, 3 planes, 1), out-of-image elements zero-filled (the builder never uses them: it zeroes
// out-of-image s2d pixels AFTER the input affine).  Called once per launch (the image pointer is the caller's).
int conv_bind_image(ConvPlan* plan, const void* image, int image_dtype, float scale, float shift) {
  ConvParams& p = plan->p;
  YX_REQUIRE(p.img_fused, "not an image-fed plan");
  YX_REQUIRE(image != nullptr && (reinterpret_cast<uintptr_t>(image) & 15) == 0, "image pointer must be 16-byte aligned");
  uint64_t es;
  switch (image_dtype) {
    case YX_U8: es = 1; break;
    case YX_F16: es = 2; break;
    case YX_F32: es = 4; break;
    default: set_error("image dtype must be YX_F16, YX_F32 or YX_U8"); return YX_ERR_INVALID;
  }
  YX_REQUIRE(((uint64_t)p.img_w * es) % 16 == 0, "image-fed stem: image rows must be a multiple of 16 bytes");
  const int B = p.n_tiles_m / (p.tiles_h * p.tiles_w);
  // (W * es / 2 fp16-sized elements, H, 3, B); box (64 elements = 128 bytes, 2(TH+2) rows, 3 planes, 1), SWIZZLE_128B
  uint64_t dims[4] = {(uint64_t)p.img_w * es / 2, (uint64_t)p.img_h, 3, (uint64_t)B};
  uint64_t st[4] = {2, (uint64_t)p.img_w * es, (uint64_t)p.img_w * p.img_h * es, (uint64_t)p.img_w * p.img_h * 3 * es};
  uint32_t box[4] = {64, (uint32_t)(2 * (p.TH + 2)), 3, 1};
  int rc = encode_map(&p.tmImg, const_cast<void*>(image), 4, dims, st, box, true, "image");
  if (rc != YX_OK) return rc;
  p.img = image; p.img_dtype = image_dtype; p.img_scale = scale; p.img_shift = shift;
  p.img_affine = (scale != 1.0f || shift != 0.0f) ? 1 : 0;
  return YX_OK;
}

#endif  // !YX_CONV_ACT_SLICE

#ifdef YX_CONV_ACT_SLICE
template <int ACT, int RES, int MODE, bool PAIR, bool SP = false, bool IMG = false>
static int launch_variant(const ConvPlan& plan, cudaStream_t stream) {
  static bool attr_set = false;
  auto kernel = conv_gemm_kernel<ACT, RES, MODE, PAIR, SP, IMG>;
  if (!attr_set) {
    YX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
    attr_set = true;
  }
  static const bool pdl = !(getenv("YX_PDL") && atoi(getenv("YX_PDL")) == 0);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(plan.grid, 1, 1);
  cfg.blockDim = dim3(plan.threads, 1, 1);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;  // the CTA pair
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl && !plan.no_pdl) {  // may start while the previous kernel of the stream drains; the kernel orders itself with griddepcontrol.wait
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  YX_CUDA(cudaLaunchKernelEx(&cfg, kernel, plan.p));
  return YX_OK;
}

template <int ACT, int RES>
static int launch_mode(const ConvPlan& plan, cudaStream_t stream) {
  if (plan.p.sp)
    return plan.p.halo ? launch_variant<ACT, RES, 1, false, true>(plan, stream) : launch_variant<ACT, RES, 0, false, true>(plan, stream);
  if (plan.p.pair)
    return plan.p.halo ? launch_variant<ACT, RES, 1, true>(plan, stream) : launch_variant<ACT, RES, 0, true>(plan, stream);
  if (plan.p.img_fused) {
    if (plan.p.img == nullptr || (reinterpret_cast<uintptr_t>(plan.p.img) & 15) != 0) {
      set_error("image-fed stem: the image pointer must be set and 16-byte aligned");
      return YX_ERR_INVALID;
    }
    return plan.p.mh == 2 ? launch_variant<ACT, 0, 2, false, false, true>(plan, stream)
                          : launch_variant<ACT, 0, 1, false, false, true>(plan, stream);
  }
  if (plan.p.halo && plan.p.rowpack)  // the stem has no residual: only RES = 0 is instantiated
    return plan.p.mh == 2 ? launch_variant<ACT, 0, 5, false>(plan, stream) : launch_variant<ACT, 0, 4, false>(plan, stream);
  if (plan.p.halo && plan.p.mh == 2) return launch_variant<ACT, RES, 2, false>(plan, stream);
  if (plan.p.halo) return launch_variant<ACT, RES, 1, false>(plan, stream);
  if (plan.p.mh == 2) return launch_variant<ACT, RES, 3, false>(plan, stream);
  return launch_variant<ACT, RES, 0, false>(plan, stream);
}

template <int ACT>
static int launch_act(const ConvPlan& plan, cudaStream_t stream) {
  // plan.store_only (tuning / profiling of an in-place residual conv): plain stores, so repeated launches do not accumulate
  const int res = (plan.p.has_res == 2 && plan.store_only) ? 0 : plan.p.has_res;
  if (res == 2) return launch_mode<ACT, 2>(plan, stream);
  return res == 1 ? launch_mode<ACT, 1>(plan, stream) : launch_mode<ACT, 0>(plan, stream);
}

#define YX_CAT2(a, b) a##b
#define YX_CAT(a, b) YX_CAT2(a, b)
int YX_CAT(conv_launch_act_, YX_CONV_ACT_SLICE)(const ConvPlan& plan, cudaStream_t stream) {
  return launch_act<YX_CONV_ACT_SLICE>(plan, stream);
}
#else   // planning TU: dispatch to the per-activation launchers
int conv_launch_act_0(const ConvPlan& plan, cudaStream_t stream);
int conv_launch_act_1(const ConvPlan& plan, cudaStream_t stream);
int conv_launch_act_2(const ConvPlan& plan, cudaStream_t stream);
int conv_launch_act_3(const ConvPlan& plan, cudaStream_t stream);
int conv_launch_act_4(const ConvPlan& plan, cudaStream_t stream);

int conv_launch(const ConvPlan& plan, cudaStream_t stream) {
  switch (plan.p.act) {
    case YX_ACT_NONE: return conv_launch_act_0(plan, stream);
    case YX_ACT_SILU: return conv_launch_act_1(plan, stream);
    case YX_ACT_HSWISH: return conv_launch_act_2(plan, stream);
    case YX_ACT_RELU: return conv_launch_act_3(plan, stream);
    case YX_ACT_LRELU: return conv_launch_act_4(plan, stream);
    default: set_error("unknown activation code"); return YX_ERR_INVALID;
  }
}
#endif  // YX_CONV_ACT_SLICE

}  // namespace yx
