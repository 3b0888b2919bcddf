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

// Drains one 128-lane accumulator: TMEM -> registers (16 columns at a time) -> +bias -> round to fp16 ->
// activation (fp32) -> (+ residual already sitting in the staging line) -> fp16 -> 128-byte-swizzled smem
// staging [group of 64 ch][row][128 B] that the TMA store reads.  `bias` points at this N tile's first channel.
template <int ACT, bool HAS_RES>
__device__ __forceinline__ void epilogue_convert(uint32_t taddr, int bn_cur, int row, bool row_valid, uint32_t sStage,
                                                 const float* __restrict__ bias) {
  for (int c0 = 0; c0 < bn_cur; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(taddr + c0, v);
    tmem_ld_wait();
    if (row_valid) {
      float bb[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + j);
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
      uint32_t it = 0, s = 0, ph = 0;  // ring slot / phase kept incrementally (no div/mod on the issue path)
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int y0 = (r / p.tiles_w) * p.TH, x0 = (r % p.tiles_w) * p.TW;
        const int n0 = nt * p.BN;
        int dy = 0, dx = 0;
        for (int tap = 0; tap < taps; ++tap) {
          int mi = 0, cx, cy;
          if (p.stride == 1) {
            cx = x0 + dx - p.pad_x;
            cy = y0 + dy - p.pad_y;
          } else {
            // input row 2*y + dy - pad.  For k=3,pad=1: dy=0 -> odd row of cell y-1; dy=1 -> even row
            // of cell y; dy=2 -> odd row of cell y.  k=1 (pad 0): even row/col of cell y.
            const int oy = dy - p.pad_y, ox = dx - p.pad_y;
            const int py = oy & 1, px = ox & 1;
            mi = py * 2 + px;
            cy = y0 + ((oy - py) >> 1);
            cx = x0 + ((ox - px) >> 1);
          }
          for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            if (p.noload && it >= (uint32_t)p.stages) {  // diagnostics: MMA rate with operands already resident
              mbar_arrive(bar_full + 8 * s);
            } else {
              mbar_expect_tx(bar_full + 8 * s, p.a_box_bytes + p.b_stage_bytes);
              tma_load_4d(sA + s * kAStageBytes, &p.tmA[mi], bar_full + 8 * s, kc * 64, cx, cy, img);
              tma_load_3d(sB + s * p.b_stage_bytes, &p.tmW, bar_full + 8 * s, kc * 64, tap, n0);
            }
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
          }
          if (++dx == p.kx) { dx = 0; ++dy; }
        }
        YX_TRACE(0, (tile - (int)blockIdx.x) / (int)gridDim.x);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      // The whole tensor pipe is fed by THIS thread: everything per k-iteration is incremental 32-bit math
      // (a div/mod + 64-bit descriptor rebuild per iteration cost ~830 cycles per 4 MMAs, 2x their execution
      // time — profiles/r01_conv_tile_trace.txt).
      uint32_t t = 0, s = 0, ph = 0;
      const uint32_t hi = sdesc_hi(1024);
      const uint32_t a_lo0 = sdesc_lo(sA), b_lo0 = sdesc_lo(sB);
      const uint32_t a_step = kAStageBytes >> 4, b_step = p.b_stage_bytes >> 4;
      uint32_t a_lo = a_lo0, b_lo = b_lo0;
      const int ks_last = (p.cin - (p.k_chunks - 1) * 64) >> 4;
      const int k_iters = taps * p.k_chunks;
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
        int kc = 0;
        for (int i = 0; i < k_iters; ++i) {
          mbar_wait(bar_full + 8 * s, ph);
          tc_fence_after();
          const int ksteps = (kc == p.k_chunks - 1) ? ks_last : 4;
          umma_f16_ss_lohi(d_tmem, a_lo, hi, b_lo, hi, idesc, accum);
          if (ksteps > 1) umma_f16_ss_lohi(d_tmem, a_lo + 2, hi, b_lo + 2, hi, idesc, 1u);
          if (ksteps > 2) umma_f16_ss_lohi(d_tmem, a_lo + 4, hi, b_lo + 4, hi, idesc, 1u);
          if (ksteps > 3) umma_f16_ss_lohi(d_tmem, a_lo + 6, hi, b_lo + 6, hi, idesc, 1u);
          accum = 1;
          umma_commit(bar_empty + 8 * s);  // frees the smem stage when these MMAs retire
          a_lo += a_step; b_lo += b_step;
          if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; a_lo = a_lo0; b_lo = b_lo0; }
          if (++kc == p.k_chunks) kc = 0;
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
      epilogue_convert<ACT, HAS_RES>(taddr, bn_cur, row, row_valid, sStage, p.bias + n0);
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
// 3x3 / stride-1 variant with HALO REUSE.
// The generic kernel re-fetches every activation byte 9x (one shifted box per tap) and each of those
// fetches is also a 16 KB shared-memory write; ncu shows the tensor-bound layers limited by exactly that
// (shared-memory fill + operand reads ~ 200 B/cycle/SM against 128).  Here ONE (TH+2)x(TW+2) halo box is
// loaded per 64-channel chunk and the nine taps are nine DESCRIPTORS into it: with TW = 8 every tile row
// is one 8-row swizzle atom, atoms are (TW+2)*128 = 1280 B apart (SBO), and a tap shift moves the start
// address by (dy*10+dx)*128 B.  Such starts are not 1024-B aligned; measured on B200: the MMA unit applies the
// 128-B swizzle XOR to ABSOLUTE shared-memory address bits (like TMA when it wrote the 1024-B aligned box), so
// the descriptor's base_offset field must stay 0 — setting it to (addr >> 7) & 7 reads garbage
// (profiles/r01_halo_descriptor_experiment.txt).
// MH = 1 or 2 stacked 128-pixel halves per CTA share every weight tile (halves the weight stream).
// Rings: A (halo, one slot per chunk) and B (one slot per (tap, chunk)) are pipelined independently.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_sdesc_sw128_halo(uint32_t smem_addr, uint32_t sbo_bytes, bool with_base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  if (with_base_offset) d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;  // swizzle phase of the start row
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int ACT, bool HAS_RES>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int MH = p.mh;
  constexpr int HW = 10;  // halo row width in pixels (TW = 8)

  const int groups = (p.BN + 63) >> 6;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + p.stages_a * p.a_stage_bytes;
  const uint32_t sStage = sB + p.stages * p.b_stage_bytes;
  const uint32_t sBar = sStage + groups * kAStageBytes;
  // barriers: fullA[4] emptyA[4] fullB[8] emptyB[8] tfull[2] tempty[2] res, tmem slot
  const uint32_t bar_fa = sBar, bar_ea = sBar + 32, bar_fb = sBar + 64, bar_eb = sBar + 128;
  const uint32_t bar_tfull = sBar + 192, bar_tempty = sBar + 208, bar_res = sBar + 224, tmem_slot = sBar + 232;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages_a; ++s) { mbar_init(bar_fa + 8 * s, 1); mbar_init(bar_ea + 8 * s, 1); }
    for (int s = 0; s < p.stages; ++s) { mbar_init(bar_fb + 8 * s, 1); mbar_init(bar_eb + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
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
  const int TH = 16 * MH;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int y0 = (r / p.tiles_w) * TH, x0 = (r % p.tiles_w) * 8;
        const int n0 = nt * p.BN;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(bar_ea + 8 * sa, pha ^ 1);
          mbar_expect_tx(bar_fa + 8 * sa, p.a_box_bytes);
          tma_load_4d(sA + sa * p.a_stage_bytes, &p.tmA[0], bar_fa + 8 * sa, kc * 64, x0 - 1, y0 - 1, img);
          if (++sa == (uint32_t)p.stages_a) { sa = 0; pha ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(bar_eb + 8 * sb, phb ^ 1);
            mbar_expect_tx(bar_fb + 8 * sb, p.b_stage_bytes);
            tma_load_3d(sB + sb * p.b_stage_bytes, &p.tmW, bar_fb + 8 * sb, kc * 64, tap, n0);
            if (++sb == (uint32_t)p.stages) { sb = 0; phb ^= 1; }
          }
        }
        YX_TRACE(0, (tile - (int)blockIdx.x) / (int)gridDim.x);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer (all per-iteration state incremental, 32-bit) ----------------
      uint32_t t = 0, sa = 0, pha = 0, sb = 0, phb = 0;
      const uint32_t a_hi = sdesc_hi(HW * 128), b_hi = sdesc_hi(1024);
      const uint32_t b_lo0 = sdesc_lo(sB), b_step = p.b_stage_bytes >> 4;
      uint32_t b_lo = b_lo0;
      const int ks_last = (p.cin - (p.k_chunks - 1) * 64) >> 4;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const int nt = tile % p.n_tiles_n;
        const int n0 = nt * p.BN;
        const int bn_cur = min(p.BN, p.cout16 - n0);
        const uint32_t idesc = make_idesc_f16(bn_cur);
        const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
        mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        YX_TRACE(1, t);
        const uint32_t d0 = tmem_base + (acc * MH) * p.acc_stride, d1 = d0 + p.acc_stride;
        uint32_t accum = 0;
        for (int kc = 0; kc < p.k_chunks; ++kc) {
          mbar_wait(bar_fa + 8 * sa, pha);
          const int ksteps = (kc == p.k_chunks - 1) ? ks_last : 4;
          // tap (dy,dx) of half h starts (16h + dy) halo rows down and dx pixels right: rows are 128 B = 8 units
          uint32_t a_tap = sdesc_lo(sA + sa * p.a_stage_bytes);
          for (int dy = 0; dy < 3; ++dy, a_tap += (HW - 3) * 8) {
            for (int dx = 0; dx < 3; ++dx, a_tap += 8) {
              mbar_wait(bar_fb + 8 * sb, phb);
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < ksteps) {
                  umma_f16_ss_lohi(d0, a_tap + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, accum | (ks > 0));
                  if (MH == 2)
                    umma_f16_ss_lohi(d1, a_tap + 16 * HW * 8 + 2 * ks, a_hi, b_lo + 2 * ks, b_hi, idesc, accum | (ks > 0));
                }
              accum = 1;
              umma_commit(bar_eb + 8 * sb);
              b_lo += b_step;
              if (++sb == (uint32_t)p.stages) { sb = 0; phb ^= 1; b_lo = b_lo0; }
            }
          }
          umma_commit(bar_ea + 8 * sa);
          if (++sa == (uint32_t)p.stages_a) { sa = 0; pha ^= 1; }
        }
        umma_commit(bar_tfull + 8 * acc);
        YX_TRACE(2, t);
      }
    }
  } else {
    // ---------------- epilogue (warps 2..5) ----------------
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    uint32_t t = 0, res_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int nt = tile % p.n_tiles_n, mt = tile / p.n_tiles_n;
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int y0 = (r / p.tiles_w) * TH, x0 = (r % p.tiles_w) * 8;
      const int n0 = nt * p.BN;
      const int bn_cur = min(p.BN, p.cout16 - n0);
      const int groups_cur = (bn_cur + 63) >> 6;
      const uint32_t acc = t & 1, acc_ph = (t >> 1) & 1;
      for (int h = 0; h < MH; ++h) {
        const int yh = y0 + 16 * h;
        if (HAS_RES && leader) {
          mbar_expect_tx(bar_res, groups_cur * kAStageBytes);
          for (int g = 0; g < groups_cur; ++g)
            tma_load_4d(sStage + g * kAStageBytes, &p.tmRes, bar_res, n0 + g * 64, x0, yh, img);
        }
        if (h == 0) {
          mbar_wait(bar_tfull + 8 * acc, acc_ph);
          tc_fence_after();
          if (leader) YX_TRACE(3, t);
        }
        if (HAS_RES) {
          mbar_wait(bar_res, res_cnt & 1);
          ++res_cnt;
        }
        const uint32_t taddr = tmem_base + (acc * MH + h) * p.acc_stride + (static_cast<uint32_t>(q * 32) << 16);
        epilogue_convert<ACT, HAS_RES>(taddr, bn_cur, row, true, sStage, p.bias + n0);
        if (h == MH - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
        }
        fence_proxy_async_smem();
        if (leader && h == MH - 1) YX_TRACE(4, t);
        named_bar_sync(1, 128);
        if (leader) {
          for (int g = 0; g < groups_cur; ++g)
            tma_store_4d(&p.tmOut, sStage + g * kAStageBytes, n0 + g * 64, x0, yh, img);
          tma_store_commit();
          tma_store_wait_read0();
          if (h == MH - 1) YX_TRACE(6, t);
        }
        named_bar_sync(1, 128);
      }
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

  // ---- halo-reuse shape for 3x3 / stride 1 (see conv3x3_halo_kernel) ------------------------------
  // YX_HALO=0 disables it (2 = diagnostic: set the descriptor base_offset); YX_HALO_MH / YX_HALO_BN force the stacked halves / N tile for experiments.
  static const int halo_env = getenv("YX_HALO") ? atoi(getenv("YX_HALO")) : 1;
  static const int halo_mh_env = getenv("YX_HALO_MH") ? atoi(getenv("YX_HALO_MH")) : 0;
  static const int halo_bn_env = getenv("YX_HALO_BN") ? atoi(getenv("YX_HALO_BN")) : 0;
  p.halo = 0;
  if (halo_env && op.ksize == 3 && op.stride == 1 && !rowpack && Hout >= 16 && Wout >= 8) {
    const double eff16 = (double)(ceil_div(Hout, 16) * 16) * (ceil_div(Wout, 8) * 8) / ((double)Hout * Wout);
    if (eff16 <= 1.25) {
      int bn, mh;
      if (p.cout16 <= 128) { bn = p.cout16; mh = 2; }
      else if (p.cout16 % 128 == 0) { bn = 128; mh = 2; }
      else if (p.cout16 <= 256) { bn = p.cout16; mh = 1; }
      else { bn = round_up(ceil_div(p.cout16, ceil_div(p.cout16, 256)), 64); mh = 1; }
      if (halo_bn_env > 0 && halo_bn_env % 16 == 0 && halo_bn_env <= 256 && (halo_bn_env % 64 == 0 || halo_bn_env >= p.cout16))
        bn = std::min(halo_bn_env, p.cout16);
      if (halo_mh_env == 1 || halo_mh_env == 2) mh = halo_mh_env;
      int stride_cols = 32;
      while (stride_cols < bn) stride_cols <<= 1;
      if (2 * mh * stride_cols > 512) mh = 1;
      const double eff32 = (double)(ceil_div(Hout, 32) * 32) / (ceil_div(Hout, 16) * 16);
      if (mh == 2 && eff32 > 1.2) mh = 1;
      p.halo = halo_env == 2 ? 2 : 1; p.mh = mh; p.BN = bn;
      p.n_tiles_n = ceil_div(p.cout16, bn);
      p.TH = 16 * mh; p.TW = 8;
      p.tiles_h = ceil_div(Hout, p.TH);
      p.tiles_w = ceil_div(Wout, 8);
      p.n_tiles_m = d.n * p.tiles_h * p.tiles_w;
      p.acc_stride = stride_cols;
      p.tmem_cols = 2 * mh * stride_cols;
      p.b_stage_bytes = bn * 128;
      p.a_box_bytes = (p.TH + 2) * 10 * 128;
      p.a_stage_bytes = round_up(p.a_box_bytes, 1024);
      p.stages_a = 2;
      const int hfixed = ceil_div(bn, 64) * kAStageBytes + kBarBytes + 1024 + p.stages_a * p.a_stage_bytes;
      p.stages = std::min(kMaxStages, (kSmemLimit - hfixed) / p.b_stage_bytes);
      YX_REQUIRE(p.stages >= 3, "halo conv: not enough shared memory for the weight ring");
      pl.smem_bytes = hfixed + p.stages * p.b_stage_bytes;
      pl.smem_bytes = std::max(pl.smem_bytes, 120 * 1024);  // one CTA per SM (it may own all 512 TMEM columns)
      pl.grid = std::min(p.n_tiles_m * p.n_tiles_n, num_sms);
    }
  }

  int rc;
  if (p.halo) {
    uint64_t dims[4] = {(uint64_t)s.c, (uint64_t)s.w, (uint64_t)s.h, (uint64_t)s.n};
    uint64_t st[4] = {2, (uint64_t)s.pitch * 2, (uint64_t)s.pitch * 2 * s.w, (uint64_t)s.nstride * 2};
    uint32_t box[4] = {64, 10, (uint32_t)(p.TH + 2), 1};
    if ((rc = encode_map(&p.tmA[0], static_cast<uint8_t*>(base) + s.offset, 4, dims, st, box, true, "A-halo")) != YX_OK)
      return rc;
    for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
  } else if (rowpack) {
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
  const int store_th = p.halo ? 16 : p.TH;  // the halo kernel stores one 16x8 half at a time
  if ((rc = encode_view(&p.tmOut, base, d, p.TW, store_th, "out")) != YX_OK) return rc;
  if (has_res) {
    if ((rc = encode_view(&p.tmRes, base, op.res, p.TW, store_th, "res")) != YX_OK) return rc;
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
  if (plan.p.halo) {
    static bool halo_attr_set = false;
    if (!halo_attr_set) {
      YX_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<ACT, HAS_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
      halo_attr_set = true;
    }
    conv3x3_halo_kernel<ACT, HAS_RES><<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.p);
  } else {
    conv_igemm_kernel<ACT, HAS_RES><<<plan.grid, kThreads, plan.smem_bytes, stream>>>(plan.p);
  }
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
