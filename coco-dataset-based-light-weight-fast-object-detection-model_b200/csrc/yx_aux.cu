// HBM-bound helper kernels of the forward pass (vectorised, NHWC fp16, 16-byte accesses):
//   K3  space-to-depth of the NCHW input image (Focus / FocusCustom)  network_blocks.py:330-361
//   K3b SPP max-pools 5/9/13 written straight into the concat buffer   network_blocks.py:239-246
//   nearest x2 upsample into a concat slice                            yolo_pafpn_p6.py:153-164
//   K2  depthwise kxk conv + bias + act                                network_blocks.py:107-120
#include <cuda_fp16.h>

#include <algorithm>

#include <cstdlib>

#include "yx_internal.h"

namespace yx {

__device__ __forceinline__ float act_f(float x, int act) {
  switch (act) {
    case YX_ACT_SILU: return x / (1.0f + __expf(-x));
    case YX_ACT_HSWISH: return x * fminf(fmaxf(x + 3.0f, 0.0f), 6.0f) * (1.0f / 6.0f);
    case YX_ACT_RELU: return fmaxf(x, 0.0f);
    case YX_ACT_LRELU: return x > 0.0f ? x : 0.1f * x;
    default: return x;
  }
}

// ------------------------------------------------------------------------------------ S2D
// One thread per output pixel: reads a 2x2 patch of each of the 3 planes (coalesced along x),
// optionally applies the predict loop's input affine (main.py:164, img.mul_(s).add_(b) in the image
// dtype), writes 16 fp16 channels (12 real + 4 zero so the stem conv's K is a multiple of 16).
template <typename T>
__device__ __forceinline__ float affine_in(T v, float scale, float shift, bool on);
template <>
__device__ __forceinline__ float affine_in<__half>(__half v, float scale, float shift, bool on) {
  float x = __half2float(v);
  if (on) {
    x = __half2float(__float2half_rn(x * scale));   // half.mul_(s): fp32 opmath, rounded to half
    x = __half2float(__float2half_rn(x + shift));   // half.add_(b)
  }
  return x;
}
template <>
__device__ __forceinline__ float affine_in<uint8_t>(uint8_t v, float scale, float shift, bool on) {
  return affine_in<__half>(__ushort2half_rn(v), scale, shift, on);   // 0..255 is exact in fp16: same bits as a half image
}
template <>
__device__ __forceinline__ float affine_in<float>(float v, float scale, float shift, bool on) {
  return on ? (v * scale) + shift : v;
}

template <typename T>
__global__ void s2d_kernel(const T* __restrict__ img, __half* __restrict__ out, int B, int H, int W, int pitch,
                           int64_t dn, int order, int padded, float scale, float shift, int affine) {
  const int Ho = H >> 1, Wo = W >> 1;
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = i % Wo, y = (i / Wo) % Ho, b = i / ((int64_t)Wo * Ho);
    __align__(16) __half v[16];
#pragma unroll
    for (int k = 12; k < 16; ++k) v[k] = __float2half(0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const T* p = img + (((int64_t)b * 3 + c) * H + 2 * y) * W + 2 * x;
      const float tl = affine_in<T>(p[0], scale, shift, affine), tr = affine_in<T>(p[1], scale, shift, affine);
      const float bl = affine_in<T>(p[W], scale, shift, affine), br = affine_in<T>(p[W + 1], scale, shift, affine);
      if (order == 1) {  // pixel_unshuffle: oc = c*4 + dy*2 + dx
        v[c * 4 + 0] = __float2half_rn(tl); v[c * 4 + 1] = __float2half_rn(tr);
        v[c * 4 + 2] = __float2half_rn(bl); v[c * 4 + 3] = __float2half_rn(br);
      } else {           // Focus: [TL, BL, TR, BR] patch-major
        v[0 + c] = __float2half_rn(tl); v[3 + c] = __float2half_rn(bl);
        v[6 + c] = __float2half_rn(tr); v[9 + c] = __float2half_rn(br);
      }
    }
    // padded layout (row-packed stem conv): row = [0 | pixels 0..Wo-1 | 0 0 0]
    const int Wrow = padded ? Wo + 4 : Wo;
    uint4* o = reinterpret_cast<uint4*>(out + b * dn + ((int64_t)y * Wrow + x + (padded ? 1 : 0)) * pitch);
    o[0] = *reinterpret_cast<const uint4*>(&v[0]);
    o[1] = *reinterpret_cast<const uint4*>(&v[8]);
    if (padded) {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      if (x == 0) { o[-2] = z; o[-1] = z; }
      if (x == Wo - 1) {
#pragma unroll
        for (int j = 2; j < 8; ++j) o[j] = z;
      }
    }
  }
}

int s2d_launch(const void* image, int image_dtype, int aux, int B, int H, int W, float scale, float shift,
               void* base, const yx_view& dst, cudaStream_t stream) {
  const int order = aux & 1, padded = (aux >> 1) & 1;
  YX_REQUIRE(H % 2 == 0 && W % 2 == 0, "image H and W must be even");
  YX_REQUIRE(dst.n == B && dst.h == H / 2 && dst.w == W / 2 + 4 * padded && dst.c == 16 && dst.pitch == 16 &&
                 dst.offset % 16 == 0,
             "s2d dst must be [B,H/2,W/2(+4 when padded),16]");
  __half* out = reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + dst.offset);
  const int64_t total = (int64_t)B * (H / 2) * (W / 2);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  const int affine = (scale != 1.0f || shift != 0.0f) ? 1 : 0;
  if (image_dtype == YX_F16)
    s2d_kernel<__half><<<blocks, threads, 0, stream>>>(static_cast<const __half*>(image), out, B, H, W, dst.pitch,
                                                       dst.nstride, order, padded, scale, shift, affine);
  else if (image_dtype == YX_F32)
    s2d_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(image), out, B, H, W, dst.pitch,
                                                      dst.nstride, order, padded, scale, shift, affine);
  else if (image_dtype == YX_U8)
    s2d_kernel<uint8_t><<<blocks, threads, 0, stream>>>(static_cast<const uint8_t*>(image), out, B, H, W, dst.pitch,
                                                        dst.nstride, order, padded, scale, shift, affine);
  else
    YX_REQUIRE(false, "image dtype must be YX_F16, YX_F32 or YX_U8");
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------ SPP
__device__ __forceinline__ uint4 hmax8(uint4 a, uint4 b) {
  uint4 r;
  __half2* ra = reinterpret_cast<__half2*>(&a);
  __half2* rb = reinterpret_cast<__half2*>(&b);
  __half2* rr = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) rr[i] = __hmax2(ra[i], rb[i]);
  return r;
}

// src: [B,h,w,C] (slice 0 of the concat buffer); dst: [B,h,w,3C] = slices 1..3 (pool 5, 9, 13).
// One thread per (pixel, 8 channels); out-of-image taps are skipped (= -inf padding of nn.MaxPool2d).
__global__ void spp_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int B, int H, int W, int C,
                           int spitch, int dpitch, int64_t sn, int64_t dn) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * H * W * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int64_t pix = i / cv;
    const int x = pix % W, y = (pix / W) % H, b = pix / ((int64_t)W * H);
    const uint4 ninf = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
    uint4 m5 = ninf, m9 = ninf, m13 = ninf;
    for (int dy = -6; dy <= 6; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      const int ady = dy < 0 ? -dy : dy;
      for (int dx = -6; dx <= 6; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        const int adx = dx < 0 ? -dx : dx;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + b * sn + ((int64_t)yy * W + xx) * spitch) + c8);
        m13 = hmax8(m13, v);
        if (ady <= 4 && adx <= 4) m9 = hmax8(m9, v);
        if (ady <= 2 && adx <= 2) m5 = hmax8(m5, v);
      }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + b * dn + ((int64_t)y * W + x) * dpitch);
    o[c8] = m5;
    o[cv + c8] = m9;
    o[2 * cv + c8] = m13;
  }
}

// Tiled version for maps that fit in shared memory (every P5 / P6 configuration: 20x20 .. 40x40): one CTA per
// (image, group of CV 8-channel vectors).  Uses the exact identities pool9 = pool5 o pool5, pool13 = pool5 o pool5 o pool5
// (max is associative and the -inf padding is its neutral element) and the separability of the 5x5 window: 30 shared
// memory reads per output vector instead of 169 global ones.
__global__ void __launch_bounds__(256) spp_tiled_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int H, int W,
                                                        int C, int CV, int spitch, int dpitch, int64_t sn, int64_t dn) {
  extern __shared__ uint4 spp_smem[];
  const int cv = C >> 3, groups = (cv + CV - 1) / CV;
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int c0 = g * CV, ncv = min(CV, cv - c0);
  const int n = H * W * ncv;
  uint4* bufA = spp_smem;
  uint4* bufB = spp_smem + H * W * CV;
  const uint4 ninf = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = i % ncv, pix = i / ncv;
    bufA[i] = __ldg(reinterpret_cast<const uint4*>(src + b * sn + (int64_t)pix * spitch) + c0 + v);
  }
  __syncthreads();
  for (int level = 0; level < 3; ++level) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {  // horizontal 5-max: A -> B
      const int v = i % ncv, pix = i / ncv, x = pix % W;
      uint4 m = bufA[i];
#pragma unroll
      for (int d = 1; d <= 2; ++d) {
        if (x - d >= 0) m = hmax8(m, bufA[i - d * ncv]);
        if (x + d < W) m = hmax8(m, bufA[i + d * ncv]);
      }
      (void)v;
      bufB[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {  // vertical 5-max: B -> A, and out to slice `level`
      const int v = i % ncv, pix = i / ncv, y = pix / W;
      uint4 m = bufB[i];
#pragma unroll
      for (int d = 1; d <= 2; ++d) {
        if (y - d >= 0) m = hmax8(m, bufB[i - d * W * ncv]);
        if (y + d < H) m = hmax8(m, bufB[i + d * W * ncv]);
      }
      bufA[i] = m;
      reinterpret_cast<uint4*>(dst + b * dn + (int64_t)pix * dpitch)[level * cv + c0 + v] = m;
    }
    __syncthreads();
  }
  (void)ninf;
}

int spp_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream) {
  YX_REQUIRE(src.c % 8 == 0 && dst.c == 3 * src.c && dst.n == src.n && dst.h == src.h && dst.w == src.w,
             "spp dst must be [B,h,w,3C]");
  YX_REQUIRE(src.offset % 16 == 0 && dst.offset % 16 == 0 && src.pitch % 8 == 0 && dst.pitch % 8 == 0, "spp alignment");
  const __half* sp = reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + src.offset);
  __half* dp = reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + dst.offset);
  static const bool direct_only = getenv("YX_SPP_DIRECT") != nullptr;
  const int cv = src.c / 8;
  const int64_t pix_bytes = (int64_t)src.h * src.w * 16 * 2;  // two buffers, per 8-channel vector
  int CV = (int)std::min<int64_t>(4, (96 * 1024) / std::max<int64_t>(pix_bytes, 1));
  CV = std::min(CV, cv);
  // small batches (bs1 latency, 8 images per GPU): fewer vectors per CTA until the grid covers the SMs -- a CTA's time is
  // seven block barriers and 6 x ceil(H*W*CV/256) passes, so at bs1 twelve CTAs of CV = 4 took 19 us for 0.6 MB
  while (CV > 1 && (int64_t)src.n * ((cv + CV - 1) / CV) < 148) --CV;
  if (CV >= 1 && !direct_only) {
    static bool attr = false;
    if (!attr) {
      YX_CUDA(cudaFuncSetAttribute(spp_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr = true;
    }
    const int groups = (cv + CV - 1) / CV;
    spp_tiled_kernel<<<src.n * groups, 256, (size_t)(pix_bytes * CV), stream>>>(sp, dp, src.h, src.w, src.c, CV, src.pitch,
                                                                             dst.pitch, src.nstride, dst.nstride);
    YX_CUDA(cudaGetLastError());
    return YX_OK;
  }
  const int64_t total = (int64_t)src.n * src.h * src.w * (src.c / 8);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  spp_kernel<<<blocks, threads, 0, stream>>>(sp, dp, src.n, src.h, src.w, src.c, src.pitch, dst.pitch, src.nstride, dst.nstride);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------ upsample
__global__ void upsample2_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int B, int H, int W, int C,
                                 int spitch, int dpitch, int64_t sn, int64_t dn) {
  // H, W are the OUTPUT dims
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * H * W * cv;
  const int Hs = H >> 1, Ws = W >> 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int64_t pix = i / cv;
    const int x = pix % W, y = (pix / W) % H, b = pix / ((int64_t)W * H);
    const uint4 v =
        __ldg(reinterpret_cast<const uint4*>(src + b * sn + ((int64_t)(y >> 1) * Ws + (x >> 1)) * spitch) + c8);
    reinterpret_cast<uint4*>(dst + b * dn + ((int64_t)y * W + x) * dpitch)[c8] = v;
  }
}

int upsample_launch(void* base, const yx_view& src, const yx_view& dst, cudaStream_t stream) {
  YX_REQUIRE(dst.h == 2 * src.h && dst.w == 2 * src.w && dst.c == src.c && dst.n == src.n && src.c % 8 == 0,
             "upsample dst must be [B,2h,2w,C]");
  YX_REQUIRE(src.offset % 16 == 0 && dst.offset % 16 == 0 && src.pitch % 8 == 0 && dst.pitch % 8 == 0, "upsample alignment");
  const int64_t total = (int64_t)dst.n * dst.h * dst.w * (dst.c / 8);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  upsample2_kernel<<<blocks, threads, 0, stream>>>(
      reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + src.offset),
      reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + dst.offset), dst.n, dst.h, dst.w, dst.c, src.pitch,
      dst.pitch, src.nstride, dst.nstride);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------ depthwise
// weights fp16 [k*k][C], bias fp32 [C]; fp32 accumulate; one thread per (output pixel, 8 channels).
__global__ void dwconv_kernel(const __half* __restrict__ src, __half* __restrict__ dst, const __half* __restrict__ wgt,
                              const float* __restrict__ bias, int B, int H, int W, int Ho, int Wo, int C, int k,
                              int stride, int act, int spitch, int dpitch, int64_t sn, int64_t dn) {
  const int cv = C >> 3;
  const int pad = k >> 1;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int64_t pix = i / cv;
    const int x = pix % Wo, y = (pix / Wo) % Ho, b = pix / ((int64_t)Wo * Ho);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int dy = 0; dy < k; ++dy) {
      const int yy = y * stride + dy - pad;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int xx = x * stride + dx - pad;
        if (xx < 0 || xx >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + b * sn + ((int64_t)yy * W + xx) * spitch) + c8);
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(wgt + (int64_t)(dy * k + dx) * C) + c8);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
        const __half2* wh = reinterpret_cast<const __half2*>(&w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = __half22float2(vh[j]), ww = __half22float2(wh[j]);
          acc[2 * j] = fmaf(a.x, ww.x, acc[2 * j]);
          acc[2 * j + 1] = fmaf(a.y, ww.y, acc[2 * j + 1]);
        }
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f0 = acc[2 * j] + bias[c8 * 8 + 2 * j], f1 = acc[2 * j + 1] + bias[c8 * 8 + 2 * j + 1];
      f0 = act_f(f0, act);   // activation on the fp32 sum, ONE rounding to fp16: the conv kernel's rule
      f1 = act_f(f1, act);
      oh[j] = __floats2half2_rn(f0, f1);
    }
    reinterpret_cast<uint4*>(dst + b * dn + ((int64_t)y * Wo + x) * dpitch)[c8] = o;
  }
}

// Register-blocked form for stride 1: one thread = 8 channels x a strip of TX consecutive output pixels.  Per filter row the
// thread loads the TX + K - 1 input vectors of the strip ONCE (the k*k-loads-per-output form above re-reads every input
// k*k times through L1) and the K weight vectors of that row; consecutive threads are consecutive channel vectors of the
// same strip, so every load / store instruction of a warp is one contiguous run of up to 512 bytes.
template <int K, int TX>
__global__ void __launch_bounds__(256) dwconv_strip_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                           const __half* __restrict__ wgt, const float* __restrict__ bias, int B,
                                                           int H, int W, int C, int act, int spitch, int dpitch, int64_t sn,
                                                           int64_t dn) {
  constexpr int PAD = K / 2;
  const int cv = C >> 3, strips = (W + TX - 1) / TX;
  const int64_t total = (int64_t)B * H * strips * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = i % cv;
    const int64_t r = i / cv;
    const int x0 = (int)(r % strips) * TX, y = (int)((r / strips) % H), b = (int)(r / ((int64_t)strips * H));
    float acc[TX][8];
#pragma unroll
    for (int t = 0; t < TX; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
      const int yy = y + dy - PAD;
      if (yy < 0 || yy >= H) continue;
      float wf[K][8];   // this filter row's weights, converted ONCE (the kernel is fp32-FMA-issue bound, not HBM bound)
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wgt + (int64_t)(dy * K + dx) * C) + c8);
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 ww = __half22float2(wh[j]);
          wf[dx][2 * j] = ww.x; wf[dx][2 * j + 1] = ww.y;
        }
      }
      const __half* row = src + b * sn + (int64_t)yy * W * spitch;
#pragma unroll
      for (int xi = 0; xi < TX + K - 1; ++xi) {
        const int xx = x0 + xi - PAD;
        if (xx < 0 || xx >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + (int64_t)xx * spitch) + c8);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
        float a[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(vh[j]);
          a[2 * j] = f.x; a[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {   // input column xi feeds output t = xi - dx through tap dx
          const int t = xi - dx;
          if (t < 0 || t >= TX) continue;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(a[j], wf[dx][j], acc[t][j]);
        }
      }
    }
    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bv[j] = bias[c8 * 8 + j];
#pragma unroll
    for (int t = 0; t < TX; ++t) {
      if (x0 + t >= W) break;
      uint4 o;
      __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f0 = acc[t][2 * j] + bv[2 * j], f1 = acc[t][2 * j + 1] + bv[2 * j + 1];
        f0 = act_f(f0, act);   // activation on the fp32 sum, ONE rounding to fp16: the conv kernel's rule
        f1 = act_f(f1, act);
        oh[j] = __floats2half2_rn(f0, f1);
      }
      reinterpret_cast<uint4*>(dst + b * dn + ((int64_t)y * W + x0 + t) * dpitch)[c8] = o;
    }
  }
}

// Shared-memory tiled form for stride 1 (round 2): a CTA owns an 8 x 16 pixel tile of 64 channels.  The (8+K-1) x (16+K-1)
// halo tile is loaded ONCE, cooperatively and coalesced (eight 16-byte channel vectors = 128 contiguous bytes per pixel), so
// every input byte leaves L2 once per tile instead of up to K times per output row, and the K x (4+K-1) vector reads of a thread's
// 4-pixel strip come from shared memory: the compute phase has no global-memory latency in it, and with three resident CTAs
// per SM one CTA's loads overlap the others' FMAs (the register-blocked strip kernel above ran at 0.4 IPC, 0.11 of HBM: every
// one of its five row passes waited for its own L2 loads).  Accumulation order per output = the strip kernel's (dy, then dx).
template <int K>
__global__ void __launch_bounds__(256, 3) dwconv_tile_kernel(const __half* __restrict__ src, __half* __restrict__ dst,
                                                          const __half* __restrict__ wgt, const float* __restrict__ bias, int B,
                                                          int H, int W, int C, int act, int spitch, int dpitch, int64_t sn,
                                                          int64_t dn, int tiles_x, int tiles_y, int cgroups) {
  constexpr int PAD = K / 2, TH = 8, TW = 16, TX = 4, HH = TH + K - 1, HW = TW + K - 1;
  __shared__ uint4 s_in[HH * HW * 8];
  const int tid = threadIdx.x;
  const int c8 = tid & 7, strip = (tid >> 3) & 3, row = tid >> 5;
  const int64_t n_tiles = (int64_t)B * tiles_y * tiles_x * cgroups;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int cg = (int)(tile % cgroups);
    int64_t r = tile / cgroups;
    const int tx = (int)(r % tiles_x); r /= tiles_x;
    const int ty = (int)(r % tiles_y);
    const int b = (int)(r / tiles_y);
    const int x0 = tx * TW, y0 = ty * TH, c0 = cg * 64;
    const __half* img = src + b * sn;
    for (int i = tid; i < HH * HW * 8; i += 256) {
      const int v = i & 7, p = i >> 3, hx = p % HW, hy = p / HW;
      const int gy = y0 - PAD + hy, gx = x0 - PAD + hx, c = c0 + v * 8;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && c < C)
        val = __ldg(reinterpret_cast<const uint4*>(img + ((int64_t)gy * W + gx) * spitch + c));
      s_in[i] = val;
    }
    __syncthreads();
    const int c = c0 + c8 * 8;
    if (c < C) {
      float acc[TX][8];
#pragma unroll
      for (int t = 0; t < TX; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
#pragma unroll
      for (int dy = 0; dy < K; ++dy) {
        float wf[K][8];
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wgt + (int64_t)(dy * K + dx) * C + c));
          const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 ww = __half22float2(wh[j]);
            wf[dx][2 * j] = ww.x; wf[dx][2 * j + 1] = ww.y;
          }
        }
        const uint4* line = s_in + ((row + dy) * HW + strip * TX) * 8 + c8;
#pragma unroll
        for (int xi = 0; xi < TX + K - 1; ++xi) {
          const uint4 v = line[xi * 8];
          const __half2* vh = reinterpret_cast<const __half2*>(&v);
          float a[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(vh[j]);
            a[2 * j] = f.x; a[2 * j + 1] = f.y;
          }
#pragma unroll
          for (int dx = 0; dx < K; ++dx) {
            const int t = xi - dx;
            if (t < 0 || t >= TX) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(a[j], wf[dx][j], acc[t][j]);
          }
        }
      }
      const int y = y0 + row;
      if (y < H) {
        float bv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bv[j] = bias[c + j];
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          const int x = x0 + strip * TX + t;
          if (x >= W) break;
          uint4 o;
          __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f0 = acc[t][2 * j] + bv[2 * j], f1 = acc[t][2 * j + 1] + bv[2 * j + 1];
            f0 = act_f(f0, act);   // activation on the fp32 sum, ONE rounding to fp16: the conv kernel's rule
            f1 = act_f(f1, act);
            oh[j] = __floats2half2_rn(f0, f1);
          }
          *reinterpret_cast<uint4*>(dst + b * dn + ((int64_t)y * W + x) * dpitch + c) = o;
        }
      }
    }
    __syncthreads();   // the next tile's loads overwrite s_in
  }
}

int dwconv_launch(void* base, const yx_op& op, const void* weights, const void* biases, cudaStream_t stream) {
  const yx_view& s = op.src;
  const yx_view& d = op.dst;
  YX_REQUIRE(op.ksize == 3 || op.ksize == 5, "depthwise ksize must be 3 or 5");
  YX_REQUIRE(op.stride == 1 || op.stride == 2, "depthwise stride must be 1 or 2");
  const int pad = op.ksize / 2;
  const int Ho = (s.h + 2 * pad - op.ksize) / op.stride + 1, Wo = (s.w + 2 * pad - op.ksize) / op.stride + 1;
  YX_REQUIRE(d.h == Ho && d.w == Wo && d.c == s.c && d.n == s.n && s.c % 8 == 0, "depthwise dst geometry");
  YX_REQUIRE(s.offset % 16 == 0 && d.offset % 16 == 0 && s.pitch % 8 == 0 && d.pitch % 8 == 0 && op.w_offset % 16 == 0,
             "depthwise alignment");
  // (read on every call: the tests switch kernels between engines of one process)
  const bool strip_env = !(getenv("YX_DW_STRIP") && atoi(getenv("YX_DW_STRIP")) == 0);
  const bool tile_env = !(getenv("YX_DW_TILE") && atoi(getenv("YX_DW_TILE")) == 0);
  if (op.stride == 1 && tile_env) {
    const int tiles_x = (Wo + 15) / 16, tiles_y = (Ho + 7) / 8, cgroups = (d.c + 63) / 64;
    const int64_t n_tiles = (int64_t)d.n * tiles_x * tiles_y * cgroups;
    const int nb = (int)std::min<int64_t>(n_tiles, 148 * 6);
    const __half* sp = reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + s.offset);
    __half* dp = reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + d.offset);
    const __half* wp = reinterpret_cast<const __half*>(static_cast<const uint8_t*>(weights) + op.w_offset);
    const float* bp = reinterpret_cast<const float*>(static_cast<const uint8_t*>(biases) + op.b_offset);
    if (op.ksize == 3)
      dwconv_tile_kernel<3><<<nb, 256, 0, stream>>>(sp, dp, wp, bp, s.n, s.h, s.w, s.c, op.act, s.pitch, d.pitch, s.nstride, d.nstride, tiles_x, tiles_y, cgroups);
    else
      dwconv_tile_kernel<5><<<nb, 256, 0, stream>>>(sp, dp, wp, bp, s.n, s.h, s.w, s.c, op.act, s.pitch, d.pitch, s.nstride, d.nstride, tiles_x, tiles_y, cgroups);
    YX_CUDA(cudaGetLastError());
    return YX_OK;
  }
  if (op.stride == 1 && strip_env) {
    constexpr int TX = 8;
    const int64_t n_threads = (int64_t)d.n * Ho * ((Wo + TX - 1) / TX) * (d.c / 8);
    const int nb = (int)std::min<int64_t>((n_threads + 255) / 256, 148 * 16);
    const __half* sp = reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + s.offset);
    __half* dp = reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + d.offset);
    const __half* wp = reinterpret_cast<const __half*>(static_cast<const uint8_t*>(weights) + op.w_offset);
    const float* bp = reinterpret_cast<const float*>(static_cast<const uint8_t*>(biases) + op.b_offset);
    if (op.ksize == 3)
      dwconv_strip_kernel<3, TX><<<nb, 256, 0, stream>>>(sp, dp, wp, bp, s.n, s.h, s.w, s.c, op.act, s.pitch, d.pitch, s.nstride, d.nstride);
    else
      dwconv_strip_kernel<5, TX><<<nb, 256, 0, stream>>>(sp, dp, wp, bp, s.n, s.h, s.w, s.c, op.act, s.pitch, d.pitch, s.nstride, d.nstride);
    YX_CUDA(cudaGetLastError());
    return YX_OK;
  }
  const int64_t total = (int64_t)d.n * Ho * Wo * (d.c / 8);
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 16);
  dwconv_kernel<<<blocks, threads, 0, stream>>>(
      reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + s.offset),
      reinterpret_cast<__half*>(static_cast<uint8_t*>(base) + d.offset),
      reinterpret_cast<const __half*>(static_cast<const uint8_t*>(weights) + op.w_offset),
      reinterpret_cast<const float*>(static_cast<const uint8_t*>(biases) + op.b_offset), s.n, s.h, s.w, Ho, Wo, s.c,
      op.ksize, op.stride, op.act, s.pitch, d.pitch, s.nstride, d.nstride);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------ tuner self-check helpers
// copy a strided NHWC view into a contiguous buffer / max |view - contiguous| (as the bits of a non-negative float)
__global__ void view_gather_kernel(const __half* __restrict__ src, __half* __restrict__ out, int H, int W, int C, int pitch,
                                   int64_t nstride, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int64_t pix = i / C;
    const int64_t b = pix / ((int64_t)H * W), r = pix % ((int64_t)H * W);
    out[i] = src[b * nstride + r * pitch + c];
  }
}
// The difference is measured in fp16 rounding STEPS at the magnitude of the larger operand (never below floor_mag):
// two launch shapes of one conv sum the same products in a different fp32 order, so their fp16 outputs may differ by a
// rounding step or two and by nothing else (the bound of tests/test_gpu_model.py::test_every_op_teacher_forced).
__global__ void view_diff_kernel(const __half* __restrict__ src, const __half* __restrict__ ref, int H, int W, int C, int pitch,
                                 int64_t nstride, int64_t total, float floor_mag, unsigned int* __restrict__ out_bits) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = i % C;
    const int64_t pix = i / C;
    const int64_t b = pix / ((int64_t)H * W), r = pix % ((int64_t)H * W);
    const float x = __half2float(src[b * nstride + r * pitch + c]), y = __half2float(ref[i]);
    const float mag = fmaxf(fmaxf(fabsf(x), fabsf(y)), floor_mag);
    int e;
    frexpf(mag, &e);                                  // mag = f * 2^e, f in [0.5, 1): one fp16 step is 2^(e - 11)
    const float d = fabsf(x - y) * exp2f((float)(11 - e));
    m = fmaxf(m, d == d ? d : 65504.f);               // NaN / inf counts as a maximal difference
  }
  atomicMax(out_bits, __float_as_uint(m));
}

int view_gather(void* base, const yx_view& v, void* out, cudaStream_t st) {
  const int64_t total = (int64_t)v.n * v.h * v.w * v.c;
  view_gather_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
      reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + v.offset), static_cast<__half*>(out), v.h, v.w, v.c, v.pitch,
      v.nstride, total);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}
int view_max_diff(void* base, const yx_view& v, const void* ref, float floor_mag, unsigned int* out_bits, cudaStream_t st) {
  const int64_t total = (int64_t)v.n * v.h * v.w * v.c;
  YX_CUDA(cudaMemsetAsync(out_bits, 0, 4, st));
  view_diff_kernel<<<(int)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, st>>>(
      reinterpret_cast<const __half*>(static_cast<uint8_t*>(base) + v.offset), static_cast<const __half*>(ref), v.h, v.w, v.c,
      v.pitch, v.nstride, total, floor_mag, out_bits);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

}  // namespace yx
