// The rows either side of the hot path, as HBM-coalesced byte/float kernels:
//   yx_preprocess_batch  decoded RGB uint8 images -> aspect-preserving Pillow-BILINEAR resize -> top-left paste into a
//                        114-filled [B,3,Hp,Wp] batch, RGB -> BGR, NCHW, fp16/fp32 (values 0..255, no normalisation)
//                        replaces: choijhanyangackr/yolox_infer/preprocess_utils.py:9-55 (PIL resize + numpy collate)
//   yx_coco_records      det[B,max_det,7] -> [x, y, w, h, score, category_id] records
//                        replaces: choijhanyangackr/common/utils.py:27-73 (convert_to_coco_format's per-box arithmetic)
// The resize is Pillow's ImagingResample restated (src/libImaging/Resample.c): per-axis 22-bit fixed-point coefficient
// tables (built on the host in double precision, exactly like precompute_coeffs / normalize_coeffs_8bpc), horizontal
// pass rounded+clipped to uint8, then vertical pass — evaluated per output pixel, bit-identical to Pillow.
#include <cuda_fp16.h>

#include "yx_internal.h"

namespace yx {

constexpr int kPrecisionBits = 22;

template <typename T> __device__ __forceinline__ T from_pixel(int v) { return static_cast<T>(static_cast<float>(v)); }
template <> __device__ __forceinline__ uint8_t from_pixel<uint8_t>(int v) { return static_cast<uint8_t>(v); }

template <typename T>
__global__ void preprocess_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ src_off,
                                  const int32_t* __restrict__ geom,      // [B][4] = h, w, new_h, new_w
                                  const int32_t* __restrict__ bounds_h,  // [B][Wp][2] first input column, taps
                                  const int32_t* __restrict__ kk_h,      // [B][Wp][ks_h]
                                  const int32_t* __restrict__ bounds_v,  // [B][Hp][2]
                                  const int32_t* __restrict__ kk_v,      // [B][Hp][ks_v]
                                  int ks_h, int ks_v, int B, int Hp, int Wp, T* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;  // x fastest: coalesced writes of the three planes
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= Wp) return;
  const int h = geom[b * 4 + 0], w = geom[b * 4 + 1], nh = geom[b * 4 + 2], nw = geom[b * 4 + 3];
  (void)h;
  int r = 114, g = 114, bl = 114;  // pad value of yolox_collate_batch
  if (y < nh && x < nw) {
    const uint8_t* img = src + src_off[b];
    const int x0 = bounds_h[((int64_t)b * Wp + x) * 2], nx = bounds_h[((int64_t)b * Wp + x) * 2 + 1];
    const int y0 = bounds_v[((int64_t)b * Hp + y) * 2], ny = bounds_v[((int64_t)b * Hp + y) * 2 + 1];
    const int32_t* kh = kk_h + ((int64_t)b * Wp + x) * ks_h;
    const int32_t* kv = kk_v + ((int64_t)b * Hp + y) * ks_v;
    const int half = 1 << (kPrecisionBits - 1);
    int v0 = half, v1 = half, v2 = half;
    for (int j = 0; j < ny; ++j) {
      const uint8_t* row = img + ((int64_t)(y0 + j) * w + x0) * 3;
      int s0 = half, s1 = half, s2 = half;
      for (int i = 0; i < nx; ++i) {
        const int k = __ldg(kh + i);
        s0 += row[3 * i + 0] * k;
        s1 += row[3 * i + 1] * k;
        s2 += row[3 * i + 2] * k;
      }
      const int kvj = __ldg(kv + j);  // horizontal pass result is a uint8 (clip8) before the vertical pass
      v0 += min(max(s0 >> kPrecisionBits, 0), 255) * kvj;
      v1 += min(max(s1 >> kPrecisionBits, 0), 255) * kvj;
      v2 += min(max(s2 >> kPrecisionBits, 0), 255) * kvj;
    }
    r = min(max(v0 >> kPrecisionBits, 0), 255);
    g = min(max(v1 >> kPrecisionBits, 0), 255);
    bl = min(max(v2 >> kPrecisionBits, 0), 255);
  }
  const int64_t plane = (int64_t)Hp * Wp;
  T* o = out + (int64_t)b * 3 * plane + (int64_t)y * Wp + x;
  o[0] = from_pixel<T>(bl);  // BGR
  o[plane] = from_pixel<T>(g);
  o[2 * plane] = from_pixel<T>(r);
}

__global__ void coco_records_kernel(const float* __restrict__ det, const int32_t* __restrict__ count, int B, int max_det,
                                    const float* __restrict__ scale, const int32_t* __restrict__ class_ids, int n_classes,
                                    float* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * max_det) return;
  const int b = i / max_det, j = i % max_det;
  float o[6] = {0, 0, 0, 0, 0, 0};
  if (j < count[b]) {
    const float* d = det + (int64_t)i * 7;
    const float s = scale[b];
    // boxes /= scale (true fp32 division, torch CPU semantics), then xyxy2xywh in place (utils.py:13-16,55-58)
    const float x1 = __fdiv_rn(d[0], s), y1 = __fdiv_rn(d[1], s), x2 = __fdiv_rn(d[2], s), y2 = __fdiv_rn(d[3], s);
    o[0] = x1; o[1] = y1; o[2] = __fsub_rn(x2, x1); o[3] = __fsub_rn(y2, y1);
    o[4] = __fmul_rn(d[4], d[5]);  // scores = output[:, 4] * output[:, 5]
    const int label = static_cast<int>(d[6]);
    o[5] = static_cast<float>(label >= 0 && label < n_classes ? class_ids[label] : 0);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) rec[(int64_t)i * 6 + k] = o[k];
}

}  // namespace yx

using namespace yx;

extern "C" int yx_preprocess_batch(const void* src, const int64_t* src_off, const int32_t* geom, const int32_t* bounds_h,
                                   const int32_t* kk_h, const int32_t* bounds_v, const int32_t* kk_v, int ks_h, int ks_v, int B,
                                   int Hp, int Wp, void* out, int out_dtype, void* stream) {
  YX_REQUIRE(src && src_off && geom && bounds_h && kk_h && bounds_v && kk_v && out, "null argument");
  YX_REQUIRE(B > 0 && Hp > 0 && Wp > 0 && ks_h > 0 && ks_v > 0 && Hp <= 65535 && B <= 65535, "bad batch geometry");
  YX_REQUIRE(out_dtype == YX_F16 || out_dtype == YX_F32 || out_dtype == YX_U8, "output dtype must be fp16, fp32 or uint8");
  const dim3 block(128), grid(ceil_div(Wp, 128), Hp, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == YX_F16)
    preprocess_kernel<__half><<<grid, block, 0, st>>>(static_cast<const uint8_t*>(src), src_off, geom, bounds_h, kk_h, bounds_v,
                                                       kk_v, ks_h, ks_v, B, Hp, Wp, static_cast<__half*>(out));
  else if (out_dtype == YX_U8)
    preprocess_kernel<uint8_t><<<grid, block, 0, st>>>(static_cast<const uint8_t*>(src), src_off, geom, bounds_h, kk_h, bounds_v,
                                                        kk_v, ks_h, ks_v, B, Hp, Wp, static_cast<uint8_t*>(out));
  else
    preprocess_kernel<float><<<grid, block, 0, st>>>(static_cast<const uint8_t*>(src), src_off, geom, bounds_h, kk_h, bounds_v,
                                                     kk_v, ks_h, ks_v, B, Hp, Wp, static_cast<float*>(out));
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

extern "C" int yx_coco_records(const float* det, const int32_t* det_count, int B, int max_det, const float* scale,
                               const int32_t* class_ids, int n_classes, float* records, void* stream) {
  YX_REQUIRE(det && det_count && scale && class_ids && records && B > 0 && max_det > 0 && n_classes > 0, "bad argument");
  const int n = B * max_det;
  coco_records_kernel<<<ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(det, det_count, B, max_det, scale,
                                                                                      class_ids, n_classes, records);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}
