// K4/K5/K6 — head decode, score thresholding + candidate compaction, per-image sort and batched
// class-aware greedy NMS for all images of the batch in one launch each.
//
//   decode (infer flavour)    choijhanyangackr/yolox_infer/postprocess_utils.py:27-52
//   decode (yolox flavour)    yolox/models/yolo_head.py:167-168,186-190,210-225
//   candidate selection       postprocess_utils.py:86-103 ; yolox/utils/boxes.py:38-59
//   NMS                       torchvision.ops.nms / batched_nms semantics (CPU kernel arithmetic:
//                             every fp32 operation rounded separately, strict '>' on IoU, stable
//                             descending score order; ties -> lower anchor index first)
//
// Arithmetic that decides KEPT INDEX SETS (IoU, coordinate-trick offsets) is written with
// __fadd_rn/__fmul_rn/__fdiv_rn so nvcc cannot contract it into FMAs: results are bit-identical to
// the CPU restatement in oracle/post_ref.c.
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "yx_internal.h"

namespace yx {

struct LevelsDev {
  int n;
  int h[8], w[8], stride[8], off[9];
};

static int make_levels(const yx_levels* lv, int A, LevelsDev* out) {
  YX_REQUIRE(lv != nullptr && lv->n_levels >= 1 && lv->n_levels <= 8, "levels: 1..8");
  out->n = lv->n_levels;
  int off = 0;
  for (int i = 0; i < lv->n_levels; ++i) {
    out->h[i] = lv->h[i]; out->w[i] = lv->w[i]; out->stride[i] = lv->stride[i]; out->off[i] = off;
    off += lv->h[i] * lv->w[i];
  }
  out->off[lv->n_levels] = off;
  YX_REQUIRE(off == A, "sum of level h*w must equal A");
  return YX_OK;
}

__device__ __forceinline__ void anchor_geom(const LevelsDev& lv, int a, float* gx, float* gy, float* s) {
  int l = 0;
  while (l + 1 < lv.n && a >= lv.off[l + 1]) ++l;
  const int r = a - lv.off[l];
  *gx = (float)(r % lv.w[l]);
  *gy = (float)(r / lv.w[l]);
  *s = (float)lv.stride[l];
}

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

// postprocess_utils.py:37-44: cx=(tx+gx)*s ; half=exp(tw)*(s/2) ; xyxy = c -/+ half   (fp32)
__device__ __forceinline__ float4 decode_box_xyxy(float t0, float t1, float t2, float t3, float gx, float gy, float s) {
  const float cx = __fmul_rn(__fadd_rn(t0, gx), s), cy = __fmul_rn(__fadd_rn(t1, gy), s);
  const float hw = __fmul_rn(expf(t2), s * 0.5f), hh = __fmul_rn(expf(t3), s * 0.5f);
  return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

// map fp32 to a uint32 whose unsigned order equals the float order
__device__ __forceinline__ uint32_t f2sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float score, int anchor) {
  return (static_cast<uint64_t>(f2sortable(score)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - (uint32_t)anchor);
}

// Warp-aggregated counter increment (callable under divergence): lanes that target the same counter elect a leader
// that performs ONE atomicAdd for all of them.  Per-candidate atomics on one address per image serialise in L2
// (~34 000 same-address atomics per image made the selection kernel 25x slower than its HBM time).
__device__ __forceinline__ int warp_agg_inc(int* ctr) {
  const unsigned active = __activemask();
  const unsigned same = __match_any_sync(active, reinterpret_cast<unsigned long long>(ctr));
  const int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(ctr, __popc(same));
  base = __shfl_sync(same, base, leader);
  return base + __popc(same & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------------
struct Workspace {
  float4* box;     // [B][A]
  float* objc;     // [B][A]
  float* col5;     // [B][A]  value written to det column 5
  float* score;    // [B][A]  NMS score
  int* label;      // [B][A]
  uint64_t* keys;  // [B][Apad]
  float4* sbox;    // [B][A]  boxes in score order (the coordinate-trick offset is added when a block is loaded)
  int* slabel;     // [B][A]
  float4* kbox;    // [B][A]  kept boxes (uncapped mode)
  int* klabel;     // [B][A]
  int* kidx;       // [B][A]  kept candidate rank
  int* count;      // [B]
  int Apad;        // keys per image (power of two >= A*K)
  int K;           // candidates per anchor: 1, or C in multi_class mode (candidate id = anchor*K + class)
  int R;           // rows per image of sbox/slabel/kbox/klabel/kidx
};

// inverse of f2sortable on the high word of a key
__device__ __forceinline__ float key_score(uint64_t key) {
  const uint32_t u = (uint32_t)(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u ^ 0x80000000u) : ~u);
}
__device__ __forceinline__ int key_id(uint64_t key) { return (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull)); }

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// K = candidates per anchor, R = rows per image that can reach the NMS (A for K == 1; min(max_nms, A*K) otherwise)
static size_t ws_layout(int B, int A, int K, int64_t R, uint8_t* base, Workspace* ws) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  };
  const size_t n = (size_t)B * A, r = (size_t)B * (size_t)R;
  const int Apad = next_pow2(std::max(A * K, 2));
  Workspace w;
  w.box = (float4*)take(n * 16); w.objc = (float*)take(n * 4); w.col5 = (float*)take(n * 4);
  w.score = (float*)take(n * 4); w.label = (int*)take(n * 4);
  w.keys = (uint64_t*)take((size_t)B * Apad * 8);
  w.sbox = (float4*)take(r * 16); w.slabel = (int*)take(r * 4);
  w.kbox = (float4*)take(r * 16); w.klabel = (int*)take(r * 4); w.kidx = (int*)take(r * 4);
  w.count = (int*)take((size_t)B * 4);
  w.Apad = Apad; w.K = K; w.R = (int)R;
  if (ws) *ws = w;
  return off;
}

// ------------------------------------------------------------------------------------------------
// decode kernels
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void decode_infer_kernel(const T* __restrict__ reg, int64_t reg_sb, int64_t reg_sa, const T* __restrict__ obj,
                                    int64_t obj_sb, int64_t obj_sa, const T* __restrict__ cls, int64_t cls_sb,
                                    int64_t cls_sa, int B, int A, int C, LevelsDev lv, const void* __restrict__ grids,
                                    const void* __restrict__ scales, int geom_dtype, float* __restrict__ boxes,
                                    float* __restrict__ obj_conf, float* __restrict__ cls_conf) {
  const int64_t total = (int64_t)B * A;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    float gx, gy, s;
    if (grids != nullptr) {  // the caller's own grids (1,A,2) / scales (1,A,1) tensors, promoted to fp32 like torch does
      if (geom_dtype == YX_F16) {
        const __half* g = static_cast<const __half*>(grids) + 2 * a;
        gx = __half2float(g[0]); gy = __half2float(g[1]); s = __half2float(static_cast<const __half*>(scales)[a]);
      } else {
        const float* g = static_cast<const float*>(grids) + 2 * a;
        gx = g[0]; gy = g[1]; s = static_cast<const float*>(scales)[a];
      }
    } else {
      anchor_geom(lv, a, &gx, &gy, &s);
    }
    const T* r = reg + b * reg_sb + a * reg_sa;
    const float4 bx = decode_box_xyxy(ldf(r), ldf(r + 1), ldf(r + 2), ldf(r + 3), gx, gy, s);
    reinterpret_cast<float4*>(boxes)[i] = bx;
    const float oc = sigmoid_f(ldf(obj + b * obj_sb + a * obj_sa));
    obj_conf[i] = oc;
    const T* c = cls + b * cls_sb + a * cls_sa;
    float* o = cls_conf + i * C;
    for (int k = 0; k < C; ++k) o[k] = __fmul_rn(sigmoid_f(ldf(c + k)), oc);
  }
}

// Fused decode + class max + threshold + compaction from raw logits: never writes [B,A,C].
template <typename T, bool VEC8>
__global__ void select_infer_kernel(const T* __restrict__ reg, int64_t reg_sb, int64_t reg_sa, const T* __restrict__ obj,
                                    int64_t obj_sb, int64_t obj_sa, const T* __restrict__ cls, int64_t cls_sb,
                                    int64_t cls_sa, int B, int A, int C, LevelsDev lv, float thr, Workspace ws) {
  const int64_t total = (int64_t)B * A;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    const float oc = sigmoid_f(ldf(obj + b * obj_sb + a * obj_sa));
    const T* c = cls + b * cls_sb + a * cls_sa;
    float best = -1.0f;
    int bi = 0;
    if (VEC8) {  // T == __half, 16-byte aligned rows, C % 8 == 0
      // score(x) = sigmoid_f(x) * oc is NON-DECREASING in the logit x over all fp16 inputs (checked exhaustively
      // by tests/test_gpu_post.py::test_score_monotonic_in_logit), so a class whose logit does not exceed the
      // logit of the running best cannot have a strictly greater score: the exact first-max of the 80 products
      // needs a sigmoid only when a new running-max logit appears (~5 of 80 on average).
      const uint4* cv = reinterpret_cast<const uint4*>(c);
      // (Tried in round 2 and measured slower than this loop's 0.21 ms at B = 64, A = 34 000, C = 80 -- kept out: requesting all
      // ten 16-byte chunks of the row before the scan, 0.29 ms (64 more registers); a warp-cooperative version streaming 32
      // rows as one contiguous run with the per-chunk results combined through shared memory, 0.28 ms, and its refinement
      // with prefix maxima that needs only ~8 sigmoids per row, 0.51 ms.  The kernel is bound by the scan's dependent
      // compare / branch chain per thread, not by HBM: DRAM traffic is the algorithmic 383 MB either way.)
      float best_logit = -INFINITY;
      {  // class 0 is always evaluated (also covers a -inf logit, whose score 0 still beats the -1 sentinel)
        const float x0 = ldf(c);
        const float s = __fmul_rn(sigmoid_f(x0), oc);
        if (s > best) { best = s; bi = 0; best_logit = x0; }
      }
      for (int k8 = 0; k8 < (C >> 3); ++k8) {
        const uint4 v = __ldg(cv + k8);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h[j]);
          if (f.x > best_logit || f.x != f.x) {
            const float s0 = __fmul_rn(sigmoid_f(f.x), oc);
            if (s0 > best) { best = s0; bi = k8 * 8 + 2 * j; best_logit = f.x; }
          }
          if (f.y > best_logit || f.y != f.y) {
            const float s1 = __fmul_rn(sigmoid_f(f.y), oc);
            if (s1 > best) { best = s1; bi = k8 * 8 + 2 * j + 1; best_logit = f.y; }
          }
        }
      }
    } else {
      for (int k = 0; k < C; ++k) {
        const float s0 = __fmul_rn(sigmoid_f(ldf(c + k)), oc);
        if (s0 > best) { best = s0; bi = k; }
      }
    }
    if (best >= thr) {  // torch.greater_equal(cls_conf_i, conf_threshold), postprocess_utils.py:87
      float gx, gy, s;
      anchor_geom(lv, a, &gx, &gy, &s);
      const T* r = reg + b * reg_sb + a * reg_sa;
      ws.box[i] = decode_box_xyxy(ldf(r), ldf(r + 1), ldf(r + 2), ldf(r + 3), gx, gy, s);
      ws.objc[i] = oc; ws.col5[i] = best; ws.score[i] = best; ws.label[i] = bi;
      const int pos = warp_agg_inc(ws.count + b);
      ws.keys[(int64_t)b * ws.Apad + pos] = make_key(best, a);
    }
  }
}

// Same selection from already decoded fp32 tensors (the reference-style two-call API).
__global__ void select_decoded_kernel(const float* __restrict__ boxes, const float* __restrict__ obj_conf,
                                      const float* __restrict__ cls_conf, int B, int A, int C, float thr, Workspace ws) {
  const int64_t total = (int64_t)B * A;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    const float* c = cls_conf + i * C;
    float best = c[0];
    int bi = 0;
    for (int k = 1; k < C; ++k) {
      const float v = c[k];
      if (v > best) { best = v; bi = k; }
    }
    if (best >= thr) {
      ws.box[i] = reinterpret_cast<const float4*>(boxes)[i];
      ws.objc[i] = obj_conf[i]; ws.col5[i] = best; ws.score[i] = best; ws.label[i] = bi;
      const int pos = warp_agg_inc(ws.count + b);
      ws.keys[(int64_t)b * ws.Apad + pos] = make_key(best, a);
    }
  }
}

// multi_class (postprocess_utils.py:90-95): one candidate per (anchor, class) with cls_conf >= thr; the candidate id
// a*C + c in the key's low word reproduces nonzero()'s row-major order for ties.  One thread per score.
__global__ void select_multiclass_kernel(const float* __restrict__ boxes, const float* __restrict__ obj_conf,
                                         const float* __restrict__ cls_conf, int B, int A, int C, float thr, Workspace ws) {
  const int64_t per = (int64_t)A * C, total = (int64_t)B * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const int id = (int)(i - (int64_t)b * per);
    const float v = cls_conf[i];
    if (id % C == 0) {  // the anchor's box / objectness travel once
      const int64_t ia = (int64_t)b * A + id / C;
      ws.box[ia] = reinterpret_cast<const float4*>(boxes)[ia];
      ws.objc[ia] = obj_conf[ia];
    }
    if (v >= thr) {
      const int pos = warp_agg_inc(ws.count + b);
      ws.keys[(int64_t)b * ws.Apad + pos] = make_key(v, id);
    }
  }
}

// rmmop (postprocess_utils.py:74-84): top-1 class per anchor, kept when top1 >= top2 * r1 and obj^2 >= top1 * r2
// (fp32 products as torch evaluates them; no confidence threshold in this mode).
__global__ void select_rmmop_kernel(const float* __restrict__ boxes, const float* __restrict__ obj_conf,
                                    const float* __restrict__ cls_conf, int B, int A, int C, float r1, float r2,
                                    Workspace ws) {
  const int64_t total = (int64_t)B * A;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    const float* c = cls_conf + i * C;
    float best = c[0], second = -INFINITY;
    int bi = 0;
    for (int k = 1; k < C; ++k) {
      const float v = c[k];
      if (v > best) { second = best; best = v; bi = k; }
      else if (v > second) second = v;
    }
    const float o = obj_conf[i];
    if (best >= __fmul_rn(second, r1) && __fmul_rn(o, o) >= __fmul_rn(best, r2)) {
      ws.box[i] = reinterpret_cast<const float4*>(boxes)[i];
      ws.objc[i] = o; ws.col5[i] = best; ws.score[i] = best; ws.label[i] = bi;
      const int pos = warp_agg_inc(ws.count + b);
      ws.keys[(int64_t)b * ws.Apad + pos] = make_key(best, a);
    }
  }
}

template <typename T> __device__ __forceinline__ T from_f(float f);
template <> __device__ __forceinline__ float from_f<float>(float f) { return f; }
template <> __device__ __forceinline__ __half from_f<__half>(float f) { return __float2half_rn(f); }
// round-trip through T: models arithmetic carried out in the tensor's dtype
template <typename T> __device__ __forceinline__ float rnd(float f);
template <> __device__ __forceinline__ float rnd<float>(float f) { return f; }
template <> __device__ __forceinline__ float rnd<__half>(float f) { return __half2float(__float2half_rn(f)); }

// yolox.utils.postprocess selection (boxes.py:38-57): in-place cxcywh -> xyxy in the tensor dtype,
// class max, score = obj*class_conf (tensor dtype), >= thr.
template <typename T>
__global__ void select_yolox_kernel(T* __restrict__ pred, int B, int A, int C, float thr, Workspace ws) {
  const int64_t total = (int64_t)B * A;
  const int D = 5 + C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    T* p = pred + i * D;
    const float cx = ldf(p), cy = ldf(p + 1), w = ldf(p + 2), h = ldf(p + 3);
    const float hw = rnd<T>(w * 0.5f), hh = rnd<T>(h * 0.5f);  // x/2 is exact unless subnormal
    const float x1 = rnd<T>(__fsub_rn(cx, hw)), y1 = rnd<T>(__fsub_rn(cy, hh));
    const float x2 = rnd<T>(__fadd_rn(cx, hw)), y2 = rnd<T>(__fadd_rn(cy, hh));
    p[0] = from_f<T>(x1); p[1] = from_f<T>(y1); p[2] = from_f<T>(x2); p[3] = from_f<T>(y2);
    const float oc = ldf(p + 4);
    float best = ldf(p + 5);
    int bi = 0;
    for (int k = 1; k < C; ++k) {
      const float v = ldf(p + 5 + k);
      if (v > best) { best = v; bi = k; }
    }
    const float sc = rnd<T>(__fmul_rn(oc, best));
    if (sc >= rnd<T>(thr)) {
      ws.box[i] = make_float4(x1, y1, x2, y2);
      ws.objc[i] = oc; ws.col5[i] = best; ws.score[i] = sc; ws.label[i] = bi;
      const int pos = warp_agg_inc(ws.count + b);
      ws.keys[(int64_t)b * ws.Apad + pos] = make_key(sc, a);
    }
  }
}

// yolox head output assembly (+ optional decode) in dtype To from raw logits of dtype Ti.
template <typename Ti, typename To>
__global__ void head_assemble_kernel(const Ti* __restrict__ reg, int64_t reg_sb, int64_t reg_sa, const Ti* __restrict__ obj,
                                     int64_t obj_sb, int64_t obj_sa, const Ti* __restrict__ cls, int64_t cls_sb,
                                     int64_t cls_sa, int B, int A, int C, LevelsDev lv, int decode, To* __restrict__ out) {
  const int64_t total = (int64_t)B * A;
  const int D = 5 + C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A, b = i / A;
    const Ti* r = reg + b * reg_sb + a * reg_sa;
    float t0 = ldf(r), t1 = ldf(r + 1), t2 = ldf(r + 2), t3 = ldf(r + 3);
    To* o = out + i * D;
    if (decode) {  // yolo_head.py:223-224, arithmetic in the tensor dtype
      float gx, gy, s;
      anchor_geom(lv, a, &gx, &gy, &s);
      t0 = rnd<To>(__fmul_rn(rnd<To>(__fadd_rn(t0, gx)), s));
      t1 = rnd<To>(__fmul_rn(rnd<To>(__fadd_rn(t1, gy)), s));
      t2 = rnd<To>(__fmul_rn(rnd<To>(expf(t2)), s));
      t3 = rnd<To>(__fmul_rn(rnd<To>(expf(t3)), s));
    }
    o[0] = from_f<To>(t0); o[1] = from_f<To>(t1); o[2] = from_f<To>(t2); o[3] = from_f<To>(t3);
    o[4] = from_f<To>(sigmoid_f(ldf(obj + b * obj_sb + a * obj_sa)));
    const Ti* c = cls + b * cls_sb + a * cls_sa;
    for (int k = 0; k < C; ++k) o[5 + k] = from_f<To>(sigmoid_f(ldf(c + k)));
  }
}

template <typename T>
__global__ void decode_outputs_kernel(T* __restrict__ out, int B, int A, int D, LevelsDev lv) {
  const int64_t total = (int64_t)B * A;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = i % A;
    float gx, gy, s;
    anchor_geom(lv, a, &gx, &gy, &s);
    T* o = out + i * D;
    o[0] = from_f<T>(__fmul_rn(rnd<T>(__fadd_rn(ldf(o), gx)), s));
    o[1] = from_f<T>(__fmul_rn(rnd<T>(__fadd_rn(ldf(o + 1), gy)), s));
    o[2] = from_f<T>(__fmul_rn(rnd<T>(expf(ldf(o + 2))), s));
    o[3] = from_f<T>(__fmul_rn(rnd<T>(expf(ldf(o + 3))), s));
  }
}

// ------------------------------------------------------------------------------------------------
// per-image bitonic sort of the candidate keys, descending (one CTA per image)
// ------------------------------------------------------------------------------------------------
#ifdef YX_POST_DBG   // experiments: phase timeline of image 0 (thread 0), read back with yx_post_dbg_read
__device__ long long g_post_dbg[64];
#define YX_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_post_dbg[i] = clock64(); } while (0)
#else
#define YX_STAMP(i) do { } while (0)
#endif
constexpr int kSortThreads = 1024;
constexpr int kSortChunk = 8192;  // keys resident in shared memory (64 KB)

__device__ __forceinline__ void cmpx(uint64_t& a, uint64_t& b, bool desc) {
  if ((a < b) == desc) { const uint64_t t = a; a = b; b = t; }
}

// Descending bitonic sort of the 8192 keys in sk[] by 1024 threads.  Warp w owns the 256 keys sk[256 w ...], thread (w, lane)
// holds the eight keys 256 w + 32 r + lane in registers: exchange distances 1..16 are warp shuffles, 32..128 are register
// swaps, only distances >= 256 go through shared memory and a block barrier (15 of the 91 phases; the all-shared-memory
// form spent 72 us per image on 91 barriers and 64-bit bank conflicts).
__device__ __forceinline__ void bitonic_inwarp(uint64_t (&v)[8], int k, int i0 /* index of v[0] */, int lane) {
  // desc(i): the pair whose lower index is i sorts descending iff (i & k) == 0; the bit k of i is the same for both
  // elements of a pair (k > distance)
  if (k > 128) {
#pragma unroll
    for (int r = 0; r < 4; ++r) cmpx(v[r], v[r + 4], ((i0 + 32 * r) & k) == 0);
  }
  if (k > 64) {
#pragma unroll
    for (int r = 0; r < 8; ++r) if ((r & 2) == 0) cmpx(v[r], v[r + 2], ((i0 + 32 * r) & k) == 0);
  }
  if (k > 32) {
#pragma unroll
    for (int r = 0; r < 8; r += 2) cmpx(v[r], v[r + 1], ((i0 + 32 * r) & k) == 0);
  }
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    if (k > j) {
      const bool lower = (lane & j) == 0;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, v[r], j);
        const bool desc = ((i0 + 32 * r) & k) == 0;
        const bool keep_max = desc == lower;
        v[r] = ((v[r] > o) == keep_max) ? v[r] : o;
      }
    }
  }
}

__device__ __forceinline__ void bitonic_sort_8192_desc(uint64_t* sk) {
  const int lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 256, i0 = wbase + lane;
  uint64_t v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = sk[i0 + 32 * r];
  for (int k = 2; k <= 256; k <<= 1) bitonic_inwarp(v, k, i0, lane);
  for (int k = 512; k <= kSortChunk; k <<= 1) {
#pragma unroll
    for (int r = 0; r < 8; ++r) sk[i0 + 32 * r] = v[r];
    __syncthreads();
    for (int j = k >> 1; j >= 256; j >>= 1) {
      for (int t = threadIdx.x; t < (kSortChunk >> 1); t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        cmpx(sk[i], sk[i | j], ((i & k) == 0));
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = sk[i0 + 32 * r];
    bitonic_inwarp(v, k, i0, lane);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) sk[i0 + 32 * r] = v[r];
  __syncthreads();
}

__global__ void __launch_bounds__(kSortThreads, 1) sort_keys_kernel(Workspace ws, int max_nms) {
  extern __shared__ uint64_t sk[];
  __shared__ int s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_cnt;
  const int b = blockIdx.x;
  YX_STAMP(0);
  const int n = ws.count[b];
  if (n <= 1) return;
  uint64_t* keys = ws.keys + (int64_t)b * ws.Apad;

  if (max_nms > 0 && n > max_nms && max_nms <= kSortChunk) {
    // ---- top-k first (postprocess_utils.py:100-103 keeps the max_nms best): MSB-first radix SELECT of the
    // k-th largest 64-bit key (up to 8 passes of 8 bits over L2-resident keys), then only those k keys are sorted,
    // entirely in shared memory / registers.  Keys are unique (anchor index in the low word), so ">= pivot" is exactly k.
    // The select stops at the first pass whose boundary bin is needed WHOLE (pivot = that prefix with the remaining bits
    // zero): three passes instead of eight on tie-free scores.
    const int lane = threadIdx.x & 31;
    __shared__ int s_done;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_need = max_nms; s_done = 0; }
    int shift_done = 0;
    for (int pass = 0; pass < 8; ++pass) {
      const int shift = 56 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += kSortThreads) s_hist[i] = 0;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += kSortThreads) {
        const unsigned long long key = keys[i];
        // (plain shared-memory atomics: a __match_any_sync aggregation of same-bin adds measured 3x SLOWER per pass)
        if (pass == 0 || (key >> (shift + 8)) == prefix) atomicAdd(&s_hist[(key >> shift) & 255ull], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int need = s_need, bin = 255;
        for (; bin > 0; --bin) {
          if (s_hist[bin] >= need) break;
          need -= s_hist[bin];
        }
        s_need = need;
        s_prefix = (prefix << 8) | (unsigned long long)bin;
        if (s_hist[bin] == need) s_done = 1;   // every key of this bin is selected: no need to look at lower bits
      }
      __syncthreads();
      YX_STAMP(1 + pass);
      shift_done = shift;
      if (s_done) break;
    }
    const unsigned long long pivot = s_prefix << shift_done;
    if (threadIdx.x == 0) s_cnt = 0;
    for (int i = threadIdx.x; i < kSortChunk; i += kSortThreads) sk[i] = 0ull;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 4 * kSortThreads) {   // compaction: four loads in flight, one shared-memory atomic per warp and key slot
      unsigned long long key[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSortThreads + threadIdx.x;
        key[u] = i < n ? keys[i] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool sel = i0 + u * kSortThreads + (int)threadIdx.x < n && key[u] >= pivot;
        const unsigned bal = __ballot_sync(0xffffffffu, sel);
        if (bal) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_cnt, __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (sel) sk[base + __popc(bal & ((1u << lane) - 1u))] = key[u];
        }
      }
    }
    __syncthreads();
    YX_STAMP(9);
    bitonic_sort_8192_desc(sk);
    YX_STAMP(10);
    for (int i = threadIdx.x; i < max_nms; i += kSortThreads) keys[i] = sk[i];
    YX_STAMP(11);
    return;
  }

  int N = 2;
  while (N < n) N <<= 1;
  for (int i = n + threadIdx.x; i < N; i += kSortThreads) keys[i] = 0;  // pads sort last
  __syncthreads();
  const int chunk = min(N, kSortChunk);
  // phase 1: fully sort every chunk in shared memory (k = 2 .. chunk)
  for (int c0 = 0; c0 < N; c0 += chunk) {
    for (int i = threadIdx.x; i < chunk; i += kSortThreads) sk[i] = keys[c0 + i];
    __syncthreads();
    for (int k = 2; k <= chunk; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (chunk >> 1); t += kSortThreads) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          cmpx(sk[i], sk[i | j], (((c0 + i) & k) == 0));
        }
        __syncthreads();
      }
    for (int i = threadIdx.x; i < chunk; i += kSortThreads) keys[c0 + i] = sk[i];
    __syncthreads();
  }
  // phase 2: merge across chunks; strides >= chunk go through global memory (L2 resident)
  for (int k = chunk << 1; k <= N; k <<= 1) {
    for (int j = k >> 1; j >= chunk; j >>= 1) {
      for (int t = threadIdx.x; t < (N >> 1); t += kSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        uint64_t a = keys[i], c = keys[i | j];
        const uint64_t a0 = a;
        cmpx(a, c, ((i & k) == 0));
        if (a != a0) { keys[i] = a; keys[i | j] = c; }
      }
      __syncthreads();
    }
    for (int c0 = 0; c0 < N; c0 += chunk) {
      for (int i = threadIdx.x; i < chunk; i += kSortThreads) sk[i] = keys[c0 + i];
      __syncthreads();
      for (int j = chunk >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (chunk >> 1); t += kSortThreads) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
          cmpx(sk[i], sk[i | j], (((c0 + i) & k) == 0));
        }
        __syncthreads();
      }
      for (int i = threadIdx.x; i < chunk; i += kSortThreads) keys[c0 + i] = sk[i];
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// batched greedy NMS: one CTA per image, candidates visited in score order in blocks of 64;
// each block is tested against the kept list (16 threads per candidate), then resolved with a
// 64x64 suppression bitmask held in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 1024;
constexpr int kKeptSmem = 1024;  // kept boxes held in shared memory when max_det <= this

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float left = fmaxf(a.x, b.x), top = fmaxf(a.y, b.y);
  const float right = fminf(a.z, b.z), bottom = fminf(a.w, b.w);
  const float w = fmaxf(0.0f, __fsub_rn(right, left)), h = fmaxf(0.0f, __fsub_rn(bottom, top));
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
  const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
  return ovr > thr;
}

// Where the NMS tail also stores an image's rows when the detections are gathered across GPUs: rank w's receive window,
// mapped into this process through CUDA IPC (stores travel over NVLink / NVSwitch as posted writes).
struct PeerDev {
  float* det[YX_MAX_PEERS];  // this rank's [B, det_rows, 7] block inside rank w's window
  int* cnt[YX_MAX_PEERS];    // this rank's [B] block inside rank w's window
  int* arrive[YX_MAX_PEERS]; // rank w's arrival counter for this rank
  int world;                 // 0: no gather
};

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_kernel(Workspace ws, int A, float nms_thr, int max_nms, int max_det, int mode, int det_rows, float* __restrict__ det,
           int* __restrict__ det_count, int* __restrict__ det_anchor, const PeerDev po) {
  __shared__ float4 s_kbox[kKeptSmem];
  __shared__ int s_klab[kKeptSmem];
  __shared__ float4 s_cbox[64];
  __shared__ int s_clab[64];
  __shared__ unsigned long long s_mask[64];
  __shared__ int s_pre[64];
  __shared__ float s_red[32];
  __shared__ int s_kept;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  YX_STAMP(16);
  int n = ws.count[b];
  if (max_nms > 0 && n > max_nms) n = max_nms;
  if (mode == YX_NMS_AUTO) mode = (4 * (int64_t)n > 100000) ? YX_NMS_VANILLA : YX_NMS_TRICK;
  const uint64_t* keys = ws.keys + (int64_t)b * ws.Apad;
  const int64_t ib = (int64_t)b * A, ic = (int64_t)b * ws.R;
  const int K = ws.K;
  float4* sbox = ws.sbox + ic;
  int* slab = ws.slabel + ic;
  int* kidx = ws.kidx + ic;
  const bool kept_in_smem = (max_det > 0 && max_det <= kKeptSmem);
  float4* kbox = kept_in_smem ? s_kbox : (ws.kbox + ic);
  int* klab = kept_in_smem ? s_klab : (ws.klabel + ic);
  const int cap = max_det > 0 ? max_det : 0x7fffffff;

  // ---- gather the selected candidates in score order (four independent key -> box chains in flight per thread) ------
  float mx = -INFINITY;
  for (int i0 = 0; i0 < n; i0 += 4 * kNmsThreads) {
    int id[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kNmsThreads + tid;
      id[u] = i < n ? key_id(keys[i]) : -1;
    }
    float4 bx[4];
    int lb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = K > 1 ? id[u] / K : id[u];
      if (id[u] >= 0) {
        bx[u] = ws.box[ib + a];
        lb[u] = K > 1 ? id[u] % K : ws.label[ib + a];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * kNmsThreads + tid;
      if (id[u] >= 0) {
        sbox[i] = bx[u];
        slab[i] = lb[u];
        mx = fmaxf(fmaxf(mx, fmaxf(bx[u].x, bx[u].y)), fmaxf(bx[u].z, bx[u].w));
      }
    }
  }
  YX_STAMP(17);
  // coordinate trick (torchvision batched_nms): boxes + label * (boxes.max() + 1).  The offset is added when a block of
  // candidates is loaded below (same two roundings as adding it here), which saves a second pass over the sorted boxes.
  float m1 = 0.0f;
  if (mode == YX_NMS_TRICK) {
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red[wid] = mx;
    __syncthreads();
    mx = s_red[0];
    for (int i = 1; i < kNmsThreads / 32; ++i) mx = fmaxf(mx, s_red[i]);
    m1 = __fadd_rn(mx, 1.0f);  // boxes.max() + 1
  }
  if (tid == 0) s_kept = 0;
  __syncthreads();
  YX_STAMP(18);

  // ---- greedy pass -------------------------------------------------------------------------------
  const bool same_class_only = (mode == YX_NMS_VANILLA);
  const int ci = tid >> 4, sub = tid & 15;  // candidate slot and sub-lane within the 16-thread group
  for (int base = 0; base < n; base += 64) {
    const int kept = s_kept;
    if (kept >= cap) break;
    const int m = min(64, n - base);
    if (tid < 64) {
      if (tid < m) {
        float4 bx = sbox[base + tid];
        const int lb = slab[base + tid];
        if (mode == YX_NMS_TRICK) {
          const float off = __fmul_rn((float)lb, m1);
          bx.x = __fadd_rn(bx.x, off); bx.y = __fadd_rn(bx.y, off); bx.z = __fadd_rn(bx.z, off); bx.w = __fadd_rn(bx.w, off);
        }
        s_cbox[tid] = bx; s_clab[tid] = lb;
      }
    }
    __syncthreads();
    if (base == 64) YX_STAMP(24);
    bool sup = false;
    unsigned long long bits = 0ull;
    if (ci < m) {
      const float4 cb = s_cbox[ci];
      const int cl = s_clab[ci];
      for (int k = sub; k < kept; k += 16) {
        if (same_class_only && klab[k] != cl) continue;
        if (iou_gt(kbox[k], cb, nms_thr)) { sup = true; break; }
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = sub * 4 + jj;
        if (j < ci && (!same_class_only || s_clab[j] == cl) && iou_gt(s_cbox[j], cb, nms_thr)) bits |= (1ull << j);
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, sup);
    const bool any_sup = ((bal >> (lane & 16)) & 0xFFFFu) != 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
    if (base == 64) YX_STAMP(25);
    if (sub == 0 && ci < m) { s_pre[ci] = any_sup ? 1 : 0; s_mask[ci] = bits; }
    __syncthreads();
    if (base == 64) YX_STAMP(26);
    // Resolve the block in score order.  The chain (candidate i survives iff no EARLIER SURVIVOR of the block suppresses
    // it) is inherently serial; every lane of warp 0 runs it redundantly on the 64 masks read straight from shared memory
    // (broadcast loads, independent of the chain, so the unrolled loop issues them ahead) with a branch-free update: one
    // AND / test / OR per candidate.  (A thread-0 loop that also appended to the kept list took ~300 cycles per candidate,
    // a shuffle-fed loop ~85; this one ~20.)  The survivors are then appended to the kept list by all lanes at once
    // (position = popcount of the lower survivor bits).
    if (wid == 0) {
      const unsigned dead_lo = __ballot_sync(0xffffffffu, lane >= m || s_pre[lane] != 0);
      const unsigned dead_hi = __ballot_sync(0xffffffffu, lane + 32 >= m || s_pre[lane + 32] != 0);
      const unsigned long long dead = ((unsigned long long)dead_hi << 32) | dead_lo;
      unsigned klo = 0u, khi = 0u;   // survivors of the block (a row's mask only holds EARLIER candidates: rows < 32 have no high half)
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const uint2 mi = *reinterpret_cast<const uint2*>(&s_mask[i]);   // (rows >= m are stale: their dead bit is set)
        const unsigned hit = i < 32 ? (mi.x & klo) : ((mi.x & klo) | (mi.y & khi));
        const bool take = !((dead >> i) & 1ull) && hit == 0u;
        if (i < 32) klo |= take ? (1u << i) : 0u;
        else khi |= take ? (1u << (i - 32)) : 0u;
      }
      unsigned long long keptmask = ((unsigned long long)khi << 32) | klo;
      // the cap (max_det) is applied afterwards: survivors do not depend on later candidates, so keeping the FIRST
      // `cap - kept` of them equals stopping the scan there
      while (__popcll(keptmask) > cap - kept) keptmask &= ~(1ull << (63 - __clzll(keptmask)));
      const int kn = kept + __popcll(keptmask);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        if ((keptmask >> c) & 1ull) {
          const int pos = kept + __popcll(keptmask & ((1ull << c) - 1ull));
          kbox[pos] = s_cbox[c]; klab[pos] = s_clab[c]; kidx[pos] = base + c;
        }
      }
      if (lane == 0) s_kept = kn;
    }
    if (base == 64) YX_STAMP(27);
    __syncthreads();
    if (base == 64) YX_STAMP(28);
    if (base == 0) YX_STAMP(23);
  }

  // ---- write detections (score order) -----------------------------------------------------------
  YX_STAMP(19);
  const int kept = s_kept;
  if (tid == 0) det_count[b] = kept;
  float* drow = det + (int64_t)b * det_rows * 7;
  for (int k = tid; k < det_rows; k += kNmsThreads) {
    float* d = drow + (int64_t)k * 7;
    if (k < kept) {
      const int r = kidx[k];
      const uint64_t key = keys[r];
      const int id = key_id(key);
      const int a = K > 1 ? id / K : id;
      const float4 bx = ws.box[ib + a];
      d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
      d[4] = ws.objc[ib + a];
      if (K > 1) { d[5] = key_score(key); d[6] = (float)(id % K); }
      else { d[5] = ws.col5[ib + a]; d[6] = (float)ws.label[ib + a]; }
      if (det_anchor) det_anchor[(int64_t)b * det_rows + k] = a;
    } else {
#pragma unroll
      for (int j = 0; j < 7; ++j) d[j] = 0.0f;
      if (det_anchor) det_anchor[(int64_t)b * det_rows + k] = -1;
    }
    // fused all-gather: the same row goes into every rank's window while it is still in registers / L1
    for (int w = 0; w < po.world; ++w) {
      float* r = po.det[w] + ((int64_t)b * det_rows + k) * 7;
#pragma unroll
      for (int j = 0; j < 7; ++j) r[j] = d[j];
    }
  }
  YX_STAMP(20);
  if (po.world > 0) {
    if (tid < po.world) po.cnt[tid][b] = kept;
    __threadfence_system();  // every thread: its own rows / count are ordered before what follows, system-wide
    __syncthreads();
    if (tid < po.world) {
      __threadfence_system();  // release by the signalling thread AFTER the barrier (cumulative over the CTA's stores)
      atomicAdd_system(po.arrive[tid], 1);
    }
  }
}

// Completes the gather on the receiving side: returns (stream-ordered) once every rank's images of the awaited step have
// arrived in this GPU's window.  A bounded wait: on timeout *status = 1 + late rank, the late rank's counts in the
// awaited window are zeroed (so no stale row can be read as a detection) and the kernel returns.
__global__ void peer_wait_kernel(const int* arrive, int world, int target, int* status, long long timeout_cycles,
                                 int* wait_cnt, int B) {
  const int lane = threadIdx.x;
  if (lane < world) {
    const long long t0 = clock64();
    const volatile int* f = arrive + lane;
    while (*f - target < 0) {  // counters only grow; difference form tolerates wrap-around
      __nanosleep(200);
      if (clock64() - t0 > timeout_cycles) {
        atomicExch(status, 1 + lane);
        if (wait_cnt != nullptr)
          for (int i = 0; i < B; ++i) wait_cnt[lane * B + i] = 0;
        break;
      }
    }
  }
  __threadfence_system();
}

static int launch_peer_wait(const void* local_arrive, int world, int target, void* status, int timeout_ms, void* wait_cnt,
                            int B, cudaStream_t st) {
  const long long timeout = timeout_ms > 0 ? (long long)timeout_ms * 2000000ll : 20000000000ll;  // ~2 GHz
  peer_wait_kernel<<<1, 32, 0, st>>>(static_cast<const int*>(local_arrive), world, target, static_cast<int*>(status), timeout,
                                     static_cast<int*>(wait_cnt), B);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

// ------------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------------
static int grid_for(int64_t total, int threads) {
  return (int)std::min<int64_t>((total + threads - 1) / threads, 148 * 8);
}

static int sort_and_nms(const Workspace& ws, int B, int A, float nms_thr, int max_nms, int max_det, int mode,
                        int det_rows, float* det, int32_t* det_count, int32_t* det_anchor, cudaStream_t st,
                        const PeerDev* peer = nullptr) {
  static bool attr = false;
  if (!attr) {
    YX_CUDA(cudaFuncSetAttribute(sort_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortChunk * 8));
    attr = true;
  }
  sort_keys_kernel<<<B, kSortThreads, kSortChunk * 8, st>>>(ws, max_nms);
  YX_CUDA(cudaGetLastError());
  PeerDev po;
  if (peer) po = *peer; else po.world = 0;
  nms_kernel<<<B, kNmsThreads, 0, st>>>(ws, A, nms_thr, max_nms, max_det, mode, det_rows, det, det_count, det_anchor, po);
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

static int64_t cand_rows(int A, int K, int max_nms) {
  const int64_t all = (int64_t)A * K;
  return (K > 1 && max_nms > 0) ? std::min<int64_t>(all, max_nms) : all;
}

static int check_ws(int B, int A, void* workspace, size_t bytes, Workspace* ws, int K = 1, int max_nms = 0) {
  YX_REQUIRE(B >= 1 && A >= 1, "B, A must be positive");
  YX_REQUIRE((int64_t)A * K < (int64_t(1) << 30), "too many candidates per image");
  YX_REQUIRE(workspace != nullptr && ((uintptr_t)workspace % 256) == 0, "workspace must be 256-byte aligned");
  const size_t need = ws_layout(B, A, K, cand_rows(A, K, max_nms), static_cast<uint8_t*>(workspace), ws);
  YX_REQUIRE(bytes >= need, "workspace too small (see yx_detect_workspace_bytes)");
  return YX_OK;
}

}  // namespace yx

using namespace yx;

#ifdef YX_POST_DBG
extern "C" int yx_post_dbg_read(long long* out64) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out64, g_post_dbg, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" size_t yx_detect_workspace_bytes(int B, int A) {
  if (B < 1 || A < 1) return 0;
  return ws_layout(B, A, 1, A, nullptr, nullptr);
}

extern "C" size_t yx_nms_workspace_bytes(int B, int A, int C, int cand_mode, int max_nms) {
  if (B < 1 || A < 1 || C < 1) return 0;
  const int K = cand_mode == YX_CAND_MULTI_CLASS ? C : 1;
  if ((int64_t)A * K >= (int64_t(1) << 30)) return 0;
  return ws_layout(B, A, K, cand_rows(A, K, max_nms), nullptr, nullptr);
}

static int decode_infer_impl(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb, int64_t obj_sa,
                             const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B, int A, int C,
                             const LevelsDev& lv, const void* grids, const void* scales, int geom_dtype, float* boxes,
                             float* obj_conf, float* cls_conf, void* stream) {
  YX_REQUIRE(B >= 1 && C >= 1, "B, C must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((int64_t)B * A, 256);
  if (logits_dtype == YX_F16)
    decode_infer_kernel<__half><<<g, 256, 0, st>>>((const __half*)reg, reg_sb, reg_sa, (const __half*)obj, obj_sb, obj_sa,
                                                   (const __half*)cls, cls_sb, cls_sa, B, A, C, lv, grids, scales, geom_dtype,
                                                   boxes, obj_conf, cls_conf);
  else if (logits_dtype == YX_F32)
    decode_infer_kernel<float><<<g, 256, 0, st>>>((const float*)reg, reg_sb, reg_sa, (const float*)obj, obj_sb, obj_sa,
                                                  (const float*)cls, cls_sb, cls_sa, B, A, C, lv, grids, scales, geom_dtype,
                                                  boxes, obj_conf, cls_conf);
  else
    YX_REQUIRE(false, "logits dtype must be YX_F16 or YX_F32");
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

extern "C" int yx_decode_infer(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                               int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B,
                               int A, int C, const yx_levels* lv_host, float* boxes, float* obj_conf, float* cls_conf,
                               void* stream) {
  LevelsDev lv;
  int rc = make_levels(lv_host, A, &lv);
  if (rc) return rc;
  return decode_infer_impl(reg, reg_sb, reg_sa, obj, obj_sb, obj_sa, cls, cls_sb, cls_sa, logits_dtype, B, A, C, lv, nullptr,
                           nullptr, YX_F32, boxes, obj_conf, cls_conf, stream);
}

// Same decode with the anchor geometry READ from the caller's grids (1,A,2) / scales (1,A,1) device tensors, exactly as
// postprocess_utils.py:37-38 uses them (any values, e.g. a custom grid offset); no pyramid description is needed.
extern "C" int yx_decode_infer_grids(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                                     int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B,
                                     int A, int C, const void* grids, const void* scales, int geom_dtype, float* boxes,
                                     float* obj_conf, float* cls_conf, void* stream) {
  YX_REQUIRE(grids != nullptr && scales != nullptr, "grids / scales must be device pointers");
  YX_REQUIRE(geom_dtype == YX_F16 || geom_dtype == YX_F32, "grids / scales dtype must be YX_F16 or YX_F32");
  LevelsDev lv;
  memset(&lv, 0, sizeof lv);
  return decode_infer_impl(reg, reg_sb, reg_sa, obj, obj_sb, obj_sa, cls, cls_sb, cls_sa, logits_dtype, B, A, C, lv, grids,
                           scales, geom_dtype, boxes, obj_conf, cls_conf, stream);
}

extern "C" int yx_nms_main_ex(const float* boxes, const float* obj_conf, const float* cls_conf, int B, int A, int C,
                              float conf_thr, float nms_thr, int max_nms, int max_det, int mode, int cand_mode,
                              float rmmop_r1, float rmmop_r2, void* workspace, size_t workspace_bytes, float* det,
                              int32_t* det_count, int32_t* det_anchor, void* stream) {
  YX_REQUIRE(mode >= 0 && mode <= 3 && C >= 1, "bad nms mode / C");
  YX_REQUIRE(cand_mode >= YX_CAND_MAX && cand_mode <= YX_CAND_RMMOP, "bad candidate mode");
  YX_REQUIRE(cand_mode != YX_CAND_RMMOP || C >= 2, "rmmop needs at least two classes");  // cls_conf_sorted[:, 1]
  const int K = cand_mode == YX_CAND_MULTI_CLASS ? C : 1;
  Workspace ws;
  int rc = check_ws(B, A, workspace, workspace_bytes, &ws, K, max_nms);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t all = (int64_t)A * K;
  const int det_rows = max_det > 0 ? max_det : (int)all;
  YX_CUDA(cudaMemsetAsync(ws.count, 0, (size_t)B * 4, st));
  if (cand_mode == YX_CAND_MULTI_CLASS)
    select_multiclass_kernel<<<grid_for((int64_t)B * all, 256), 256, 0, st>>>(boxes, obj_conf, cls_conf, B, A, C, conf_thr, ws);
  else if (cand_mode == YX_CAND_RMMOP)
    select_rmmop_kernel<<<grid_for((int64_t)B * A, 256), 256, 0, st>>>(boxes, obj_conf, cls_conf, B, A, C, rmmop_r1, rmmop_r2, ws);
  else
    select_decoded_kernel<<<grid_for((int64_t)B * A, 256), 256, 0, st>>>(boxes, obj_conf, cls_conf, B, A, C, conf_thr, ws);
  YX_CUDA(cudaGetLastError());
  return sort_and_nms(ws, B, A, nms_thr, max_nms, max_det, mode, det_rows, det, det_count, det_anchor, st);
}

extern "C" int yx_nms_main(const float* boxes, const float* obj_conf, const float* cls_conf, int B, int A, int C,
                           float conf_thr, float nms_thr, int max_nms, int max_det, int mode, void* workspace,
                           size_t workspace_bytes, float* det, int32_t* det_count, int32_t* det_anchor, void* stream) {
  return yx_nms_main_ex(boxes, obj_conf, cls_conf, B, A, C, conf_thr, nms_thr, max_nms, max_det, mode, YX_CAND_MAX, 0.0f,
                        0.0f, workspace, workspace_bytes, det, det_count, det_anchor, stream);
}

static int detect_main_impl(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                            int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B,
                            int A, int C, const yx_levels* lv_host, float conf_thr, float nms_thr, int max_nms,
                            int max_det, int mode, void* workspace, size_t workspace_bytes, float* det,
                            int32_t* det_count, int32_t* det_anchor, void* stream, const yx::PeerDev* peer) {
  LevelsDev lv;
  int rc = make_levels(lv_host, A, &lv);
  if (rc) return rc;
  Workspace ws;
  rc = check_ws(B, A, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  YX_REQUIRE(mode >= 0 && mode <= 3 && C >= 1, "bad nms mode / C");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int det_rows = max_det > 0 ? max_det : A;
  YX_CUDA(cudaMemsetAsync(ws.count, 0, (size_t)B * 4, st));
  const int g = grid_for((int64_t)B * A, 128);
  if (logits_dtype == YX_F16) {
    const bool vec = (C % 8 == 0) && (cls_sa % 8 == 0) && (cls_sb % 8 == 0) && (((uintptr_t)cls) % 16 == 0);
    if (vec)
      select_infer_kernel<__half, true><<<g, 128, 0, st>>>((const __half*)reg, reg_sb, reg_sa, (const __half*)obj, obj_sb,
                                                           obj_sa, (const __half*)cls, cls_sb, cls_sa, B, A, C, lv, conf_thr, ws);
    else
      select_infer_kernel<__half, false><<<g, 128, 0, st>>>((const __half*)reg, reg_sb, reg_sa, (const __half*)obj, obj_sb,
                                                            obj_sa, (const __half*)cls, cls_sb, cls_sa, B, A, C, lv, conf_thr, ws);
  } else if (logits_dtype == YX_F32) {
    select_infer_kernel<float, false><<<g, 128, 0, st>>>((const float*)reg, reg_sb, reg_sa, (const float*)obj, obj_sb, obj_sa,
                                                         (const float*)cls, cls_sb, cls_sa, B, A, C, lv, conf_thr, ws);
  } else {
    YX_REQUIRE(false, "logits dtype must be YX_F16 or YX_F32");
  }
  YX_CUDA(cudaGetLastError());
  return sort_and_nms(ws, B, A, nms_thr, max_nms, max_det, mode, det_rows, det, det_count, det_anchor, st, peer);
}

extern "C" int yx_detect_main(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                              int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype, int B,
                              int A, int C, const yx_levels* lv_host, float conf_thr, float nms_thr, int max_nms,
                              int max_det, int mode, void* workspace, size_t workspace_bytes, float* det,
                              int32_t* det_count, int32_t* det_anchor, void* stream) {
  return detect_main_impl(reg, reg_sb, reg_sa, obj, obj_sb, obj_sa, cls, cls_sb, cls_sa, logits_dtype, B, A, C, lv_host,
                          conf_thr, nms_thr, max_nms, max_det, mode, workspace, workspace_bytes, det, det_count, det_anchor,
                          stream, nullptr);
}

extern "C" int yx_detect_main_gather(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                                     int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int logits_dtype,
                                     int B, int A, int C, const yx_levels* lv_host, float conf_thr, float nms_thr,
                                     int max_nms, int max_det, int mode, void* workspace, size_t workspace_bytes,
                                     float* det, int32_t* det_count, int32_t* det_anchor, const yx_peer_out* peer,
                                     void* stream) {
  YX_REQUIRE(peer != nullptr && peer->world >= 1 && peer->world <= YX_MAX_PEERS, "peer: world must be 1..YX_MAX_PEERS");
  YX_REQUIRE(max_det > 0, "gathered detections need a fixed row count (max_det > 0)");
  YX_REQUIRE(peer->local_arrive != nullptr && peer->status != nullptr, "peer: local_arrive / status missing");
  PeerDev po;
  po.world = peer->world;
  for (int w = 0; w < peer->world; ++w) {
    YX_REQUIRE(peer->det[w] && peer->cnt[w] && peer->arrive[w], "peer: missing window pointer");
    po.det[w] = static_cast<float*>(peer->det[w]);
    po.cnt[w] = static_cast<int*>(peer->cnt[w]);
    po.arrive[w] = static_cast<int*>(peer->arrive[w]);
  }
  YX_REQUIRE(peer->wait_mode >= YX_PEER_WAIT_NONE && peer->wait_mode <= YX_PEER_WAIT_BEFORE, "peer: bad wait_mode");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  // BEFORE: the awaited step is an EARLIER one (its rows arrived while this step's network ran), so the wait costs nothing
  // and no rank is lock-stepped to the slowest one; the consumer reads that earlier step's window after this call.
  if (peer->wait_mode == YX_PEER_WAIT_BEFORE &&
      (rc = launch_peer_wait(peer->local_arrive, peer->world, peer->wait_target, peer->status, peer->timeout_ms,
                             peer->wait_cnt, B, st)) != YX_OK)
    return rc;
  rc = detect_main_impl(reg, reg_sb, reg_sa, obj, obj_sb, obj_sa, cls, cls_sb, cls_sa, logits_dtype, B, A, C, lv_host,
                        conf_thr, nms_thr, max_nms, max_det, mode, workspace, workspace_bytes, det, det_count,
                        det_anchor, stream, &po);
  if (rc) return rc;
  if (peer->wait_mode == YX_PEER_WAIT_AFTER)
    return launch_peer_wait(peer->local_arrive, peer->world, peer->wait_target, peer->status, peer->timeout_ms,
                            peer->wait_cnt, B, st);
  return YX_OK;
}

extern "C" int yx_peer_wait(const void* local_arrive, int world, int wait_target, void* status, int timeout_ms,
                            void* wait_cnt, int B, void* stream) {
  YX_REQUIRE(local_arrive != nullptr && status != nullptr && world >= 1 && world <= YX_MAX_PEERS && B >= 0, "bad peer wait arguments");
  return launch_peer_wait(local_arrive, world, wait_target, status, timeout_ms, wait_cnt, B, static_cast<cudaStream_t>(stream));
}

extern "C" int yx_head_assemble(const void* reg, int64_t reg_sb, int64_t reg_sa, const void* obj, int64_t obj_sb,
                                int64_t obj_sa, const void* cls, int64_t cls_sb, int64_t cls_sa, int B, int A, int C,
                                const yx_levels* lv_host, int decode, void* out, int out_dtype, void* stream) {
  LevelsDev lv;
  int rc = make_levels(lv_host, A, &lv);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((int64_t)B * A, 128);
  // engine logits are fp16
  if (out_dtype == YX_F16)
    head_assemble_kernel<__half, __half><<<g, 128, 0, st>>>((const __half*)reg, reg_sb, reg_sa, (const __half*)obj, obj_sb,
                                                            obj_sa, (const __half*)cls, cls_sb, cls_sa, B, A, C, lv, decode,
                                                            (__half*)out);
  else if (out_dtype == YX_F32)
    head_assemble_kernel<__half, float><<<g, 128, 0, st>>>((const __half*)reg, reg_sb, reg_sa, (const __half*)obj, obj_sb,
                                                           obj_sa, (const __half*)cls, cls_sb, cls_sa, B, A, C, lv, decode,
                                                           (float*)out);
  else
    YX_REQUIRE(false, "out dtype must be YX_F16 or YX_F32");
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

extern "C" int yx_decode_outputs(void* outputs, int dtype, int B, int A, int C, const yx_levels* lv_host, void* stream) {
  LevelsDev lv;
  int rc = make_levels(lv_host, A, &lv);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((int64_t)B * A, 256);
  if (dtype == YX_F16)
    decode_outputs_kernel<__half><<<g, 256, 0, st>>>((__half*)outputs, B, A, 5 + C, lv);
  else if (dtype == YX_F32)
    decode_outputs_kernel<float><<<g, 256, 0, st>>>((float*)outputs, B, A, 5 + C, lv);
  else
    YX_REQUIRE(false, "dtype must be YX_F16 or YX_F32");
  YX_CUDA(cudaGetLastError());
  return YX_OK;
}

extern "C" int yx_postprocess_yolox(void* prediction, int dtype, int B, int A, int C, float conf_thr, float nms_thr,
                                    int mode, void* workspace, size_t workspace_bytes, float* det, int32_t* det_count,
                                    int32_t* det_anchor, void* stream) {
  Workspace ws;
  int rc = check_ws(B, A, workspace, workspace_bytes, &ws);
  if (rc) return rc;
  YX_REQUIRE(mode >= 0 && mode <= 3 && C >= 1, "bad nms mode / C");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YX_CUDA(cudaMemsetAsync(ws.count, 0, (size_t)B * 4, st));
  const int g = grid_for((int64_t)B * A, 128);
  if (dtype == YX_F16)
    select_yolox_kernel<__half><<<g, 128, 0, st>>>((__half*)prediction, B, A, C, conf_thr, ws);
  else if (dtype == YX_F32)
    select_yolox_kernel<float><<<g, 128, 0, st>>>((float*)prediction, B, A, C, conf_thr, ws);
  else
    YX_REQUIRE(false, "dtype must be YX_F16 or YX_F32");
  YX_CUDA(cudaGetLastError());
  return sort_and_nms(ws, B, A, nms_thr, /*max_nms=*/0, /*max_det=*/0, mode, A, det, det_count, det_anchor, st);
}
