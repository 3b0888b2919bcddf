/* TEST INFRASTRUCTURE ONLY — never linked into or called by the product library.
 *
 * Plain-C CPU restatement of the reference's post-processing path:
 *   decode (infer flavour)   choijhanyangackr/yolox_infer/postprocess_utils.py:27-52
 *   decode (yolox flavour)   yolox/models/yolo_head.py:210-225
 *   candidate selection      postprocess_utils.py:86-103 ; yolox/utils/boxes.py:38-59
 *   greedy NMS               torchvision.ops.nms CPU kernel (third-party, torchvision 0.26.0;
 *                            not vendored in /root/reference; called at boxes.py:62,68 and
 *                            yolox_infer/nms.py:19,40): stable descending sort, suppress j>i when
 *                            inter/(area_i+area_j-inter) > thr (strict), areas=(x2-x1)*(y2-y1)
 *   batched NMS              torchvision.ops.boxes.batched_nms: "coordinate trick"
 *                            (boxes + label*(max_coord+1)) or "vanilla" (per class, then re-sorted
 *                            by score) — both restated; the caller picks like torchvision does.
 *
 * Parity pin: tests/test_oracle_post.py checks these against torchvision's own CPU ops on seeded
 * inputs and against tests/golden/post_*.npz produced by the reference (make_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libpost_ref.so oracle/post_ref.c -lm
 * -ffp-contract=off matters: the IoU test must round every operation like torchvision's kernel.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- fp16 -> fp32 ------------------------------------------------------------------- */
static float h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) bits = sign;
    else {
      int e = -1;
      do { man <<= 1; e++; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(112 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
  else bits = sign | ((exp + 112) << 23) | (man << 13);
  float f; memcpy(&f, &bits, 4); return f;
}

static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* ---- decode, infer flavour -------------------------------------------------------------
 * reg [A,4], obj [A], cls [A,C] logits (fp16 bit patterns if is_half else fp32), one image.
 * level l covers anchors row-major (y*W+x) of an (lh[l] x lw[l]) map with stride ls[l].
 * out: boxes [A,4] xyxy, obj_conf [A], cls_conf [A,C] (= sigmoid(cls)*sigmoid(obj)).  */
void yxref_decode_infer(const void* reg, const void* obj, const void* cls, int is_half, int A, int C,
                        int n_levels, const int* lh, const int* lw, const int* ls,
                        float* boxes, float* obj_conf, float* cls_conf) {
  int a = 0;
  for (int l = 0; l < n_levels; ++l)
    for (int y = 0; y < lh[l]; ++y)
      for (int x = 0; x < lw[l]; ++x, ++a) {
        float t[4], o;
        for (int k = 0; k < 4; ++k)
          t[k] = is_half ? h2f(((const uint16_t*)reg)[a * 4 + k]) : ((const float*)reg)[a * 4 + k];
        o = is_half ? h2f(((const uint16_t*)obj)[a]) : ((const float*)obj)[a];
        float s = (float)ls[l];
        float cx = (t[0] + (float)x) * s, cy = (t[1] + (float)y) * s;
        float hw = expf(t[2]) * (s / 2), hh = expf(t[3]) * (s / 2);
        boxes[a * 4 + 0] = cx - hw; boxes[a * 4 + 1] = cy - hh;
        boxes[a * 4 + 2] = cx + hw; boxes[a * 4 + 3] = cy + hh;
        float oc = sigmoidf_(o);
        obj_conf[a] = oc;
        for (int c = 0; c < C; ++c) {
          float v = is_half ? h2f(((const uint16_t*)cls)[(size_t)a * C + c]) : ((const float*)cls)[(size_t)a * C + c];
          cls_conf[(size_t)a * C + c] = sigmoidf_(v) * oc;
        }
      }
}

/* ---- decode, yolox flavour (fp32): pred [A,5+C] = [tx,ty,tw,th,sig(obj),sig(cls)] in place -> cxcywh */
void yxref_decode_yolox(float* pred, int A, int C, int n_levels, const int* lh, const int* lw, const int* ls) {
  int a = 0, D = 5 + C;
  for (int l = 0; l < n_levels; ++l)
    for (int y = 0; y < lh[l]; ++y)
      for (int x = 0; x < lw[l]; ++x, ++a) {
        float s = (float)ls[l];
        float* p = pred + (size_t)a * D;
        p[0] = (p[0] + (float)x) * s; p[1] = (p[1] + (float)y) * s;
        p[2] = expf(p[2]) * s; p[3] = expf(p[3]) * s;
      }
}

/* ---- stable descending argsort ------------------------------------------------------- */
typedef struct { float s; int i; } si_t;
static int cmp_desc(const void* a, const void* b) {
  const si_t* x = (const si_t*)a; const si_t* y = (const si_t*)b;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i); /* ties: lower index first (stable) */
}
void yxref_argsort_desc(const float* s, int n, int* order) {
  si_t* t = (si_t*)malloc(sizeof(si_t) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) { t[i].s = s[i]; t[i].i = i; }
  qsort(t, (size_t)n, sizeof(si_t), cmp_desc);
  for (int i = 0; i < n; ++i) order[i] = t[i].i;
  free(t);
}

/* ---- torchvision nms (CPU kernel semantics). keep: indices into boxes, score-descending. */
int yxref_nms(const float* boxes, const float* scores, int n, float thr, int* keep) {
  if (n == 0) return 0;
  int* order = (int*)malloc(sizeof(int) * (size_t)n);
  unsigned char* sup = (unsigned char*)calloc((size_t)n, 1);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  yxref_argsort_desc(scores, n, order);
  for (int i = 0; i < n; ++i)
    area[i] = (boxes[i * 4 + 2] - boxes[i * 4 + 0]) * (boxes[i * 4 + 3] - boxes[i * 4 + 1]);
  int nk = 0;
  for (int _i = 0; _i < n; ++_i) {
    int i = order[_i];
    if (sup[i]) continue;
    keep[nk++] = i;
    float ix1 = boxes[i * 4], iy1 = boxes[i * 4 + 1], ix2 = boxes[i * 4 + 2], iy2 = boxes[i * 4 + 3], ia = area[i];
    for (int _j = _i + 1; _j < n; ++_j) {
      int j = order[_j];
      if (sup[j]) continue;
      float xx1 = fmaxf(ix1, boxes[j * 4]), yy1 = fmaxf(iy1, boxes[j * 4 + 1]);
      float xx2 = fminf(ix2, boxes[j * 4 + 2]), yy2 = fminf(iy2, boxes[j * 4 + 3]);
      float w = fmaxf(0.0f, xx2 - xx1), h = fmaxf(0.0f, yy2 - yy1);
      float inter = w * h;
      float ovr = inter / (ia + area[j] - inter);
      if (ovr > thr) sup[j] = 1;
    }
  }
  free(order); free(sup); free(area);
  return nk;
}

/* ---- batched_nms, coordinate trick: offsets = label * (max(boxes)+1), all fp32 ------------ */
int yxref_batched_nms_trick(const float* boxes, const float* scores, const float* labels, int n,
                            float thr, int* keep) {
  if (n == 0) return 0;
  float mx = boxes[0];
  for (int i = 1; i < n * 4; ++i) if (boxes[i] > mx) mx = boxes[i];
  float m1 = mx + 1.0f;
  float* nb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    float off = labels[i] * m1;
    for (int k = 0; k < 4; ++k) nb[i * 4 + k] = boxes[i * 4 + k] + off;
  }
  int nk = yxref_nms(nb, scores, n, thr, keep);
  free(nb);
  return nk;
}

/* ---- batched_nms, vanilla: nms per class (ascending class id), then keep sorted by score desc.
 * torchvision: keep_indices = where(keep_mask); return keep_indices[sort(scores[keep_indices], desc)]
 * -> ties ordered by ascending index. */
int yxref_batched_nms_vanilla(const float* boxes, const float* scores, const float* labels, int n,
                              float thr, int* keep) {
  if (n == 0) return 0;
  unsigned char* mask = (unsigned char*)calloc((size_t)n, 1);
  unsigned char* done = (unsigned char*)calloc((size_t)n, 1);
  int* idx = (int*)malloc(sizeof(int) * (size_t)n);
  int* k2 = (int*)malloc(sizeof(int) * (size_t)n);
  float* b2 = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  float* s2 = (float*)malloc(sizeof(float) * (size_t)n);
  for (int i = 0; i < n; ++i) {
    if (done[i]) continue;
    float lab = labels[i];
    int m = 0;
    for (int j = i; j < n; ++j)
      if (labels[j] == lab) { done[j] = 1; idx[m] = j; memcpy(b2 + m * 4, boxes + j * 4, 16); s2[m] = scores[j]; ++m; }
    int nk = yxref_nms(b2, s2, m, thr, k2);
    for (int t = 0; t < nk; ++t) mask[idx[k2[t]]] = 1;
  }
  int m = 0;
  for (int i = 0; i < n; ++i) if (mask[i]) { idx[m] = i; s2[m] = scores[i]; ++m; }
  yxref_argsort_desc(s2, m, k2);
  for (int t = 0; t < m; ++t) keep[t] = idx[k2[t]];
  free(mask); free(done); free(idx); free(k2); free(b2); free(s2);
  return m;
}

/* ---- yolox_nms_torch_batch, one image (postprocess_utils.py:73-127) -----------------------------
 * boxes [A,4], obj_conf [A], cls_conf [A,C].  mode: 0 = coordinate trick, 1 = vanilla, 2 = class-agnostic.
 * cand: how candidates are formed --
 *   0  default      (:86-89)  one per anchor: first-max class, kept when max >= conf_thr
 *   1  multi_class  (:90-95)  one per (anchor, class) with cls_conf >= conf_thr, in nonzero() (row-major) order
 *   2  rmmop        (:74-84)  one per anchor: top-1 class, kept when top1 >= top2 * r1 and obj^2 >= top1 * r2
 *                             (no conf threshold; descending sort ties resolved to the lower class index)
 * det_out [max_det,7] = [x1,y1,x2,y2,obj,score,label]; anchor_out = source anchor per row.
 * max_nms <= 0 disables the top-k cap.  Returns number of detections. */
int yxref_nms_image_main_ex(const float* boxes, const float* obj_conf, const float* cls_conf, int A, int C,
                            float conf_thr, float nms_thr, int max_nms, int max_det, int mode, int cand,
                            float r1, float r2, float* det_out, int* anchor_out) {
  const size_t cap = cand == 1 ? (size_t)A * (size_t)C : (size_t)A;
  float* cb = (float*)malloc(sizeof(float) * 4 * (cap ? cap : 1));
  float* cs = (float*)malloc(sizeof(float) * (cap ? cap : 1));
  float* cl = (float*)malloc(sizeof(float) * (cap ? cap : 1));
  float* co = (float*)malloc(sizeof(float) * (cap ? cap : 1));
  int* ca = (int*)malloc(sizeof(int) * (cap ? cap : 1));
  int n = 0;
  for (int a = 0; a < A; ++a) {
    const float* c = cls_conf + (size_t)a * C;
    if (cand == 1) {
      for (int k = 0; k < C; ++k)
        if (c[k] >= conf_thr) {
          memcpy(cb + (size_t)n * 4, boxes + a * 4, 16); cs[n] = c[k]; cl[n] = (float)k; co[n] = obj_conf[a]; ca[n] = a; ++n;
        }
      continue;
    }
    float best = c[0]; int bi = 0;
    for (int k = 1; k < C; ++k) if (c[k] > best) { best = c[k]; bi = k; }  /* first max wins */
    int pass;
    if (cand == 2) {
      float second = -INFINITY;
      for (int k = 0; k < C; ++k) if (k != bi && c[k] > second) second = c[k];
      const float o = obj_conf[a];
      pass = (best >= second * r1) && (o * o >= best * r2);
    } else {
      pass = best >= conf_thr;
    }
    if (pass) {
      memcpy(cb + (size_t)n * 4, boxes + a * 4, 16); cs[n] = best; cl[n] = (float)bi; co[n] = obj_conf[a]; ca[n] = a; ++n;
    }
  }
  if (max_nms > 0 && n > max_nms) { /* argsort(desc)[:max_nms], detections reordered by score */
    int* ord = (int*)malloc(sizeof(int) * (size_t)n);
    yxref_argsort_desc(cs, n, ord);
    float* b2 = (float*)malloc(sizeof(float) * 4 * (size_t)max_nms);
    float* s2 = (float*)malloc(sizeof(float) * (size_t)max_nms);
    float* l2 = (float*)malloc(sizeof(float) * (size_t)max_nms);
    float* o2 = (float*)malloc(sizeof(float) * (size_t)max_nms);
    int* a2 = (int*)malloc(sizeof(int) * (size_t)max_nms);
    for (int t = 0; t < max_nms; ++t) {
      int j = ord[t];
      memcpy(b2 + t * 4, cb + (size_t)j * 4, 16); s2[t] = cs[j]; l2[t] = cl[j]; o2[t] = co[j]; a2[t] = ca[j];
    }
    free(cb); free(cs); free(cl); free(co); free(ca); free(ord);
    cb = b2; cs = s2; cl = l2; co = o2; ca = a2; n = max_nms;
  }
  int* keep = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int nk = mode == 2 ? yxref_nms(cb, cs, n, nms_thr, keep)
         : mode == 1 ? yxref_batched_nms_vanilla(cb, cs, cl, n, nms_thr, keep)
                     : yxref_batched_nms_trick(cb, cs, cl, n, nms_thr, keep);
  if (nk > max_det) nk = max_det;
  for (int t = 0; t < nk; ++t) {
    int j = keep[t];
    memcpy(det_out + t * 7, cb + (size_t)j * 4, 16);
    det_out[t * 7 + 4] = co[j]; det_out[t * 7 + 5] = cs[j]; det_out[t * 7 + 6] = cl[j];
    anchor_out[t] = ca[j];
  }
  free(cb); free(cs); free(cl); free(co); free(ca); free(keep);
  return nk;
}

int yxref_nms_image_main(const float* boxes, const float* obj_conf, const float* cls_conf, int A, int C,
                         float conf_thr, float nms_thr, int max_nms, int max_det, int mode,
                         float* det_out, int* anchor_out) {
  return yxref_nms_image_main_ex(boxes, obj_conf, cls_conf, A, C, conf_thr, nms_thr, max_nms, max_det, mode, 0, 0.0f,
                                 0.0f, det_out, anchor_out);
}

/* ---- yolox.utils.postprocess, one image (boxes.py:38-75), fp32 ----------------------------
 * pred [A,5+C] = [cx,cy,w,h,obj,cls...] ; converted to xyxy IN PLACE like the reference (:38-43).
 * det_out [A,7] = [x1,y1,x2,y2,obj,class_conf,class_pred]; no caps. */
int yxref_postprocess_image(float* pred, int A, int C, float conf_thr, float nms_thr, int mode,
                            float* det_out, int* anchor_out) {
  int D = 5 + C;
  float* cb = (float*)malloc(sizeof(float) * 4 * (size_t)A);
  float* cs = (float*)malloc(sizeof(float) * (size_t)A);
  float* cl = (float*)malloc(sizeof(float) * (size_t)A);
  float* cc = (float*)malloc(sizeof(float) * (size_t)A);
  float* co = (float*)malloc(sizeof(float) * (size_t)A);
  int* ca = (int*)malloc(sizeof(int) * (size_t)A);
  int n = 0;
  for (int a = 0; a < A; ++a) {
    float* p = pred + (size_t)a * D;
    float x1 = p[0] - p[2] / 2, y1 = p[1] - p[3] / 2, x2 = p[0] + p[2] / 2, y2 = p[1] + p[3] / 2;
    p[0] = x1; p[1] = y1; p[2] = x2; p[3] = y2;
    float best = p[5]; int bi = 0;
    for (int k = 1; k < C; ++k) if (p[5 + k] > best) { best = p[5 + k]; bi = k; }
    float score = p[4] * best;
    if (score >= conf_thr) {
      memcpy(cb + n * 4, p, 16); cs[n] = score; cl[n] = (float)bi; cc[n] = best; co[n] = p[4]; ca[n] = a; ++n;
    }
  }
  int* keep = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int nk = mode == 2 ? yxref_nms(cb, cs, n, nms_thr, keep)
         : mode == 1 ? yxref_batched_nms_vanilla(cb, cs, cl, n, nms_thr, keep)
                     : yxref_batched_nms_trick(cb, cs, cl, n, nms_thr, keep);
  for (int t = 0; t < nk; ++t) {
    int j = keep[t];
    memcpy(det_out + t * 7, cb + j * 4, 16);
    det_out[t * 7 + 4] = co[j]; det_out[t * 7 + 5] = cc[j]; det_out[t * 7 + 6] = cl[j];
    anchor_out[t] = ca[j];
  }
  free(cb); free(cs); free(cl); free(cc); free(co); free(ca); free(keep);
  return nk;
}
