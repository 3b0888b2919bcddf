// COCO bounding-box evaluation (host code): per-(image, category) greedy matching at ten IoU thresholds, accumulation into
// the 101-point interpolated precision / recall tables and the twelve summary numbers.  This is the computation the
// reference delegates to pycocotools' COCOeval(iouType="bbox") with its C++ accelerator
// (yolox/evaluators/coco_evaluator.py:198-215, yolox/layers/fast_coco_eval_api.py:19-147, yolox/layers/csrc/cocoeval/),
// restated from the published COCO evaluation protocol.  It runs after the hot path, on the gathered detections of a
// whole dataset; it is CPU work in the reference too (one pass over <= 100 detections per image and category).
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <vector>

#include "yx_internal.h"

namespace {

constexpr int kT = 10, kR = 101, kA = 4, kM = 3;
const int kMaxDets[kM] = {1, 10, 100};
const double kArea[kA][2] = {{0.0, 1e10}, {0.0, 1024.0}, {1024.0, 9216.0}, {9216.0, 1e10}};

struct Box { double x, y, w, h; };

inline double box_iou(const Box& d, const Box& g, bool crowd) {
  const double w = std::min(d.x + d.w, g.x + g.w) - std::max(d.x, g.x);
  if (w <= 0) return 0.0;
  const double h = std::min(d.y + d.h, g.y + g.h) - std::max(d.y, g.y);
  if (h <= 0) return 0.0;
  const double i = w * h;
  const double u = crowd ? d.w * d.h : d.w * d.h + g.w * g.h - i;
  return i / u;
}

}  // namespace

using namespace yx;

extern "C" int yx_cocoeval_bbox(const int64_t* gt_image, const int32_t* gt_category, const double* gt_bbox,
                                const double* gt_area, const int32_t* gt_iscrowd, int64_t n_gt, const int64_t* dt_image,
                                const int32_t* dt_category, const double* dt_bbox, const double* dt_score, int64_t n_dt,
                                const int64_t* image_ids, int64_t n_images, const int32_t* category_ids, int32_t n_categories,
                                double* stats12, double* precision_out, double* recall_out) {
  YX_REQUIRE(n_gt >= 0 && n_dt >= 0 && n_images >= 0 && n_categories >= 0 && stats12 != nullptr, "cocoeval: bad sizes");
  YX_REQUIRE((n_gt == 0 || (gt_image && gt_category && gt_bbox && gt_area && gt_iscrowd)) &&
                 (n_dt == 0 || (dt_image && dt_category && dt_bbox && dt_score)) && (n_images == 0 || image_ids) &&
                 (n_categories == 0 || category_ids),
             "cocoeval: null array");
  // evaluated images / categories: unique, ascending (COCOeval.evaluate: np.unique of params.imgIds / catIds)
  std::vector<int64_t> imgs(image_ids, image_ids + n_images);
  std::sort(imgs.begin(), imgs.end());
  imgs.erase(std::unique(imgs.begin(), imgs.end()), imgs.end());
  std::vector<int32_t> cats(category_ids, category_ids + n_categories);
  std::sort(cats.begin(), cats.end());
  cats.erase(std::unique(cats.begin(), cats.end()), cats.end());
  const int K = (int)cats.size();
  const int64_t I = (int64_t)imgs.size();
  std::map<int64_t, int64_t> img_index;
  for (int64_t i = 0; i < I; ++i) img_index[imgs[i]] = i;
  std::map<int32_t, int> cat_index;
  for (int k = 0; k < K; ++k) cat_index[cats[k]] = k;

  // bucket annotations by (category, image), keeping their input order (the sorts below are stable, as mergesort is)
  std::vector<std::vector<int64_t>> gts((size_t)K * I), dts((size_t)K * I);
  for (int64_t g = 0; g < n_gt; ++g) {
    auto ii = img_index.find(gt_image[g]);
    auto kk = cat_index.find(gt_category[g]);
    if (ii != img_index.end() && kk != cat_index.end()) gts[(size_t)kk->second * I + ii->second].push_back(g);
  }
  for (int64_t d = 0; d < n_dt; ++d) {
    auto ii = img_index.find(dt_image[d]);
    auto kk = cat_index.find(dt_category[d]);
    if (ii != img_index.end() && kk != cat_index.end()) dts[(size_t)kk->second * I + ii->second].push_back(d);
  }

  double iou_thr[kT], rec_thr[kR];
  for (int t = 0; t < kT; ++t) iou_thr[t] = 0.5 + t * ((0.95 - 0.5) / 9.0);  // np.linspace(.5, .95, 10)
  iou_thr[kT - 1] = 0.95;
  for (int r = 0; r < kR; ++r) rec_thr[r] = 0.01 * r;

  std::vector<double> precision((size_t)kT * kR * K * kA * kM, -1.0), recall((size_t)kT * K * kA * kM, -1.0);
  auto P = [&](int t, int r, int k, int a, int m) -> double& { return precision[((((size_t)t * kR + r) * K + k) * kA + a) * kM + m]; };
  auto Rc = [&](int t, int k, int a, int m) -> double& { return recall[(((size_t)t * K + k) * kA + a) * kM + m]; };

  std::vector<double> ious;
  for (int k = 0; k < K; ++k) {
    // The greedy matching visits detections in score order, so the result for a cap of 1 / 10 detections is a prefix of
    // the result for 100; each (area range, cap) cell is simply evaluated on its own.
    for (int a = 0; a < kA; ++a)
      for (int m = 0; m < kM; ++m) {
        std::vector<double> all_score;
        std::vector<std::vector<uint8_t>> all_match(kT), all_ignore(kT);
        int64_t npig = 0;
        bool any_cell = false;
        for (int64_t i = 0; i < I; ++i) {
          const std::vector<int64_t>& G = gts[(size_t)k * I + i];
          const std::vector<int64_t>& D = dts[(size_t)k * I + i];
          if (G.empty() && D.empty()) continue;  // evaluateImg returns None
          any_cell = true;
          // detections: score-descending (stable), at most maxDets[-1] enter computeIoU, at most maxDets[m] are matched
          std::vector<int64_t> dord(D);
          std::stable_sort(dord.begin(), dord.end(), [&](int64_t x, int64_t y) { return dt_score[x] > dt_score[y]; });
          if ((int)dord.size() > kMaxDets[kM - 1]) dord.resize(kMaxDets[kM - 1]);
          const int nd = std::min<int>((int)dord.size(), kMaxDets[m]);
          // ground truths: ignored ones (crowd, or area outside the range) last (stable)
          std::vector<int64_t> gord(G);
          std::vector<uint8_t> gig(G.size());
          auto ignored_gt = [&](int64_t g) { return gt_iscrowd[g] != 0 || gt_area[g] < kArea[a][0] || gt_area[g] > kArea[a][1]; };
          std::stable_sort(gord.begin(), gord.end(), [&](int64_t x, int64_t y) { return (int)ignored_gt(x) < (int)ignored_gt(y); });
          const int ng = (int)gord.size();
          for (int g = 0; g < ng; ++g) {
            gig[g] = ignored_gt(gord[g]) ? 1 : 0;
            if (!gig[g]) ++npig;
          }
          ious.assign((size_t)nd * std::max(ng, 1), 0.0);
          for (int d = 0; d < nd; ++d) {
            const double* db = dt_bbox + 4 * dord[d];
            const Box bd{db[0], db[1], db[2], db[3]};
            for (int g = 0; g < ng; ++g) {
              const double* gb = gt_bbox + 4 * gord[g];
              ious[(size_t)d * ng + g] = box_iou(bd, Box{gb[0], gb[1], gb[2], gb[3]}, gt_iscrowd[gord[g]] != 0);
            }
          }
          for (int d = 0; d < nd; ++d) all_score.push_back(dt_score[dord[d]]);
          std::vector<uint8_t> gtm(ng);
          for (int t = 0; t < kT; ++t) {
            std::fill(gtm.begin(), gtm.end(), 0);
            for (int d = 0; d < nd; ++d) {
              double best = std::min(iou_thr[t], 1.0 - 1e-10);
              int mi = -1;
              for (int g = 0; g < ng; ++g) {
                if (gtm[g] && gt_iscrowd[gord[g]] == 0) continue;       // already matched (crowds may match many)
                if (mi > -1 && gig[mi] == 0 && gig[g] == 1) break;       // a counted match beats any ignored gt
                if (ious[(size_t)d * ng + g] < best) continue;
                best = ious[(size_t)d * ng + g];
                mi = g;
              }
              uint8_t is_match = 0, is_ignored = 0;
              if (mi >= 0) {
                is_match = 1;
                is_ignored = gig[mi];
                gtm[mi] = 1;
              } else {
                const double* db = dt_bbox + 4 * dord[d];
                const double area = db[2] * db[3];                       // loadRes: area = w * h for bbox results
                is_ignored = (area < kArea[a][0] || area > kArea[a][1]) ? 1 : 0;
              }
              all_match[t].push_back(is_match);
              all_ignore[t].push_back(is_ignored);
            }
          }
        }
        if (!any_cell || npig == 0) continue;  // accumulate: "if npig == 0: continue" leaves -1
        const int64_t n = (int64_t)all_score.size();
        std::vector<int64_t> ord(n);
        std::iota(ord.begin(), ord.end(), 0);
        std::stable_sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) { return all_score[x] > all_score[y]; });
        std::vector<double> rc(n), pr(n);
        for (int t = 0; t < kT; ++t) {
          double tp = 0, fp = 0;
          for (int64_t j = 0; j < n; ++j) {
            const int64_t o = ord[j];
            if (!all_ignore[t][o]) {
              if (all_match[t][o]) tp += 1; else fp += 1;
            }
            rc[j] = tp / (double)npig;
            // the reference's first choice is its C++ accelerator (coco_evaluator.py:204-205, cocoeval.cpp:332-335):
            // precision = tp / (tp + fp) exactly, 0 while nothing counts; pycocotools' Python fallback divides by
            // fp + tp + np.spacing(1) instead, one ulp away on some entries (oracle/cocoeval_ref.py keeps both forms)
            pr[j] = (tp + fp) > 0 ? tp / (tp + fp) : 0.0;
          }
          Rc(t, k, a, m) = n ? rc[n - 1] : 0.0;
          for (int64_t j = n - 1; j > 0; --j)
            if (pr[j] > pr[j - 1]) pr[j - 1] = pr[j];
          for (int r = 0; r < kR; ++r) {  // np.searchsorted(rc, recThrs, side="left")
            const int64_t pi = std::lower_bound(rc.begin(), rc.end(), rec_thr[r]) - rc.begin();
            P(t, r, k, a, m) = pi < n ? pr[pi] : 0.0;
          }
        }
      }
  }

  // ---- summarize (COCOeval.summarize, the twelve detection numbers) --------------------------------------
  auto mean_p = [&](int t_lo, int t_hi, int a, int m) {
    double s = 0;
    int64_t c = 0;
    for (int t = t_lo; t < t_hi; ++t)
      for (int r = 0; r < kR; ++r)
        for (int k = 0; k < K; ++k) {
          const double v = P(t, r, k, a, m);
          if (v > -1) { s += v; ++c; }
        }
    return c ? s / (double)c : -1.0;
  };
  auto mean_r = [&](int a, int m) {
    double s = 0;
    int64_t c = 0;
    for (int t = 0; t < kT; ++t)
      for (int k = 0; k < K; ++k) {
        const double v = Rc(t, k, a, m);
        if (v > -1) { s += v; ++c; }
      }
    return c ? s / (double)c : -1.0;
  };
  stats12[0] = mean_p(0, kT, 0, 2);
  stats12[1] = mean_p(0, 1, 0, 2);   // IoU 0.50
  stats12[2] = mean_p(5, 6, 0, 2);   // IoU 0.75
  stats12[3] = mean_p(0, kT, 1, 2);
  stats12[4] = mean_p(0, kT, 2, 2);
  stats12[5] = mean_p(0, kT, 3, 2);
  stats12[6] = mean_r(0, 0);
  stats12[7] = mean_r(0, 1);
  stats12[8] = mean_r(0, 2);
  stats12[9] = mean_r(1, 2);
  stats12[10] = mean_r(2, 2);
  stats12[11] = mean_r(3, 2);
  if (precision_out) std::copy(precision.begin(), precision.end(), precision_out);
  if (recall_out) std::copy(recall.begin(), recall.end(), recall_out);
  return YX_OK;
}
