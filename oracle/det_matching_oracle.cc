// ORACLE (test infrastructure, NOT product code): CPU restatement of the
// reference's DetectionMatching op, nms_net/matching_module/det_matching.cc.
//
//   ordering   det_matching.cc:44-51,95-98  (std::sort on indices ascending by
//              score, then std::reverse; GTs sorted ascending by the ignore flag)
//   outputs    det_matching.cc:101-117      (labels 0, weights 1, assignment -1)
//   greedy     det_matching.cc:119-159
//
// The orderings deliberately use the same libstdc++ std::sort / std::reverse
// calls as the reference so that ties are broken the way a build of the
// reference with this toolchain would break them.  Plain C ABI, called from
// tests through ctypes.  Build: `make -C oracle`.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <numeric>
#include <vector>

namespace {

template <typename V>
std::vector<size_t> ascending_order(const V* v, size_t n) {
  std::vector<size_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(),
            [v](size_t a, size_t b) { return v[a] < v[b]; });
  return idx;
}

}  // namespace

extern "C" {

// iou: [n_dets, n_gt] row-major; score: [n_dets]; ignore: [n_gt] (0/1 bytes).
// labels/weights: [n_dets] float; assignment: [n_dets] int32.
// det_order_out (optional, may be null): the visiting order, for tests.
int oracle_detection_matching(const float* iou, const float* score,
                              const uint8_t* ignore, int n_dets, int n_gt,
                              float* labels, float* weights,
                              int32_t* assignment, int64_t* det_order_out) {
  if (n_dets < 0 || n_gt < 0) return 1;
  const float iou_thresh = 0.5f;  // det_matching.cc:73 (cfg value is unused)

  std::vector<size_t> det_order = ascending_order(score, (size_t)n_dets);
  std::reverse(det_order.begin(), det_order.end());
  std::vector<bool> ign(n_gt);
  for (int g = 0; g < n_gt; ++g) ign[g] = ignore[g] != 0;
  std::vector<size_t> gt_order(n_gt);
  std::iota(gt_order.begin(), gt_order.end(), 0);
  std::sort(gt_order.begin(), gt_order.end(),
            [&ign](size_t a, size_t b) { return ign[a] < ign[b]; });

  for (int d = 0; d < n_dets; ++d) {
    labels[d] = 0.f;
    weights[d] = 1.f;
    assignment[d] = -1;
    if (det_order_out) det_order_out[d] = (int64_t)det_order[d];
  }

  std::vector<bool> taken(n_gt, false);
  for (int rank = 0; rank < n_dets; ++rank) {
    const size_t det = det_order[rank];
    const float* row = iou + det * (size_t)n_gt;
    float best = iou_thresh;
    int match = -1;
    for (int k = 0; k < n_gt; ++k) {
      const size_t gt = gt_order[k];
      if (taken[gt] && !ign[gt]) continue;      // :134 regular GT already used
      if (match > -1 && ign[gt]) break;         // :138 matched, reached crowd GTs
      if (row[gt] < best) continue;             // :142 not better (NaN passes)
      best = row[gt];                           // :147-148 ">=": later GT wins ties
      match = (int)gt;
    }
    if (match > -1) {                           // :151-158
      taken[match] = true;
      labels[det] = 1.f;
      assignment[det] = match;
      if (ign[match]) weights[det] = 0.f;
    }
  }
  return 0;
}

}  // extern "C"
