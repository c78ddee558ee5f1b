// The output stage of SiftFeatureMatcher::Match (src/feature/matching.cc:749-839) as data: turns the
// batched results of b2_match_pairs / b2_verify_pairs / b2_verify_relative_pose into the exact rows the
// reference writes to its `matches` and `two_view_geometries` tables, so that a caller can insert a whole
// batch in ONE transaction instead of one mutex-guarded write per pair.  SQLite itself
// stays with the caller (the storage engine is out of scope); this header fixes the WIRE FORMAT:
//
//   Database::ImagePairToPairId / SwapImagePair      src/base/database.h:337-365  (kMaxNumImages = 2^31 - 1)
//   Database::WriteMatches                           src/base/database.cc:680-697 (columns swapped if id1 > id2)
//   Database::WriteTwoViewGeometry                   src/base/database.cc:699-755 (geometry inverted if id1 > id2;
//        DAGSfM stores qvec in the column named F and tvec in the column named E, never binds H)
//   FeatureMatchesToBlob                             src/base/database.cc:90-98   (row-major uint32 N x 2)
//   schema / statements                              src/base/database.cc:1233-1261, 1121-1130
//   the two min_num_inliers gates                    src/feature/matching.cc:824-831
//   self-match and duplicate filtering               src/feature/matching.cc:766-781
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_set>
#include <utility>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {

constexpr int64_t kMaxNumImages = 2147483647;  // std::numeric_limits<int32_t>::max()

inline bool SwapImagePair(uint32_t image_id1, uint32_t image_id2) { return image_id1 > image_id2; }
inline int64_t ImagePairToPairId(uint32_t image_id1, uint32_t image_id2) {
  return SwapImagePair(image_id1, image_id2) ? kMaxNumImages * image_id2 + image_id1 : kMaxNumImages * image_id1 + image_id2;
}
inline void PairIdToImagePair(int64_t pair_id, uint32_t* image_id1, uint32_t* image_id2) {
  *image_id2 = static_cast<uint32_t>(pair_id % kMaxNumImages);
  *image_id1 = static_cast<uint32_t>((pair_id - *image_id2) / kMaxNumImages);
}

// matching.cc:766-781: drops self-matches and repeated pairs (in either order), keeps first occurrences in order.
inline std::vector<std::pair<uint32_t, uint32_t>> UniqueImagePairs(const std::vector<std::pair<uint32_t, uint32_t>>& image_pairs) {
  std::vector<std::pair<uint32_t, uint32_t>> out;
  std::unordered_set<int64_t> seen;
  seen.reserve(image_pairs.size());
  for (const auto& p : image_pairs) {
    if (p.first == p.second) continue;
    if (!seen.insert(ImagePairToPairId(p.first, p.second)).second) continue;
    out.push_back(p);
  }
  return out;
}

// INSERT INTO matches(pair_id, rows, cols, data) VALUES(?, ?, ?, ?);
struct MatchesRow {
  int64_t pair_id = 0, rows = 0, cols = 2;
  std::vector<uint8_t> data;  // rows x 2 uint32, row-major
};
// INSERT INTO two_view_geometries(pair_id, rows, cols, data, config, F, E, H) VALUES(?, ?, ?, ?, ?, ?, ?, ?);  (H stays NULL)
struct TwoViewGeometryRow {
  int64_t pair_id = 0, rows = 0, cols = 2;
  std::vector<uint8_t> data;  // inlier matches, rows x 2 uint32, row-major
  int64_t config = 0;
  std::vector<uint8_t> F;     // qvec: 4 doubles (w, x, y, z); empty when there are no inliers
  std::vector<uint8_t> E;     // tvec: 3 doubles; empty when there are no inliers
};

namespace detail {
inline std::vector<uint8_t> MatchesBlob(const uint32_t* m, int64_t n, bool swap) {
  std::vector<uint8_t> b((size_t)n * 8);
  uint32_t* o = reinterpret_cast<uint32_t*>(b.data());
  for (int64_t i = 0; i < n; ++i) {
    o[2 * i] = m[2 * i + (swap ? 1 : 0)];
    o[2 * i + 1] = m[2 * i + (swap ? 0 : 1)];
  }
  return b;
}
// InvertPose (base/pose.cc:192-196) as TwoViewGeometry::Invert applies it (two_view_geometry.cc:100-102)
inline void InvertPose(const double* q, const double* t, double* qi, double* ti) {
  qi[0] = q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = -q[3];
  double n[4] = {qi[0], qi[1], qi[2], qi[3]};
  const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2] + n[3] * n[3]);
  if (nn == 0) n[0] = 1.0;  // NormalizeQuaternion: a zero quaternion becomes (1, x, y, z)
  else for (double& x : n) x /= nn;
  const double w = n[0], x = n[1], y = n[2], z = n[3];
  const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                          {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                          {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
  for (int r = 0; r < 3; ++r) ti[r] = -(R[r][0] * t[0] + R[r][1] * t[1] + R[r][2] * t[2]);
}
}  // namespace detail

// One pair.  matches: the n_matches x 2 raw matches of b2_match_pairs; result / inlier_matches: the pair's
// b2_verify_pairs output; pose: its b2_verify_relative_pose output or NULL (qvec = tvec = 0, config of `result`).
// Applies matching.cc:824-831: fewer than min_num_inliers raw matches -> an empty matches row; fewer than
// min_num_inliers inlier matches -> the row of a default-constructed TwoViewGeometry (config UNDEFINED).
inline void MakeRows(uint32_t image_id1, uint32_t image_id2, const uint32_t* matches, int64_t n_matches,
                     const b2_two_view_result& result, const uint32_t* inlier_matches, const b2_relative_pose* pose,
                     int min_num_inliers, MatchesRow* mrow, TwoViewGeometryRow* grow) {
  const bool swap = SwapImagePair(image_id1, image_id2);
  const int64_t pair_id = ImagePairToPairId(image_id1, image_id2);
  if (n_matches < min_num_inliers) n_matches = 0;
  mrow->pair_id = pair_id;
  mrow->rows = n_matches;
  mrow->cols = 2;
  mrow->data = detail::MatchesBlob(matches, n_matches, swap);
  *grow = TwoViewGeometryRow();
  grow->pair_id = pair_id;
  const int64_t n_inl = result.n_inliers;
  if (n_inl < min_num_inliers || n_inl <= 0) return;  // TwoViewGeometry(): no inliers, config 0, empty F / E blobs
  grow->rows = n_inl;
  grow->data = detail::MatchesBlob(inlier_matches, n_inl, swap);
  grow->config = pose ? pose->config : result.config;
  double q[4] = {0, 0, 0, 0}, t[3] = {0, 0, 0};
  if (pose) { std::memcpy(q, pose->qvec, sizeof q); std::memcpy(t, pose->tvec, sizeof t); }
  if (swap) {
    double qi[4], ti[3];
    detail::InvertPose(q, t, qi, ti);
    std::memcpy(q, qi, sizeof q);
    std::memcpy(t, ti, sizeof t);
  }
  grow->F.resize(sizeof q);
  std::memcpy(grow->F.data(), q, sizeof q);
  grow->E.resize(sizeof t);
  std::memcpy(grow->E.data(), t, sizeof t);
}

// A whole batch in the layout of the C ABI: image_ids maps the ABI's image indices to database image_ids.
inline void MakeRowsBatch(int64_t n_pairs, const uint32_t* pairs, const uint32_t* image_ids, const int64_t* match_offsets,
                          const uint32_t* matches, const b2_two_view_result* results, const uint32_t* inlier_matches,
                          const b2_relative_pose* poses /* may be NULL */, int min_num_inliers,
                          std::vector<MatchesRow>* mrows, std::vector<TwoViewGeometryRow>* grows) {
  mrows->resize((size_t)n_pairs);
  grows->resize((size_t)n_pairs);
  for (int64_t p = 0; p < n_pairs; ++p) {
    const int64_t a = match_offsets[p], n = match_offsets[p + 1] - a;
    MakeRows(image_ids[pairs[2 * p]], image_ids[pairs[2 * p + 1]], matches + 2 * a, n, results[p], inlier_matches + 2 * a,
             poses ? &poses[p] : nullptr, min_num_inliers, &(*mrows)[(size_t)p], &(*grows)[(size_t)p]);
  }
}

}  // namespace dagsfm_b200
