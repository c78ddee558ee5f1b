// Header-only C++ adaptors that give the C ABI (include/dagsfm_b200.h) the reference's own
// signatures for the matching path, so that the reference's callers compile unchanged:
//
//   SiftMatchGPU                      lib/SiftGPU/SiftGPU.h:276-373   (subset colmap uses)
//   CreateSiftGPUMatcher              src/feature/sift.h:229-230, sift.cc:877-939
//   MatchSiftFeaturesGPU              src/feature/sift.h:235-239, sift.cc:941-985
//
// The reference cannot be built in this environment (Eigen / glog / Boost absent), so the
// adaptors are templates over the few members they touch: a descriptor matrix only needs
// rows(), cols() and data() (row-major uint8, as Eigen::Matrix<uint8_t,Dyn,Dyn,RowMajor>),
// FeatureMatches is any std::vector of {uint32 point2D_idx1, point2D_idx2}
// (src/feature/types.h:86-104).  Drop this header next to the reference's sift.cc, replace its
// `#include "SiftGPU/SiftGPU.h"` and link libdagsfm_b200.so -- see INTEGRATION.md.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {

struct FeatureMatch {  // src/feature/types.h:86-98
  uint32_t point2D_idx1 = 0xffffffffu;
  uint32_t point2D_idx2 = 0xffffffffu;
};
using FeatureMatches = std::vector<FeatureMatch>;
static_assert(sizeof(FeatureMatch) == 8, "FeatureMatch must be two packed uint32");

struct SiftMatchingOptions {  // src/feature/sift.h:116-165 (fields the matcher reads)
  int num_threads = -1;
  bool use_gpu = true;
  std::string gpu_index = "-1";
  double max_ratio = 0.8;
  double max_distance = 0.7;
  bool cross_check = true;
  int max_num_matches = 32768;
  double max_error = 4.0;  // guided matching
  bool Check() const { return max_ratio > 0 && max_distance > 0 && max_num_matches > 0 && max_error > 0; }
};

// The slice of SiftMatchGPU that colmap's sift.cc drives.
class SiftMatchGPU {
 public:
  enum { SIFTMATCH_SAME_AS_SIFTGPU = 0, SIFTMATCH_GLSL = 2, SIFTMATCH_CUDA = 3, SIFTMATCH_CUDA_DEVICE0 = 3 };
  int gpu_index = 0;  // public field used by colmap (sift.cc:932)

  explicit SiftMatchGPU(int max_sift = 4096) : max_sift_(max_sift) {}
  SiftMatchGPU(const SiftMatchGPU&) = delete;
  SiftMatchGPU& operator=(SiftMatchGPU&& o) noexcept {
    std::swap(h_, o.h_); std::swap(max_sift_, o.max_sift_); std::swap(device_, o.device_); gpu_index = o.gpu_index;
    return *this;
  }
  ~SiftMatchGPU() { if (h_) b2_match_destroy(h_); }

  void SetLanguage(int language) { device_ = language >= SIFTMATCH_CUDA_DEVICE0 ? language - SIFTMATCH_CUDA_DEVICE0 : 0; }
  // returns 0 when no usable context exists (colmap then fails CreateSiftGPUMatcher, sift.cc:905-907)
  int VerifyContextGL() {
    if (h_) return 1;
    return b2_match_create(device_, &h_) == B2_OK ? 1 : 0;
  }
  bool Allocate(int max_sift, int /*mbm*/) {
    max_sift_ = (max_sift + 31) / 32 * 32;  // SiftMatchCU.cpp:84-87 rounds up to x32
    return VerifyContextGL() != 0;
  }
  void SetMaxSift(int max_sift) { max_sift_ = max_sift; }
  int GetMaxSift() const { return max_sift_; }
  // SiftMatchCU.cpp:99-112: features beyond max_sift are silently clamped
  void SetDescriptors(int index, int num, const unsigned char* descriptors, int /*id*/ = -1) {
    if (!h_ || index < 0 || index > 1) return;
    // the C ABI reads a NULL pointer as "keep the previous upload"; an explicitly empty set
    // (Eigen's data() of a 0-row matrix may be NULL) must still replace it
    static const unsigned char kEmpty = 0;
    b2_match_set_descriptors(h_, index, std::min(num, max_sift_), descriptors ? descriptors : &kEmpty);
  }
  // returns the number of matches, -1 on a device error (SiftMatchCU.cpp:193-196)
  int GetSiftMatch(int max_match, uint32_t match_buffer[][2], float distmax = 0.7f, float ratiomax = 0.8f,
                   int mutual_best_match = 1) {
    if (!h_) return -1;
    b2_match_options o;
    o.max_ratio = ratiomax; o.max_distance = distmax; o.cross_check = mutual_best_match; o.max_num_matches = max_match;
    int32_t n = 0;
    if (b2_match_run(h_, &o, &match_buffer[0][0], &n) != B2_OK) return -1;
    return n;
  }

  // Guided matching keeps host copies of the two slots: a NULL descriptor / keypoint pointer of
  // MatchGuidedSiftFeaturesGPU means "same image as the previous call" (sift.cc:1009-1035).
  b2_matcher* handle() const { return h_; }
  std::vector<unsigned char> guided_desc[2];
  std::vector<float> guided_xy[2];

 private:
  b2_matcher* h_ = nullptr;
  int max_sift_ = 4096;
  int device_ = 0;
};

// sift.cc:877-939
inline bool CreateSiftGPUMatcher(const SiftMatchingOptions& match_options, SiftMatchGPU* sift_match_gpu) {
  if (!match_options.Check() || !sift_match_gpu) return false;
  int gpu = 0;
  try { gpu = std::stoi(match_options.gpu_index); } catch (...) { gpu = -1; }
  *sift_match_gpu = SiftMatchGPU(match_options.max_num_matches);
  sift_match_gpu->SetLanguage(gpu >= 0 ? SiftMatchGPU::SIFTMATCH_CUDA_DEVICE0 + gpu : SiftMatchGPU::SIFTMATCH_CUDA);
  if (sift_match_gpu->VerifyContextGL() == 0) return false;
  if (!sift_match_gpu->Allocate(match_options.max_num_matches, match_options.cross_check)) {
    std::fprintf(stderr, "ERROR: Not enough GPU memory to match %d features. Reduce the maximum number of matches.\n",
                 match_options.max_num_matches);
    return false;
  }
  sift_match_gpu->gpu_index = std::max(gpu, 0);
  return true;
}

// sift.cc:941-985.  Descriptors: any type with rows(), cols(), data() (row-major uint8).
template <class Descriptors, class Matches>
void MatchSiftFeaturesGPU(const SiftMatchingOptions& match_options, const Descriptors* descriptors1,
                          const Descriptors* descriptors2, SiftMatchGPU* sift_match_gpu, Matches* matches) {
  if (descriptors1 != nullptr) {
    if (sift_match_gpu->GetMaxSift() < (int)descriptors1->rows())
      std::printf("WARNING: Clamping features from %d to %d - consider increasing the maximum number of matches.\n",
                  (int)descriptors1->rows(), sift_match_gpu->GetMaxSift());
    sift_match_gpu->SetDescriptors(0, (int)descriptors1->rows(), descriptors1->data());
  }
  if (descriptors2 != nullptr) {
    if (sift_match_gpu->GetMaxSift() < (int)descriptors2->rows())
      std::printf("WARNING: Clamping features from %d to %d - consider increasing the maximum number of matches.\n",
                  (int)descriptors2->rows(), sift_match_gpu->GetMaxSift());
    sift_match_gpu->SetDescriptors(1, (int)descriptors2->rows(), descriptors2->data());
  }
  matches->resize(static_cast<size_t>(match_options.max_num_matches));
  const int num_matches = sift_match_gpu->GetSiftMatch(
      match_options.max_num_matches, reinterpret_cast<uint32_t(*)[2]>(matches->data()),
      static_cast<float>(match_options.max_distance), static_cast<float>(match_options.max_ratio),
      match_options.cross_check);
  if (num_matches < 0) {
    std::fprintf(stderr, "ERROR: Feature matching failed. This is probably caused by insufficient GPU memory. "
                         "Consider reducing the maximum number of features and/or matches.\n");
    matches->clear();
  } else {
    matches->resize(num_matches);
  }
}

// sift.cc:987-1066.  Keypoints: any container of structs with float members x, y (FeatureKeypoint);
// Geometry: any struct with `config`, `F`, `H` (3x3, indexable as M(r, c)) and `inlier_matches`
// (TwoViewGeometry, two_view_geometry.h:43-113).  Configurations without a guided filter leave
// inlier_matches untouched, as the reference does (:1049-1051).
template <class Keypoints, class Descriptors, class Geometry>
void MatchGuidedSiftFeaturesGPU(const SiftMatchingOptions& match_options, const Keypoints* keypoints1,
                                const Keypoints* keypoints2, const Descriptors* descriptors1,
                                const Descriptors* descriptors2, SiftMatchGPU* sift_match_gpu,
                                Geometry* two_view_geometry) {
  auto stage = [&](int slot, const Keypoints* kp, const Descriptors* d) {
    if (d == nullptr) return;
    const size_t n = std::min<size_t>((size_t)d->rows(), (size_t)sift_match_gpu->GetMaxSift());
    sift_match_gpu->guided_desc[slot].assign(d->data(), d->data() + n * 128);
    sift_match_gpu->guided_xy[slot].resize(2 * n);
    for (size_t i = 0; i < n; ++i) {
      sift_match_gpu->guided_xy[slot][2 * i] = (*kp)[i].x;
      sift_match_gpu->guided_xy[slot][2 * i + 1] = (*kp)[i].y;
    }
  };
  stage(0, keypoints1, descriptors1);
  stage(1, keypoints2, descriptors2);
  const int cfg = (int)two_view_geometry->config;
  if (!(cfg == 2 || cfg == 3 || cfg == 4 || cfg == 5 || cfg == 6)) return;
  b2_guided_geometry g;
  g.config = cfg;
  g.reserved = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      g.F[3 * r + c] = two_view_geometry->F(r, c);
      g.H[3 * r + c] = two_view_geometry->H(r, c);
    }
  static const unsigned char kEmptyDesc = 0;
  static const float kEmptyXy = 0;
  const unsigned char* dp[2];
  const float* kp[2];
  int32_t cnt[2];
  for (int k = 0; k < 2; ++k) {
    cnt[k] = (int32_t)(sift_match_gpu->guided_desc[k].size() / 128);
    dp[k] = cnt[k] ? sift_match_gpu->guided_desc[k].data() : &kEmptyDesc;
    kp[k] = cnt[k] ? sift_match_gpu->guided_xy[k].data() : &kEmptyXy;
  }
  b2_match_options o;
  o.max_ratio = (float)match_options.max_ratio;
  o.max_distance = (float)match_options.max_distance;
  o.cross_check = match_options.cross_check;
  o.max_num_matches = match_options.max_num_matches;
  const uint32_t pair[2] = {0, 1};
  int64_t offsets[2] = {0, 0}, total = 0;
  const int64_t cap = std::min(cnt[0], cnt[1]) > 0 ? (int64_t)cnt[0] : 0;
  two_view_geometry->inlier_matches.resize((size_t)std::max<int64_t>(cap, 1));
  b2_matcher* h = sift_match_gpu->handle();
  const bool ok = h != nullptr && b2_match_set_images(h, 2, dp, cnt) == B2_OK &&
                  b2_match_set_keypoints(h, 2, kp, cnt) == B2_OK &&
                  b2_match_guided_pairs(h, 1, pair, &g, match_options.max_error, &o, offsets,
                                        reinterpret_cast<uint32_t*>(two_view_geometry->inlier_matches.data()), cap,
                                        &total) == B2_OK;
  if (!ok) {
    std::fprintf(stderr, "ERROR: Feature matching failed. This is probably caused by insufficient GPU memory. "
                         "Consider reducing the maximum number of features and/or matches.\n");
    two_view_geometry->inlier_matches.clear();
  } else {
    two_view_geometry->inlier_matches.resize((size_t)total);
  }
}

}  // namespace dagsfm_b200
