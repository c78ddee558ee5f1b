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
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
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
    for (int k = 0; k < 2; ++k) { guided_desc[k].swap(o.guided_desc[k]); guided_xy[k].swap(o.guided_xy[k]); }
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
    const int n = std::max(0, std::min(num, max_sift_));
    if (!descriptors && n > 0) return;  // nothing to copy from: keep the previous upload
    b2_match_set_descriptors(h_, index, n, descriptors ? descriptors : &kEmpty);
    if (descriptors || n == 0) {  // host copy for GetGuidedSiftMatch, whose filter needs both images in one call
      guided_desc[index].assign(descriptors ? descriptors : &kEmpty, (descriptors ? descriptors : &kEmpty) + (size_t)n * 128);
      guided_xy[index].clear();
    }
  }
  // returns the number of matches, -1 on a device error (SiftMatchCU.cpp:193-196)
  int GetSiftMatch(int max_match, uint32_t match_buffer[][2], float distmax = 0.7f, float ratiomax = 0.8f,
                   int mutual_best_match = 1) {
    if (!h_ || max_match < 0) return -1;
    // max_match is the size of the caller's buffer; the feature clamp is max_sift (SiftMatchCU.cpp:108), applied at upload
    b2_match_options o;
    o.max_ratio = ratiomax; o.max_distance = distmax; o.cross_check = mutual_best_match; o.max_num_matches = std::max(max_sift_, 1);
    int32_t n = 0;
    if (max_match >= o.max_num_matches) {
      if (b2_match_run(h_, &o, &match_buffer[0][0], &n) != B2_OK) return -1;
      return n;
    }
    std::vector<uint32_t> tmp(2 * (size_t)o.max_num_matches);
    if (b2_match_run(h_, &o, tmp.data(), &n) != B2_OK) return -1;
    n = std::min(n, max_match);
    for (int i = 0; i < n; ++i) { match_buffer[i][0] = tmp[2 * i]; match_buffer[i][1] = tmp[2 * i + 1]; }
    return n;
  }

  int CreateContextGL() { return VerifyContextGL(); }
  // "-cuda [device_id]" (SiftGPU.h:308-312); other switches are OpenGL-only and ignored
  void SetDeviceParam(int argc, char** argv) {
    for (int i = 0; i + 1 < argc; ++i)
      if (std::string(argv[i]) == "-cuda") device_ = std::max(0, std::atoi(argv[i + 1]));
  }

  // Guided matching (SiftGPU.h:339-362): per image SetDescriptors, then SetFeautreLocation [sic]; then
  // GetGuidedSiftMatch.  `locations` is a vector of [float x, float y, float skip[gap]]; colmap passes its
  // FeatureKeypoint array with gap = 4 (sift.cc:1016-1018).  The two slots are kept on the host: the C ABI takes
  // the images of a guided call together.
  void SetFeautreLocation(int index, const float* locations, int gap = 0) {
    if (index < 0 || index > 1) return;
    const size_t n = guided_desc[index].size() / 128;
    guided_xy[index].resize(2 * n);
    for (size_t i = 0; i < n; ++i) {
      guided_xy[index][2 * i] = locations[i * (size_t)(2 + gap)];
      guided_xy[index][2 * i + 1] = locations[i * (size_t)(2 + gap) + 1];
    }
  }
  struct SiftKeypoint { float x, y, s, o; };  // SiftGPU::SiftKeypoint (SiftGPU.h:112-115)
  void SetFeatureLocation(int index, const SiftKeypoint* keys) { SetFeautreLocation(index, reinterpret_cast<const float*>(keys), 2); }

  // H / F: row-major float[9] or NULL (exactly one of them, as colmap calls it; both NULL = plain GetSiftMatch).
  // hdistmax / fdistmax are SQUARED pixel thresholds (colmap passes max_error^2 for both, sift.cc:1058-1062).
  // Returns the number of matches, -1 on a device error.
  int GetGuidedSiftMatch(int max_match, uint32_t match_buffer[][2], float* H, float* F, float distmax = 0.7f,
                         float ratiomax = 0.8f, float hdistmax = 32, float fdistmax = 16, int mutual_best_match = 1) {
    if (!h_) return -1;
    if (!H && !F) return GetSiftMatch(max_match, match_buffer, distmax, ratiomax, mutual_best_match);
    if (H && F) {
      std::fprintf(stderr, "ERROR: GetGuidedSiftMatch with both H and F is not supported (colmap passes one)\n");
      return -1;
    }
    b2_guided_geometry g;
    g.config = H ? 4 /* PLANAR */ : 3 /* UNCALIBRATED */;
    g.reserved = 0;
    for (int i = 0; i < 9; ++i) { g.F[i] = F ? (double)F[i] : 0.0; g.H[i] = H ? (double)H[i] : 0.0; }
    static const unsigned char kEmptyDesc = 0;
    static const float kEmptyXy = 0;
    const unsigned char* dp[2];
    const float* kp[2];
    int32_t cnt[2];
    for (int k = 0; k < 2; ++k) {
      cnt[k] = (int32_t)(guided_desc[k].size() / 128);
      if (guided_xy[k].size() != 2 * (size_t)cnt[k]) return -1;  // SetFeautreLocation missing for this slot
      dp[k] = cnt[k] ? guided_desc[k].data() : &kEmptyDesc;
      kp[k] = cnt[k] ? guided_xy[k].data() : &kEmptyXy;
    }
    b2_match_options o;  // feature clamp = max_sift (already applied by SetDescriptors); max_match only bounds the output
    o.max_ratio = ratiomax; o.max_distance = distmax; o.cross_check = mutual_best_match; o.max_num_matches = std::max(max_sift_, 1);
    const uint32_t pair[2] = {0, 1};
    int64_t offsets[2] = {0, 0}, total = 0;
    const double max_error = std::sqrt((double)(H ? hdistmax : fdistmax));
    std::vector<uint32_t> buf(2 * (size_t)std::max(cnt[0], 1));
    if (b2_match_set_images(h_, 2, dp, cnt) != B2_OK || b2_match_set_keypoints(h_, 2, kp, cnt) != B2_OK ||
        b2_match_guided_pairs(h_, 1, pair, &g, max_error, &o, offsets, buf.data(), (int64_t)std::max(cnt[0], 1), &total) != B2_OK)
      return -1;
    const int n = (int)std::min<int64_t>(total, max_match);
    for (int i = 0; i < n; ++i) { match_buffer[i][0] = buf[2 * i]; match_buffer[i][1] = buf[2 * i + 1]; }
    return n;
  }

  b2_matcher* handle() const { return h_; }
  std::vector<unsigned char> guided_desc[2];  // host copies of the two slots (guided matching)
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

// sift.cc:987-1079.  Keypoints: a contiguous container (data(), size()) of FeatureKeypoint-like structs whose first two
// floats are x, y, sizeof a multiple of 4 (colmap: 6 floats, passed with gap = 4); Geometry: any struct with `config`,
// `F`, `H` (3x3, indexable as M(r, c)) and `inlier_matches` (TwoViewGeometry, two_view_geometry.h:43-113).
// Configurations without a guided filter leave inlier_matches untouched, as the reference does (:1049-1051).
template <class Keypoints, class Descriptors, class Geometry>
void MatchGuidedSiftFeaturesGPU(const SiftMatchingOptions& match_options, const Keypoints* keypoints1,
                                const Keypoints* keypoints2, const Descriptors* descriptors1,
                                const Descriptors* descriptors2, SiftMatchGPU* sift_match_gpu,
                                Geometry* two_view_geometry) {
  using Keypoint = typename std::remove_reference<decltype((*keypoints1)[0])>::type;
  static_assert(sizeof(Keypoint) % sizeof(float) == 0 && sizeof(Keypoint) >= 2 * sizeof(float), "keypoints are float records");
  const int gap = (int)(sizeof(Keypoint) / sizeof(float)) - 2;  // kFeatureShapeNumElems = 4 for colmap's FeatureKeypoint
  auto stage = [&](int index, const Keypoints* kp, const Descriptors* d) {
    if (d == nullptr) return;  // NULL: same image as in the previous call (sift.cc:1009, :1022)
    if (sift_match_gpu->GetMaxSift() < (int)d->rows())
      std::printf("WARNING: Clamping features from %d to %d - consider increasing the maximum number of matches.\n",
                  (int)d->rows(), sift_match_gpu->GetMaxSift());
    sift_match_gpu->SetDescriptors(index, (int)d->rows(), d->data());
    static const float kNoKeypoint[2] = {0, 0};
    sift_match_gpu->SetFeautreLocation(index, kp->size() ? reinterpret_cast<const float*>(kp->data()) : kNoKeypoint, gap);
  };
  stage(0, keypoints1, descriptors1);
  stage(1, keypoints2, descriptors2);
  float F[9], H[9];
  float* F_ptr = nullptr;
  float* H_ptr = nullptr;
  const int cfg = (int)two_view_geometry->config;
  if (cfg == 2 || cfg == 3) {  // CALIBRATED, UNCALIBRATED
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[3 * r + c] = (float)two_view_geometry->F(r, c);
    F_ptr = F;
  } else if (cfg == 4 || cfg == 5 || cfg == 6) {  // PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] = (float)two_view_geometry->H(r, c);
    H_ptr = H;
  } else {
    return;
  }
  two_view_geometry->inlier_matches.resize(static_cast<size_t>(match_options.max_num_matches));
  const int num_matches = sift_match_gpu->GetGuidedSiftMatch(
      match_options.max_num_matches, reinterpret_cast<uint32_t(*)[2]>(two_view_geometry->inlier_matches.data()), H_ptr, F_ptr,
      static_cast<float>(match_options.max_distance), static_cast<float>(match_options.max_ratio),
      static_cast<float>(match_options.max_error * match_options.max_error),
      static_cast<float>(match_options.max_error * match_options.max_error), match_options.cross_check);
  if (num_matches < 0) {
    std::fprintf(stderr, "ERROR: Feature matching failed. This is probably caused by insufficient GPU memory. "
                         "Consider reducing the maximum number of features and/or matches.\n");
    two_view_geometry->inlier_matches.clear();
  } else {
    two_view_geometry->inlier_matches.resize(num_matches);
  }
}

// SiftGPU.h:369: the factory the reference loads (SIFTGPU_EXPORT_EXTERN).  The caller owns the object.
inline SiftMatchGPU* CreateNewSiftMatchGPU(int max_sift = 4096) { return new SiftMatchGPU(max_sift); }

}  // namespace dagsfm_b200
