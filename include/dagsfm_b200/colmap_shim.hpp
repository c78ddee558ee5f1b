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
  bool Check() const { return max_ratio > 0 && max_distance > 0 && max_num_matches > 0; }
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

}  // namespace dagsfm_b200
