// Header-only C++ adaptor that gives the verification C ABI (include/dagsfm_b200.h, VERIFY section)
// the reference's own class for this path, so that the reference's callers compile unchanged:
//
//   TwoViewGeometry                 src/estimators/two_view_geometry.h:43-303
//     ::Estimate                    src/estimators/two_view_geometry.cc:113-126   (DAGSfM: calibrated pairs
//                                                                continue with EstimateWithRelativePose)
//     ::EstimateMultiple            :128-167
//     ::EstimateCalibrated / EstimateUncalibrated / EstimateWithRelativePose   :232-489
//     ::Invert                      :95-111
//   RANSACOptions                   src/optim/ransac.h:46-72
//   SetPRNGSeed                     src/util/random.cc:44-53 (thread-local PRNG of the verifier thread)
//
// The reference cannot be built in this environment (Eigen / glog / Boost absent), so the adaptor is a
// template over the few members it touches: a Camera needs ModelId(), Width(), Height(),
// HasPriorFocalLength() and Params() (a contiguous container of double); a point needs operator()(int)
// (Eigen::Vector2d); FeatureMatches is any std::vector of {uint32 point2D_idx1, point2D_idx2}.
// Matrices / vectors are small row-major value types indexable as M(r, c) / v(i) like Eigen's.
//
// One call verifies one pair (the reference's calling convention, matching.cc:571-608); the handle is
// cached per thread and device.  For throughput use the batched C ABI directly.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <random>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {

#ifndef DAGSFM_B200_FEATURE_MATCH_DEFINED
#define DAGSFM_B200_FEATURE_MATCH_DEFINED
struct FeatureMatch2 {  // src/feature/types.h:86-98
  FeatureMatch2() = default;
  FeatureMatch2(uint32_t a, uint32_t b) : point2D_idx1(a), point2D_idx2(b) {}
  uint32_t point2D_idx1 = 0xffffffffu;
  uint32_t point2D_idx2 = 0xffffffffu;
};
#endif

struct Mat3 {  // stands in for Eigen::Matrix3d (row-major storage)
  std::array<double, 9> m{};
  double& operator()(int r, int c) { return m[3 * r + c]; }
  double operator()(int r, int c) const { return m[3 * r + c]; }
  const double* data() const { return m.data(); }
  double* data() { return m.data(); }
  static Mat3 Zero() { return Mat3(); }
  static Mat3 Identity() { Mat3 r; r(0, 0) = r(1, 1) = r(2, 2) = 1; return r; }
  void transposeInPlace() { std::swap(m[1], m[3]); std::swap(m[2], m[6]); std::swap(m[5], m[7]); }
  Mat3 inverse() const {
    const Mat3& a = *this;
    Mat3 r;
    const double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1), c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2),
                 c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    const double id = 1.0 / (a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02);
    r(0, 0) = c00 * id; r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id; r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
    r(1, 0) = c01 * id; r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id; r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
    r(2, 0) = c02 * id; r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id; r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
    return r;
  }
  bool operator==(const Mat3& o) const { return m == o.m; }
};
template <int N>
struct VecN {  // Eigen::Vector3d / Vector4d
  std::array<double, N> v{};
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  static VecN Zero() { return VecN(); }
  double norm() const { double s = 0; for (double x : v) s += x * x; return std::sqrt(s); }
  bool operator==(const VecN& o) const { return v == o.v; }
};
using Vec3 = VecN<3>;
using Vec4 = VecN<4>;

struct RANSACOptions {  // src/optim/ransac.h:46-72
  double max_error = 0.0;
  double min_inlier_ratio = 0.1;
  double confidence = 0.99;
  size_t min_num_trials = 0;
  size_t max_num_trials = std::numeric_limits<size_t>::max();
  bool Check() const {
    return max_error > 0 && min_inlier_ratio >= 0 && min_inlier_ratio <= 1 && confidence >= 0 && confidence <= 1 &&
           min_num_trials <= max_num_trials;
  }
};

// util/random.cc:38-53: the verifier threads' thread-local std::mt19937, seeded from the clock unless
// SetPRNGSeed was called on the thread.  Each Estimate call draws one 32-bit seed for the device-side stream.
inline std::mt19937*& ThreadPRNG() {
  thread_local std::mt19937* prng = nullptr;
  return prng;
}
inline void SetPRNGSeed(unsigned seed) {
  delete ThreadPRNG();
  ThreadPRNG() = new std::mt19937(seed);
}
inline uint32_t NextPairSeed() {
  if (ThreadPRNG() == nullptr)
    SetPRNGSeed(static_cast<unsigned>(std::chrono::system_clock::now().time_since_epoch().count()));
  return static_cast<uint32_t>((*ThreadPRNG())());
}

template <class FeatureMatchT = FeatureMatch2>
struct TwoViewGeometryT {
  using FeatureMatches = std::vector<FeatureMatchT>;
  static_assert(sizeof(FeatureMatchT) == 8, "FeatureMatch must be two packed uint32");

  enum ConfigurationType {  // two_view_geometry.h:47-102
    UNDEFINED = 0, DEGENERATE = 1, CALIBRATED = 2, UNCALIBRATED = 3, PLANAR = 4, PANORAMIC = 5,
    PLANAR_OR_PANORAMIC = 6, WATERMARK = 7, MULTIPLE = 8,
  };

  struct Options {  // two_view_geometry.h:105-157
    size_t min_num_inliers = 15;
    double min_E_F_inlier_ratio = 0.95;
    double max_H_inlier_ratio = 0.8;
    double watermark_min_inlier_ratio = 0.7;
    double watermark_border_size = 0.1;
    bool detect_watermark = true;
    bool multiple_ignore_watermark = true;
    RANSACOptions ransac_options;
    int gpu_index = 0;  // not in the reference: the device that verifies
    bool Check() const {
      return min_E_F_inlier_ratio >= 0 && min_E_F_inlier_ratio <= 1 && max_H_inlier_ratio >= 0 && max_H_inlier_ratio <= 1 &&
             watermark_min_inlier_ratio >= 0 && watermark_min_inlier_ratio <= 1 && watermark_border_size >= 0 &&
             watermark_border_size <= 1 && ransac_options.Check();
    }
  };

  TwoViewGeometryT() = default;  // two_view_geometry.h:159-166: everything zero, config UNDEFINED

  // two_view_geometry.cc:95-111
  void Invert() {
    F.transposeInPlace();
    E.transposeInPlace();
    H = H.inverse();
    // InvertPose (base/pose.cc:192-196): q^-1 = (w, -x, -y, -z); t' = -(q^-1 * t) with the NORMALISED quaternion
    // (NormalizeQuaternion, :82-91: a zero quaternion becomes (1, x, y, z))
    const Vec4 oq = qvec;
    const Vec3 ot = tvec;
    qvec(0) = oq(0); qvec(1) = -oq(1); qvec(2) = -oq(2); qvec(3) = -oq(3);
    Vec4 n = qvec;
    const double nn = n.norm();
    if (nn == 0) n(0) = 1.0;
    else for (int i = 0; i < 4; ++i) n(i) /= nn;
    const double w = n(0), x = n(1), y = n(2), z = n(3);
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                            {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                            {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
    for (int r = 0; r < 3; ++r) tvec(r) = -(R[r][0] * ot(0) + R[r][1] * ot(1) + R[r][2] * ot(2));
    for (auto& match : inlier_matches) std::swap(match.point2D_idx1, match.point2D_idx2);
  }

  // two_view_geometry.cc:113-126
  template <class Camera, class Points>
  void Estimate(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                const FeatureMatches& matches, const Options& options) {
    if (camera1.HasPriorFocalLength() && camera2.HasPriorFocalLength())
      EstimateWithRelativePose(camera1, points1, camera2, points2, matches, options);
    else
      EstimateUncalibrated(camera1, points1, camera2, points2, matches, options);
  }

  // two_view_geometry.cc:128-167.  As in the reference every round is a full Estimate, so a single
  // surviving geometry of a calibrated pair carries its relative pose.
  template <class Camera, class Points>
  void EstimateMultiple(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                        const FeatureMatches& matches, const Options& options) {
    Run(camera1, points1, camera2, points2, matches, options, camera1.HasPriorFocalLength(), camera2.HasPriorFocalLength(),
        /*multiple=*/true, /*pose=*/camera1.HasPriorFocalLength() && camera2.HasPriorFocalLength());
  }

  // two_view_geometry.cc:292-425: E + F + H regardless of the cameras' prior flags
  template <class Camera, class Points>
  void EstimateCalibrated(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                          const FeatureMatches& matches, const Options& options) {
    Run(camera1, points1, camera2, points2, matches, options, true, true, false, false);
  }
  // two_view_geometry.cc:427-489: F + H only
  template <class Camera, class Points>
  void EstimateUncalibrated(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                            const FeatureMatches& matches, const Options& options) {
    Run(camera1, points1, camera2, points2, matches, options, false, false, false, false);
  }
  // two_view_geometry.cc:232-290
  template <class Camera, class Points>
  void EstimateWithRelativePose(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                                const FeatureMatches& matches, const Options& options) {
    Run(camera1, points1, camera2, points2, matches, options, true, true, false, true);
  }

  // two_view_geometry.cc:169-230: relative pose of an ALREADY estimated geometry (config, E / H, inlier_matches), as the
  // incremental mapper calls it for its initial pair (sfm/incremental_mapper.cc:1161).  False when the configuration
  // carries no epipolar geometry or homography to decompose.
  template <class Camera, class Points>
  bool EstimateRelativePose(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
                            int gpu_index = 0) {
    if (config != CALIBRATED && config != UNCALIBRATED && config != PLANAR && config != PANORAMIC && config != PLANAR_OR_PANORAMIC)
      return false;
    b2_verifier* v = Verifier(gpu_index);
    if (!v) return false;
    const b2_camera cams[2] = {PackCamera(camera1, true), PackCamera(camera2, true)};
    const std::vector<double> xy1 = PackPoints(points1), xy2 = PackPoints(points2);
    static const double kNone[2] = {0, 0};
    const double* xy[2] = {xy1.empty() ? kNone : xy1.data(), xy2.empty() ? kNone : xy2.data()};
    const int32_t n_pts[2] = {(int32_t)points1.size(), (int32_t)points2.size()};
    b2_two_view_result res = {};
    res.config = config;
    res.n_inliers = (int32_t)inlier_matches.size();
    for (int i = 0; i < 9; ++i) { res.E[i] = E.m[i]; res.F[i] = F.m[i]; res.H[i] = H.m[i]; }
    const uint32_t pair[2] = {0, 1};
    const int64_t offsets[2] = {0, (int64_t)inlier_matches.size()};
    static const uint32_t kNoMatch[2] = {0, 0};
    const uint32_t* iptr = inlier_matches.empty() ? kNoMatch : reinterpret_cast<const uint32_t*>(inlier_matches.data());
    b2_relative_pose rp;
    if (b2_verify_set_images(v, 2, cams, xy, n_pts) != B2_OK || b2_verify_relative_pose(v, 1, pair, offsets, &res, iptr, &rp) != B2_OK) {
      std::fprintf(stderr, "ERROR: relative pose failed: %s\n", b2_last_error());
      return false;
    }
    config = rp.config;
    for (int i = 0; i < 4; ++i) qvec(i) = rp.qvec[i];
    for (int i = 0; i < 3; ++i) tvec(i) = rp.tvec[i];
    tri_angle = rp.tri_angle;
    return true;
  }

  int config = UNDEFINED;
  Mat3 E, F, H;
  Vec4 qvec;
  Vec3 tvec;
  FeatureMatches inlier_matches;
  double tri_angle = 0;
  size_t E_num_inliers = 0, F_num_inliers = 0, H_num_inliers = 0, T_num_tracks = 0;

 private:
  struct Handle {
    b2_verifier* v = nullptr;
    int device = -1;
    ~Handle() { if (v) b2_verify_destroy(v); }
  };
  static b2_verifier* Verifier(int device) {
    thread_local Handle h;
    if (h.v && h.device != device) { b2_verify_destroy(h.v); h.v = nullptr; }
    if (!h.v) {
      if (b2_verify_create(device, &h.v) != B2_OK) { h.v = nullptr; return nullptr; }
      h.device = device;
    }
    return h.v;
  }
  template <class Camera>
  static b2_camera PackCamera(const Camera& c, bool prior) {
    b2_camera o;
    o.model = (int32_t)c.ModelId();
    o.width = (int32_t)c.Width();
    o.height = (int32_t)c.Height();
    o.has_prior_focal_length = prior ? 1 : 0;
    for (double& p : o.params) p = 0;
    const auto& prm = c.Params();
    for (size_t i = 0; i < prm.size() && i < 12; ++i) o.params[i] = prm[i];
    return o;
  }
  template <class Points>
  static std::vector<double> PackPoints(const Points& pts) {
    std::vector<double> xy(2 * pts.size());
    for (size_t i = 0; i < pts.size(); ++i) { xy[2 * i] = pts[i](0); xy[2 * i + 1] = pts[i](1); }
    return xy;
  }

  template <class Camera, class Points>
  void Run(const Camera& camera1, const Points& points1, const Camera& camera2, const Points& points2,
           const FeatureMatches& matches, const Options& options, bool prior1, bool prior2, bool multiple, bool pose) {
    if (!options.Check()) {  // options.Check() is a glog CHECK in the reference: abort
      std::fprintf(stderr, "Check failed: TwoViewGeometry::Options\n");
      std::abort();
    }
    *this = TwoViewGeometryT();
    b2_verifier* v = Verifier(options.gpu_index);
    if (!v) {
      std::fprintf(stderr, "ERROR: two-view verification failed: %s\n", b2_last_error());
      config = DEGENERATE;
      return;
    }
    const b2_camera cams[2] = {PackCamera(camera1, prior1), PackCamera(camera2, prior2)};
    const std::vector<double> xy1 = PackPoints(points1), xy2 = PackPoints(points2);
    static const double kNone[2] = {0, 0};
    const double* xy[2] = {xy1.empty() ? kNone : xy1.data(), xy2.empty() ? kNone : xy2.data()};
    const int32_t n_pts[2] = {(int32_t)points1.size(), (int32_t)points2.size()};
    b2_two_view_options o;
    b2_two_view_default_options(&o);
    o.min_num_inliers = (int32_t)std::min<size_t>(options.min_num_inliers, 0x7fffffff);
    o.detect_watermark = options.detect_watermark ? 1 : 0;
    o.min_E_F_inlier_ratio = options.min_E_F_inlier_ratio;
    o.max_H_inlier_ratio = options.max_H_inlier_ratio;
    o.watermark_min_inlier_ratio = options.watermark_min_inlier_ratio;
    o.watermark_border_size = options.watermark_border_size;
    o.max_error = options.ransac_options.max_error;
    o.min_inlier_ratio = options.ransac_options.min_inlier_ratio;
    o.confidence = options.ransac_options.confidence;
    const size_t kMax = (size_t)std::numeric_limits<int64_t>::max();
    o.min_num_trials = (int64_t)std::min(options.ransac_options.min_num_trials, kMax);
    o.max_num_trials = (int64_t)std::min(options.ransac_options.max_num_trials, kMax);
    const uint32_t pair[2] = {0, 1};
    const int64_t offsets[2] = {0, (int64_t)matches.size()};
    const uint32_t seed = NextPairSeed();
    b2_two_view_result res;
    static const uint32_t kNoMatch[2] = {0, 0};
    const uint32_t* mptr = matches.empty() ? kNoMatch : reinterpret_cast<const uint32_t*>(matches.data());
    std::vector<FeatureMatchT> inl(std::max<size_t>(matches.size(), 1));
    int rc = b2_verify_set_images(v, 2, cams, xy, n_pts);
    if (rc == B2_OK)
      rc = multiple ? b2_verify_pairs_multiple(v, 1, pair, offsets, mptr, &o, options.multiple_ignore_watermark ? 1 : 0, &seed,
                                               &res, reinterpret_cast<uint32_t*>(inl.data()))
                    : b2_verify_pairs(v, 1, pair, offsets, mptr, &o, &seed, &res, reinterpret_cast<uint32_t*>(inl.data()));
    if (rc != B2_OK) {
      std::fprintf(stderr, "ERROR: two-view verification failed: %s\n", b2_last_error());
      config = DEGENERATE;
      return;
    }
    config = res.config;
    for (int i = 0; i < 9; ++i) { E.m[i] = res.E[i]; F.m[i] = res.F[i]; H.m[i] = res.H[i]; }
    E_num_inliers = (size_t)res.E_num_inliers;
    F_num_inliers = (size_t)res.F_num_inliers;
    H_num_inliers = (size_t)res.H_num_inliers;
    inl.resize((size_t)res.n_inliers);
    inlier_matches = std::move(inl);
    if (pose && config != MULTIPLE) {
      b2_relative_pose rp;
      const uint32_t* iptr = inlier_matches.empty() ? kNoMatch : reinterpret_cast<const uint32_t*>(inlier_matches.data());
      if (b2_verify_relative_pose(v, 1, pair, offsets, &res, iptr, &rp) == B2_OK) {
        config = rp.config;
        for (int i = 0; i < 4; ++i) qvec(i) = rp.qvec[i];
        for (int i = 0; i < 3; ++i) tvec(i) = rp.tvec[i];
        tri_angle = rp.tri_angle;
      } else {
        std::fprintf(stderr, "ERROR: relative pose failed: %s\n", b2_last_error());
      }
    }
  }
};

using TwoViewGeometry = TwoViewGeometryT<>;

}  // namespace dagsfm_b200
