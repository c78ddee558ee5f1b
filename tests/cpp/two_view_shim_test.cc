// Replays the reference's src/estimators/two_view_geometry_test.cc (TestDefault, TestInvert) on the C++ adaptor
// include/dagsfm_b200/two_view_shim.hpp -- host-only, runs without a device -- and, when called with the argument
// "estimate", drives TwoViewGeometry::Estimate / EstimateMultiple / EstimateUncalibrated through the C ABI on a
// synthetic pair with a known relative pose.  Built by tests/test_two_view_shim.py against the product library
// (GPU) or against the CUDA-emulator build of the same sources (CPU test of the adaptor).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "dagsfm_b200/two_view_shim.hpp"

using namespace dagsfm_b200;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

struct Camera {  // the members of colmap::Camera the adaptor reads
  int model = 2;
  size_t w = 1000, h = 1000;
  bool prior = true;
  std::vector<double> params{1200.0, 500.0, 500.0, 0.0};
  int ModelId() const { return model; }
  size_t Width() const { return w; }
  size_t Height() const { return h; }
  bool HasPriorFocalLength() const { return prior; }
  const std::vector<double>& Params() const { return params; }
};
struct P2 { double x, y; double operator()(int i) const { return i == 0 ? x : y; } };

static bool approx(double a, double b, double tol = 1e-12) { return std::fabs(a - b) <= tol; }

static int test_default_and_invert() {
  TwoViewGeometry g;  // TestDefault
  CHECK(g.config == TwoViewGeometry::UNDEFINED);
  CHECK(g.F == Mat3::Zero() && g.E == Mat3::Zero() && g.H == Mat3::Zero());
  CHECK(g.qvec == Vec4::Zero() && g.tvec == Vec3::Zero());
  CHECK(g.inlier_matches.empty());
  // TestInvert
  g.config = TwoViewGeometry::CALIBRATED;
  g.F = g.E = g.H = Mat3::Identity();
  g.qvec(0) = 1; g.qvec(1) = g.qvec(2) = g.qvec(3) = 0;  // ComposeIdentityQuaternion
  g.tvec(0) = 0; g.tvec(1) = 1; g.tvec(2) = 2;
  g.inlier_matches.resize(2);
  g.inlier_matches[0] = FeatureMatch2(0, 1);
  g.inlier_matches[1] = FeatureMatch2(2, 3);
  for (int round = 0; round < 2; ++round) {
    g.Invert();
    const double s = round == 0 ? -1.0 : 1.0;
    CHECK(g.config == TwoViewGeometry::CALIBRATED);
    CHECK(g.F == Mat3::Identity() && g.E == Mat3::Identity() && g.H == Mat3::Identity());
    CHECK(approx(g.qvec(0), 1) && approx(g.qvec(1), 0) && approx(g.qvec(2), 0) && approx(g.qvec(3), 0));
    CHECK(approx(g.tvec(0), 0) && approx(g.tvec(1), s * 1) && approx(g.tvec(2), s * 2));
    CHECK(g.inlier_matches[0].point2D_idx1 == (round == 0 ? 1u : 0u) && g.inlier_matches[0].point2D_idx2 == (round == 0 ? 0u : 1u));
    CHECK(g.inlier_matches[1].point2D_idx1 == (round == 0 ? 3u : 2u) && g.inlier_matches[1].point2D_idx2 == (round == 0 ? 2u : 3u));
  }
  // a general pose: inverting twice is the identity, and x2 = R x1 + t  <=>  x1 = R' x2 + t'
  TwoViewGeometry p;
  const double a = 0.3;
  p.qvec(0) = std::cos(a / 2); p.qvec(2) = std::sin(a / 2);  // rotation about y
  p.tvec(0) = 0.5; p.tvec(1) = -0.2; p.tvec(2) = 1.0;
  p.H = Mat3::Identity(); p.H(0, 2) = 3; p.H(1, 1) = 2;
  TwoViewGeometry q = p;
  q.Invert();
  CHECK(approx(q.H(0, 2), -3) && approx(q.H(1, 1), 0.5));
  // R(a) (0.5,-0.2,1) negated and rotated back
  const double c = std::cos(a), s = std::sin(a);
  CHECK(approx(q.tvec(0), -(c * 0.5 - s * 1.0)) && approx(q.tvec(1), 0.2) && approx(q.tvec(2), -(s * 0.5 + c * 1.0)));
  q.Invert();
  for (int i = 0; i < 4; ++i) CHECK(approx(q.qvec(i), p.qvec(i)));
  for (int i = 0; i < 3; ++i) CHECK(approx(q.tvec(i), p.tvec(i)));
  // options defaults of the reference (two_view_geometry.h:105-157, ransac.h:46-72)
  TwoViewGeometry::Options o;
  CHECK(o.min_num_inliers == 15 && o.min_E_F_inlier_ratio == 0.95 && o.max_H_inlier_ratio == 0.8 && o.detect_watermark);
  CHECK(o.ransac_options.max_error == 0.0 && o.ransac_options.confidence == 0.99 && o.ransac_options.min_inlier_ratio == 0.1);
  CHECK(!o.Check());  // max_error must be set by the caller, as RANSACOptions::Check demands
  o.ransac_options.max_error = 4.0;
  CHECK(o.Check());
  return 0;
}

static int test_estimate() {
  std::mt19937 rng(7);
  std::normal_distribution<double> noise(0.0, 0.3);
  std::uniform_real_distribution<double> u(-1.0, 1.0), px(0.0, 1000.0);
  const double ang = 0.15, t[3] = {-1.0, 0.1, 0.2};
  const double R[3][3] = {{std::cos(ang), 0, std::sin(ang)}, {0, 1, 0}, {-std::sin(ang), 0, std::cos(ang)}};
  std::vector<P2> p1, p2;
  const int n_in = 200, n_out = 40;
  for (int i = 0; i < n_in; ++i) {
    const double X[3] = {2 * u(rng), 2 * u(rng), 8 + u(rng)};
    double Y[3];
    for (int r = 0; r < 3; ++r) Y[r] = R[r][0] * X[0] + R[r][1] * X[1] + R[r][2] * X[2] + t[r];
    p1.push_back({1200 * X[0] / X[2] + 500 + noise(rng), 1200 * X[1] / X[2] + 500 + noise(rng)});
    p2.push_back({1200 * Y[0] / Y[2] + 500 + noise(rng), 1200 * Y[1] / Y[2] + 500 + noise(rng)});
  }
  for (int i = 0; i < n_out; ++i) { p1.push_back({px(rng), px(rng)}); p2.push_back({px(rng), px(rng)}); }
  TwoViewGeometry::FeatureMatches matches;
  for (uint32_t i = 0; i < p1.size(); ++i) matches.emplace_back(i, i);
  TwoViewGeometry::Options opt;   // the matcher's values (feature/sift.h:142-158, matching.cc:557-569)
  opt.ransac_options.max_error = 4.0;
  opt.ransac_options.confidence = 0.999;
  opt.ransac_options.min_num_trials = 30;
  opt.ransac_options.max_num_trials = 10000;
  opt.ransac_options.min_inlier_ratio = 0.25;
  Camera cam;
  SetPRNGSeed(5);
  TwoViewGeometry g;
  g.Estimate(cam, p1, cam, p2, matches, opt);
  CHECK(g.config == TwoViewGeometry::CALIBRATED);
  CHECK(g.inlier_matches.size() >= 0.95 * n_in && g.inlier_matches.size() <= (size_t)n_in + 3);
  CHECK(g.E_num_inliers >= 0.9 * n_in && g.F_num_inliers >= 0.9 * n_in);
  CHECK(approx(g.qvec.norm(), 1.0, 1e-9) && approx(g.tvec.norm(), 1.0, 1e-9));
  CHECK(std::fabs(g.qvec(0) - std::cos(ang / 2)) < 2e-2 && std::fabs(std::fabs(g.qvec(2)) - std::sin(ang / 2)) < 2e-2);
  const double tn = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  CHECK((g.tvec(0) * t[0] + g.tvec(1) * t[1] + g.tvec(2) * t[2]) / tn > 0.98);
  CHECK(g.tri_angle > 0.08 && g.tri_angle < 0.18);
  // same seed, same answer (the reference's SetPRNGSeed contract)
  SetPRNGSeed(5);
  TwoViewGeometry g2;
  g2.Estimate(cam, p1, cam, p2, matches, opt);
  CHECK(g2.config == g.config && g2.inlier_matches.size() == g.inlier_matches.size() && g2.E == g.E && g2.qvec == g.qvec);
  // a camera without prior focal length: EstimateUncalibrated, no pose (qvec stays zero as constructed)
  Camera nop = cam;
  nop.prior = false;
  TwoViewGeometry gu;
  gu.Estimate(cam, p1, nop, p2, matches, opt);
  CHECK(gu.config == TwoViewGeometry::UNCALIBRATED && gu.E == Mat3::Zero() && gu.qvec == Vec4::Zero() && gu.tri_angle == 0);
  CHECK(gu.inlier_matches.size() >= 0.95 * n_in);
  // Invert maps the geometry to the swapped pair: F^T is the fundamental matrix of (image 2, image 1)
  TwoViewGeometry gi = g;
  gi.Invert();
  CHECK(gi.F(0, 1) == g.F(1, 0) && gi.inlier_matches[0].point2D_idx1 == g.inlier_matches[0].point2D_idx2);
  // too few matches: DEGENERATE without touching the device results
  TwoViewGeometry::FeatureMatches few(matches.begin(), matches.begin() + 10);
  TwoViewGeometry gd;
  gd.Estimate(cam, p1, cam, p2, few, opt);
  CHECK(gd.config == TwoViewGeometry::DEGENERATE && gd.inlier_matches.empty() && gd.qvec == Vec4::Zero());
  // EstimateMultiple on a single rigid motion returns that one geometry, pose included
  SetPRNGSeed(9);
  TwoViewGeometry gm;
  gm.EstimateMultiple(cam, p1, cam, p2, matches, opt);
  CHECK(gm.config == TwoViewGeometry::CALIBRATED && gm.inlier_matches.size() >= 0.95 * n_in && approx(gm.qvec.norm(), 1.0, 1e-9));
  // EstimateRelativePose on a stored geometry (incremental_mapper.cc:1161): same pose as the one Estimate attached;
  // configurations without E / H to decompose are refused
  TwoViewGeometry gr = g;
  gr.qvec = Vec4::Zero(); gr.tvec = Vec3::Zero(); gr.tri_angle = 0;
  CHECK(gr.EstimateRelativePose(cam, p1, cam, p2));
  CHECK(gr.qvec == g.qvec && gr.tvec == g.tvec && gr.tri_angle == g.tri_angle && gr.config == g.config);
  TwoViewGeometry gw = g;
  gw.config = TwoViewGeometry::WATERMARK;
  CHECK(!gw.EstimateRelativePose(cam, p1, cam, p2));
  gw.config = TwoViewGeometry::DEGENERATE;
  CHECK(!gw.EstimateRelativePose(cam, p1, cam, p2));
  std::printf("two-view shim ok: %zu inliers, tri_angle %.4f\n", g.inlier_matches.size(), g.tri_angle);
  return 0;
}

int main(int argc, char** argv) {
  if (test_default_and_invert()) return 1;
  std::printf("two-view shim host tests ok\n");
  if (argc > 1 && !std::strcmp(argv[1], "estimate")) return test_estimate();
  return 0;
}
