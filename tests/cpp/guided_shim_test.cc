// Replays the reference's TestMatchGuidedSiftFeaturesGPU (src/feature/sift_test.cc:580-676) through
// dagsfm_b200::MatchGuidedSiftFeaturesGPU with stand-ins for Eigen / colmap types.
// Built by tests/test_zz_guided_gpu.py; exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "dagsfm_b200/colmap_shim.hpp"

struct Desc {
  std::vector<unsigned char> v;
  long n = 0;
  long rows() const { return n; }
  long cols() const { return 128; }
  const unsigned char* data() const { return v.data(); }
};
struct Keypoint { float x = 0, y = 0, a11 = 1, a12 = 0, a21 = 0, a22 = 1; };  // FeatureKeypoint, types.h:40-72
using Keypoints = std::vector<Keypoint>;
struct Mat3 {
  double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double operator()(int r, int c) const { return m[3 * r + c]; }
};
struct Geometry {  // TwoViewGeometry stand-in
  int config = 0;
  Mat3 E, F, H;
  dagsfm_b200::FeatureMatches inlier_matches;
};

static Desc make(int n, unsigned seed) {
  Desc d;
  d.n = n;
  d.v.resize((size_t)n * 128);
  unsigned s = seed * 2654435761u + 12345u;
  for (int i = 0; i < n; ++i) {
    double norm = 0, r[128];
    for (int k = 0; k < 128; ++k) { s = s * 1664525u + 1013904223u; const double u = (s >> 8) / 16777216.0; r[k] = u * u; norm += r[k] * r[k]; }
    norm = std::sqrt(norm);
    for (int k = 0; k < 128; ++k) { double x = std::floor(512.0 * r[k] / norm + 0.5); d.v[(size_t)i * 128 + k] = (unsigned char)(x > 255 ? 255 : x); }
  }
  return d;
}

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  using namespace dagsfm_b200;
  SiftMatchingOptions opt;
  opt.gpu_index = "0";
  SiftMatchGPU gpu;
  CHECK(CreateSiftGPUMatcher(opt, &gpu));
  Keypoints empty_kp, k1(2), k2(2);
  k1[0].x = 1; k1[1].x = 2;
  k2[0].x = 2; k2[1].x = 1;
  Desc empty, d1 = make(2, 7), d2 = d1;
  memcpy(&d2.v[0], &d1.v[128], 128);  // descriptors1.colwise().reverse()
  memcpy(&d2.v[128], &d1.v[0], 128);
  Geometry g;
  g.config = 6;  // PLANAR_OR_PANORAMIC, H = identity
  MatchGuidedSiftFeaturesGPU(opt, &k1, &k2, &d1, &d2, &gpu, &g);
  CHECK(g.inlier_matches.size() == 2);
  CHECK(g.inlier_matches[0].point2D_idx1 == 0 && g.inlier_matches[0].point2D_idx2 == 1);
  CHECK(g.inlier_matches[1].point2D_idx1 == 1 && g.inlier_matches[1].point2D_idx2 == 0);
  // NULL = same images as before (sift_test.cc:623-650)
  MatchGuidedSiftFeaturesGPU<Keypoints, Desc, Geometry>(opt, nullptr, nullptr, nullptr, nullptr, &gpu, &g);
  CHECK(g.inlier_matches.size() == 2);
  MatchGuidedSiftFeaturesGPU<Keypoints, Desc, Geometry>(opt, &k1, nullptr, &d1, nullptr, &gpu, &g);
  CHECK(g.inlier_matches.size() == 2);
  MatchGuidedSiftFeaturesGPU<Keypoints, Desc, Geometry>(opt, nullptr, &k2, nullptr, &d2, &gpu, &g);
  CHECK(g.inlier_matches.size() == 2);
  // moving a keypoint out of the 4 px disc removes its match (sift_test.cc:652-657)
  k1[0].x = 100;
  MatchGuidedSiftFeaturesGPU(opt, &k1, &k2, &d1, &d2, &gpu, &g);
  CHECK(g.inlier_matches.size() == 1);
  CHECK(g.inlier_matches[0].point2D_idx1 == 1 && g.inlier_matches[0].point2D_idx2 == 0);
  // empty inputs
  MatchGuidedSiftFeaturesGPU(opt, &empty_kp, &k2, &empty, &d2, &gpu, &g);
  CHECK(g.inlier_matches.empty());
  MatchGuidedSiftFeaturesGPU(opt, &k1, &empty_kp, &d1, &empty, &gpu, &g);
  CHECK(g.inlier_matches.empty());
  MatchGuidedSiftFeaturesGPU(opt, &empty_kp, &empty_kp, &empty, &empty, &gpu, &g);
  CHECK(g.inlier_matches.empty());
  // a configuration without a guided filter leaves inlier_matches alone (sift.cc:1049-1051)
  g.config = 7;
  g.inlier_matches.assign(3, FeatureMatch());
  MatchGuidedSiftFeaturesGPU(opt, &k1, &k2, &d1, &d2, &gpu, &g);
  CHECK(g.inlier_matches.size() == 3);
  // the SiftMatchGPU interface itself (SiftGPU.h:339-362), as a caller outside colmap's wrapper would drive it: the
  // factory, SetDescriptors + SetFeautreLocation per image, GetGuidedSiftMatch with H or with F
  SiftMatchGPU* raw = CreateNewSiftMatchGPU(64);
  char arg0[] = "-cuda", arg1[] = "0";
  char* argv[] = {arg0, arg1};
  raw->SetDeviceParam(2, argv);
  raw->SetLanguage(SiftMatchGPU::SIFTMATCH_CUDA);
  CHECK(raw->CreateContextGL() != 0 && raw->Allocate(64, 1) && raw->GetMaxSift() == 64);
  k1[0].x = 1; k1[1].x = 2;
  uint32_t buf[8][2];
  float Hf[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  raw->SetDescriptors(0, 2, d1.data());
  raw->SetFeautreLocation(0, reinterpret_cast<const float*>(k1.data()), 4);
  raw->SetDescriptors(1, 2, d2.data());
  CHECK(raw->GetGuidedSiftMatch(8, buf, Hf, nullptr, 0.7f, 0.8f, 16.f, 16.f, 1) == -1);   // second image has no locations yet
  raw->SetFeautreLocation(1, reinterpret_cast<const float*>(k2.data()), 4);
  CHECK(raw->GetGuidedSiftMatch(8, buf, Hf, nullptr, 0.7f, 0.8f, 16.f, 16.f, 1) == 2);
  CHECK(buf[0][0] == 0 && buf[0][1] == 1 && buf[1][0] == 1 && buf[1][1] == 0);
  // F of a pure x-translation: x2' F x1 = y1 - y2 = 0 for these keypoints whatever their x
  float Ff[9] = {0, 0, 0, 0, 0, -1, 0, 1, 0};
  k1[0].x = 500;
  raw->SetFeautreLocation(0, reinterpret_cast<const float*>(k1.data()), 4);
  CHECK(raw->GetGuidedSiftMatch(8, buf, nullptr, Ff, 0.7f, 0.8f, 16.f, 16.f, 1) == 2);
  CHECK(raw->GetGuidedSiftMatch(8, buf, Hf, nullptr, 0.7f, 0.8f, 16.f, 16.f, 1) == 1);     // H = I rejects the moved keypoint
  CHECK(raw->GetGuidedSiftMatch(8, buf, Hf, Ff) == -1);                                   // both constraints: unsupported
  SiftMatchGPU::SiftKeypoint sk[2] = {{1, 0, 1, 0}, {2, 0, 1, 0}};
  raw->SetFeatureLocation(0, sk);
  CHECK(raw->GetGuidedSiftMatch(8, buf, Hf, nullptr, 0.7f, 0.8f, 16.f, 16.f, 1) == 2);
  CHECK(raw->GetGuidedSiftMatch(1, buf, Hf, nullptr, 0.7f, 0.8f, 16.f, 16.f, 1) == 1);     // max_match clamps
  delete raw;
  std::printf("guided shim ok\n");
  return 0;
}
