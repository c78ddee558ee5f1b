// Replays the shape of the reference's TestMatchSiftFeaturesCPU / ...CPUvsGPU
// (src/feature/sift_test.cc:300-325, 496-505) through the C++ shim with a stand-in for
// Eigen's row-major uint8 matrix.  Built by tests/test_shim_cpp.py; exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dagsfm_b200/colmap_shim.hpp"

struct Desc {  // minimal FeatureDescriptors stand-in
  std::vector<unsigned char> v;
  long n = 0;
  long rows() const { return n; }
  long cols() const { return 128; }
  const unsigned char* data() const { return v.data(); }
};

static Desc make(int n, unsigned seed) {
  Desc d;
  d.n = n;
  d.v.resize((size_t)n * 128);
  unsigned s = seed * 2654435761u + 12345u;
  for (int i = 0; i < n; ++i) {
    double norm = 0;
    double r[128];
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
  opt.max_num_matches = 1000;
  SiftMatchGPU gpu;
  CHECK(CreateSiftGPUMatcher(opt, &gpu));
  Desc d1 = make(100, 1), d2 = d1, empty;
  for (int i = 0; i < 100; ++i) memcpy(&d2.v[(size_t)i * 128], &d1.v[(size_t)(99 - i) * 128], 128);  // reversed rows
  FeatureMatches m;
  MatchSiftFeaturesGPU(opt, &d1, &d2, &gpu, &m);
  CHECK(m.size() == 100);
  for (size_t i = 0; i < m.size(); ++i) CHECK(m[i].point2D_idx1 == i && m[i].point2D_idx2 == 99 - i);
  MatchSiftFeaturesGPU<Desc, FeatureMatches>(opt, nullptr, nullptr, &gpu, &m);  // reuse previous upload
  CHECK(m.size() == 100);
  MatchSiftFeaturesGPU(opt, &empty, &d2, &gpu, &m);
  CHECK(m.empty());
  MatchSiftFeaturesGPU(opt, &d1, &empty, &gpu, &m);
  CHECK(m.empty());
  // clamping to max_num_matches (SiftMatchCU.cpp:108)
  opt.max_num_matches = 64;
  SiftMatchGPU gpu2;
  CHECK(CreateSiftGPUMatcher(opt, &gpu2));
  MatchSiftFeaturesGPU(opt, &d1, &d1, &gpu2, &m);
  CHECK(m.size() == 64);
  std::printf("shim ok\n");
  return 0;
}
