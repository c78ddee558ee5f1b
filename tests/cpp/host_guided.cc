// Host build of the guided-matching row function (match_guided.cuh is plain C++) driven the way
// the CUDA kernel drives it: one call per row and direction, then the cross-check of
// match_cross_kernel.  Lets the CPU suite compare the device logic, including the integer
// threshold tables of match_thresholds.h, with the oracle.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../dagsfm_b200/csrc/match_guided.cuh"
#include "../../dagsfm_b200/csrc/match_thresholds.h"

extern "C" int host_guided_match(const float* kp1, const float* kp2, const uint8_t* d1, int n1, const uint8_t* d2,
                                 int n2, int config, const double* F, const double* H, double max_error,
                                 float max_ratio, float max_distance, int cross_check, uint32_t* out, int cap) {
  const int kDotClamp = 262144;
  static b2::HostThresholds th;
  if (!th.build(max_ratio, max_distance, kDotClamp)) return -3;
  const b2::GuidedGeom g = b2::make_guided_geom(config, F, H);
  if (g.kind == 0) return -1;
  const float max_residual = (float)(max_error * max_error);
  std::vector<int> m12(n1, -1), m21(n2, -1);
  for (int i = 0; i < n1; ++i) {
    uint32_t x[32];
    memcpy(x, d1 + (size_t)i * 128, 128);
    m12[i] = b2::guided_row_best(g, max_residual, true, x, kp1[2 * i], kp1[2 * i + 1], d2, kp2, n2, th.thr_dist,
                                 th.ratio_lim.data(), kDotClamp);
  }
  for (int j = 0; j < n2; ++j) {
    uint32_t x[32];
    memcpy(x, d2 + (size_t)j * 128, 128);
    m21[j] = b2::guided_row_best(g, max_residual, false, x, kp2[2 * j], kp2[2 * j + 1], d1, kp1, n1, th.thr_dist,
                                 th.ratio_lim.data(), kDotClamp);
  }
  int n = 0;
  for (int i = 0; i < n1; ++i) {
    const int j = m12[i];
    bool ok = j >= 0;
    if (ok && cross_check) ok = (m21[j] == i);
    if (!ok) continue;
    if (n < cap) { out[2 * n] = (uint32_t)i; out[2 * n + 1] = (uint32_t)j; }
    ++n;
  }
  return n;
}

// Unguided use of the same tables: FindBestMatchesOneWay decisions from (best, second) pairs.
extern "C" int host_threshold_accept(float max_ratio, float max_distance, int best, int second) {
  static b2::HostThresholds th;
  static float r = -1, d = -1;
  if (r != max_ratio || d != max_distance) { th.build(max_ratio, max_distance, 262144); r = max_ratio; d = max_distance; }
  if (best <= 0 || best < th.thr_dist) return 0;
  return second <= th.ratio_lim[best < 262144 ? best : 262144];
}
