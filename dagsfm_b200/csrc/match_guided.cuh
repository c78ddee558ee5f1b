// Guided descriptor matching: MatchGuidedSiftFeaturesCPU / ...GPU of the reference
// (src/feature/sift.cc:824-875, :987-1066) -- the distance matrix entry of (i1, i2) is forced to 0
// when the float32 geometric residual of (keypoint1[i1], keypoint2[i2]) under the pair's F or H
// exceeds max_error^2 (sift.cc:96-103), then FindBestMatches runs unchanged.
//
// Guided matching runs on verified pairs only and the filter leaves a few candidates per row (an
// epipolar band or a disc), so the dense tensor-core contraction would be >95 % wasted work: one
// THREAD owns one row, evaluates the filter for every column (10 float ops) and takes the exact
// 128-byte dot (dp4a) only for the columns that pass.  Columns are visited in ascending order with
// the reference's strict-'>' update, so ties resolve exactly as in FindBestMatchesOneWay
// (sift.cc:121-133).  The row function below is plain C++: tests/cpp/host_guided.cc compiles it for
// the host and checks it against the oracle without a GPU.  TU built with --fmad=false (the filter
// must be the same IEEE float operations as the oracle's).
#pragma once
#include <cstdint>

#ifndef B2_HD
#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif
#endif

namespace b2 {

struct GuidedGeom {  // per pair
  int32_t kind;      // 0: no guided filter for this configuration, 1: F (Sampson), 2: H (transfer)
  float m[9];        // F or H, row-major, cast from double like the reference's .cast<float>()
};

// TwoViewGeometry::ConfigurationType -> filter (sift.cc:838-866): CALIBRATED / UNCALIBRATED use F,
// PLANAR / PANORAMIC / PLANAR_OR_PANORAMIC use H, anything else has no guided filter.
B2_HD GuidedGeom make_guided_geom(int config, const double* F_rowmajor, const double* H_rowmajor) {
  GuidedGeom g;
  g.kind = (config == 2 || config == 3) ? 1 : (config == 4 || config == 5 || config == 6) ? 2 : 0;
  const double* src = (g.kind == 1) ? F_rowmajor : H_rowmajor;
  for (int k = 0; k < 9; ++k) g.m[k] = (g.kind != 0 && src) ? (float)src[k] : 0.0f;
  return g;
}

// `row_is_image1`: the matrix is dists(i1, i2); in the reverse direction (cross-check) the row is
// an image-2 keypoint and the columns are image-1 keypoints, the filter's argument order stays.
B2_HD bool guided_skip(const GuidedGeom& g, float max_residual, float x1, float y1, float x2, float y2) {
  const float* M = g.m;
  if (g.kind == 1) {
    const float Fx1_0 = M[0] * x1 + M[1] * y1 + M[2] * 1.0f;
    const float Fx1_1 = M[3] * x1 + M[4] * y1 + M[5] * 1.0f;
    const float Fx1_2 = M[6] * x1 + M[7] * y1 + M[8] * 1.0f;
    const float Ftx2_0 = M[0] * x2 + M[3] * y2 + M[6] * 1.0f;
    const float Ftx2_1 = M[1] * x2 + M[4] * y2 + M[7] * 1.0f;
    const float x2tFx1 = x2 * Fx1_0 + y2 * Fx1_1 + 1.0f * Fx1_2;
    return x2tFx1 * x2tFx1 / (Fx1_0 * Fx1_0 + Fx1_1 * Fx1_1 + Ftx2_0 * Ftx2_0 + Ftx2_1 * Ftx2_1) > max_residual;
  }
  const float h0 = M[0] * x1 + M[1] * y1 + M[2] * 1.0f;
  const float h1 = M[3] * x1 + M[4] * y1 + M[5] * 1.0f;
  const float h2 = M[6] * x1 + M[7] * y1 + M[8] * 1.0f;
  const float e0 = h0 / h2 - x2, e1 = h1 / h2 - y2;
  return e0 * e0 + e1 * e1 > max_residual;
}

// Exact dot of two 128-byte descriptors given as 32 little-endian words each.
B2_HD int dot128(const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
  unsigned u = 0;  // unsigned bytes: at most 128 * 255 * 255 < 2^31
#pragma unroll
  for (int k = 0; k < 32; ++k) u = __dp4a(a[k], b[k], u);
  return (int)u;
#else
  int s = 0;
  for (int k = 0; k < 32; ++k)
    for (int q = 0; q < 4; ++q) s += (int)((a[k] >> (8 * q)) & 255u) * (int)((b[k] >> (8 * q)) & 255u);
  return s;
#endif
}

// FindBestMatchesOneWay for ONE row of the guided distance matrix.  xdesc: the row's descriptor
// (32 words); (rx, ry): its keypoint; ydesc / ykp: the column image's descriptors (128 B rows) and
// keypoints (x, y pairs), n_y of them.  thr_dist / ratio_lim / dot_clamp: the integer forms of the
// max_distance and max_ratio tests (match_api.cu: ThresholdTables).  Returns the matched column or -1.
B2_HD int guided_row_best(const GuidedGeom& g, float max_residual, bool row_is_image1, const uint32_t* xdesc,
                          float rx, float ry, const uint8_t* ydesc, const float* ykp, int n_y, int thr_dist,
                          const int* ratio_lim, int dot_clamp) {
  int best = 0, second = 0, best_j = -1;
  for (int j = 0; j < n_y; ++j) {
    const float cx = ykp[2 * j], cy = ykp[2 * j + 1];
    const bool skip = row_is_image1 ? guided_skip(g, max_residual, rx, ry, cx, cy)
                                    : guided_skip(g, max_residual, cx, cy, rx, ry);
    if (skip) continue;  // dists(i1, i2) = 0: can be neither best nor second-best (both start at 0, strict '>')
    const int d = dot128(xdesc, reinterpret_cast<const uint32_t*>(ydesc + (size_t)j * 128));
    if (d > best) {
      best_j = j;
      second = best;
      best = d;
    } else if (d > second) {
      second = d;
    }
  }
  if (best_j < 0 || best < thr_dist) return -1;
  if (second > ratio_lim[best < dot_clamp ? best : dot_clamp]) return -1;
  return best_j;
}

}  // namespace b2
