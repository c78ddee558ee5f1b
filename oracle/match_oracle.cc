// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
// product path (dagsfm_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// CPU restatement of the reference's descriptor matcher, following
//   src/feature/sift.cc:76-109   ComputeSiftDistanceMatrix
//   src/feature/sift.cc:111-162  FindBestMatchesOneWay
//   src/feature/sift.cc:164-198  FindBestMatches
//   src/feature/sift.cc:810-822  MatchSiftFeaturesCPU
//   src/feature/sift.cc:824-875  MatchGuidedSiftFeaturesCPU (+ the guided_filter branch of :96-103)
//   src/feature/utils.cc:47-76   L2NormalizeFeatureDescriptors / FeatureDescriptorsToUnsignedByte
//   src/feature/sift_test.cc:243-253 CreateRandomFeatureDescriptors (the test fixture)
//   src/util/random.{h,cc}       SetPRNGSeed / RandomReal (std::mt19937 +
//                                std::uniform_real_distribution of THIS libstdc++)
// The reference cannot be compiled here (Eigen / glog absent), so this file restates
// it in plain C++; the arithmetic is integer (exact) plus four float32 operations
// per row (kDistNorm*dot, min, acosf, max_ratio*), written exactly as the reference
// writes them.  Parity pins: tests/test_oracle_match.py replays the reference's own
// unit tests sift_test.cc:300-325 and :448-578 (expected match counts 2, 100, 100,
// 98, 99, 100, 98).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

namespace {

// sift.cc:76-109 (guided_filter == nullptr branch).  The reference casts both
// descriptor sets to int matrices and takes row-by-row int dot products into a
// row-major N1 x N2 int matrix.
void ComputeSiftDistanceMatrix(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                               std::vector<int>* dists) {
  std::vector<int> a((size_t)n1 * 128), b((size_t)n2 * 128);
  for (size_t i = 0; i < a.size(); ++i) a[i] = d1[i];
  for (size_t i = 0; i < b.size(); ++i) b[i] = d2[i];
  dists->resize((size_t)n1 * n2);
  for (int i1 = 0; i1 < n1; ++i1) {
    const int* r1 = a.data() + (size_t)i1 * 128;
    for (int i2 = 0; i2 < n2; ++i2) {
      const int* r2 = b.data() + (size_t)i2 * 128;
      int s = 0;
      for (int k = 0; k < 128; ++k) s += r1[k] * r2[k];
      (*dists)[(size_t)i1 * n2 + i2] = s;
    }
  }
}

// sift.cc:111-162.  `dists` is rows x cols, row-major.
size_t FindBestMatchesOneWay(const int* dists, int rows, int cols, const float max_ratio,
                             const float max_distance, std::vector<int>* matches) {
  const float kDistNorm = 1.0f / (512.0f * 512.0f);
  size_t num_matches = 0;
  matches->assign(rows, -1);
  for (int i1 = 0; i1 < rows; ++i1) {
    int best_i2 = -1;
    int best_dist = 0;
    int second_best_dist = 0;
    for (int i2 = 0; i2 < cols; ++i2) {
      const int dist = dists[(size_t)i1 * cols + i2];
      if (dist > best_dist) {
        best_i2 = i2;
        second_best_dist = best_dist;
        best_dist = dist;
      } else if (dist > second_best_dist) {
        second_best_dist = dist;
      }
    }
    if (best_i2 == -1) continue;
    const float best_dist_normed = std::acos(std::min(kDistNorm * best_dist, 1.0f));
    if (best_dist_normed > max_distance) continue;
    const float second_best_dist_normed = std::acos(std::min(kDistNorm * second_best_dist, 1.0f));
    if (best_dist_normed >= max_ratio * second_best_dist_normed) continue;
    num_matches += 1;
    (*matches)[i1] = best_i2;
  }
  return num_matches;
}

// sift.cc:164-198 + 810-822.
void FindBestMatches(const std::vector<int>& dists, int n1, int n2, float max_ratio, float max_distance,
                     int cross_check, std::vector<uint32_t>* out);

int MatchSiftFeaturesCPU(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                         float max_distance, int cross_check, std::vector<uint32_t>* out) {
  out->clear();
  std::vector<int> dists;
  ComputeSiftDistanceMatrix(d1, n1, d2, n2, &dists);
  FindBestMatches(dists, n1, n2, max_ratio, max_distance, cross_check, out);
  return (int)(out->size() / 2);
}

// sift.cc:164-198 FindBestMatches on a row-major n1 x n2 matrix.
void FindBestMatches(const std::vector<int>& dists, int n1, int n2, float max_ratio, float max_distance,
                     int cross_check, std::vector<uint32_t>* out) {
  out->clear();
  std::vector<int> m12;
  FindBestMatchesOneWay(dists.data(), n1, n2, max_ratio, max_distance, &m12);
  if (cross_check) {
    // dists.transpose(): Eigen materialises the transposed (col-major -> row-major) copy
    std::vector<int> tr((size_t)n1 * n2);
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n2; ++j) tr[(size_t)j * n1 + i] = dists[(size_t)i * n2 + j];
    std::vector<int> m21;
    FindBestMatchesOneWay(tr.data(), n2, n1, max_ratio, max_distance, &m21);
    for (size_t i1 = 0; i1 < m12.size(); ++i1) {
      if (m12[i1] != -1 && m21[m12[i1]] != -1 && m21[m12[i1]] == static_cast<int>(i1)) {
        out->push_back((uint32_t)i1);
        out->push_back((uint32_t)m12[i1]);
      }
    }
  } else {
    for (size_t i1 = 0; i1 < m12.size(); ++i1) {
      if (m12[i1] != -1) {
        out->push_back((uint32_t)i1);
        out->push_back((uint32_t)m12[i1]);
      }
    }
  }
}

// sift.cc:824-875 MatchGuidedSiftFeaturesCPU with the guided_filter branch of
// ComputeSiftDistanceMatrix (:96-103): dists(i1, i2) = 0 where the float32 geometric residual of
// (keypoint1[i1], keypoint2[i2]) exceeds max_error^2, else the descriptor dot.  config follows
// TwoViewGeometry::ConfigurationType: 2/3 (CALIBRATED/UNCALIBRATED) -> Sampson error w.r.t. F,
// 4/5/6 (PLANAR/PANORAMIC/PLANAR_OR_PANORAMIC) -> transfer error w.r.t. H, anything else: the
// reference returns without touching inlier_matches (here: returns -1, output empty).
// The filter is evaluated in float32 like the reference's Eigen::Matrix3f expressions, sums
// left to right, no FMA contraction (the reference's own rounding depends on how Eigen was
// compiled, so parity of a residual within one ulp of the threshold is unpinned).
int MatchGuidedSiftFeaturesCPU(const float* kp1 /*[n1][2]*/, const float* kp2, const uint8_t* d1, int n1,
                               const uint8_t* d2, int n2, int config, const double* F_rm, const double* H_rm,
                               double max_error, float max_ratio, float max_distance, int cross_check,
                               std::vector<uint32_t>* out) {
  out->clear();
  const float max_residual = (float)(max_error * max_error);
  float F[9], H[9];
  for (int k = 0; k < 9; ++k) {
    F[k] = (float)F_rm[k];
    H[k] = (float)H_rm[k];
  }
  const bool use_f = (config == 2 || config == 3), use_h = (config == 4 || config == 5 || config == 6);
  if (!use_f && !use_h) return -1;
  std::vector<int> dists;
  ComputeSiftDistanceMatrix(d1, n1, d2, n2, &dists);
  for (int i1 = 0; i1 < n1; ++i1) {
    const float x1 = kp1[2 * i1], y1 = kp1[2 * i1 + 1];
    for (int i2 = 0; i2 < n2; ++i2) {
      const float x2 = kp2[2 * i2], y2 = kp2[2 * i2 + 1];
      bool skip;
      if (use_f) {
        const float Fx1_0 = F[0] * x1 + F[1] * y1 + F[2] * 1.0f;
        const float Fx1_1 = F[3] * x1 + F[4] * y1 + F[5] * 1.0f;
        const float Fx1_2 = F[6] * x1 + F[7] * y1 + F[8] * 1.0f;
        const float Ftx2_0 = F[0] * x2 + F[3] * y2 + F[6] * 1.0f;
        const float Ftx2_1 = F[1] * x2 + F[4] * y2 + F[7] * 1.0f;
        const float x2tFx1 = x2 * Fx1_0 + y2 * Fx1_1 + 1.0f * Fx1_2;
        skip = x2tFx1 * x2tFx1 / (Fx1_0 * Fx1_0 + Fx1_1 * Fx1_1 + Ftx2_0 * Ftx2_0 + Ftx2_1 * Ftx2_1) > max_residual;
      } else {
        const float h0 = H[0] * x1 + H[1] * y1 + H[2] * 1.0f;
        const float h1 = H[3] * x1 + H[4] * y1 + H[5] * 1.0f;
        const float h2 = H[6] * x1 + H[7] * y1 + H[8] * 1.0f;
        const float e0 = h0 / h2 - x2, e1 = h1 / h2 - y2;
        skip = e0 * e0 + e1 * e1 > max_residual;
      }
      if (skip) dists[(size_t)i1 * n2 + i2] = 0;
    }
  }
  FindBestMatches(dists, n1, n2, max_ratio, max_distance, cross_check, out);
  return (int)(out->size() / 2);
}

}  // namespace

extern "C" {

// MatchSiftFeaturesCPU.  out: [cap][2] uint32.  Returns #matches (or -needed if cap too small).
int orc_match_sift(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                   float max_distance, int cross_check, uint32_t* out, int cap) {
  std::vector<uint32_t> m;
  const int n = MatchSiftFeaturesCPU(d1, n1, d2, n2, max_ratio, max_distance, cross_check, &m);
  if (n > cap) return -n;
  if (n > 0) memcpy(out, m.data(), (size_t)n * 2 * sizeof(uint32_t));
  return n;
}

// MatchGuidedSiftFeaturesCPU.  Returns #matches, -1 when the configuration has no guided filter,
// INT_MIN + needed if cap is too small.
int orc_match_guided(const float* kp1, const float* kp2, const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                     int config, const double* F, const double* H, double max_error, float max_ratio,
                     float max_distance, int cross_check, uint32_t* out, int cap) {
  std::vector<uint32_t> m;
  const int n = MatchGuidedSiftFeaturesCPU(kp1, kp2, d1, n1, d2, n2, config, F, H, max_error, max_ratio,
                                           max_distance, cross_check, &m);
  if (n < 0) return -1;
  if (n > cap) return -2147483647 + n;
  if (n > 0) memcpy(out, m.data(), (size_t)n * 2 * sizeof(uint32_t));
  return n;
}

// One-way result (for kernel-level tests): matches[rows], -1 = none.
void orc_best_one_way(const uint8_t* d1, int n1, const uint8_t* d2, int n2, float max_ratio,
                      float max_distance, int* matches) {
  std::vector<int> dists, m;
  ComputeSiftDistanceMatrix(d1, n1, d2, n2, &dists);
  FindBestMatchesOneWay(dists.data(), n1, n2, max_ratio, max_distance, &m);
  if (n1 > 0) memcpy(matches, m.data(), (size_t)n1 * sizeof(int));
}

// sift_test.cc:243-253 CreateRandomFeatureDescriptors: SetPRNGSeed(0) (a fresh
// std::mt19937(0), random.cc:40-55), d = pow(U(0,1), 2) (std::pow(float,int) promotes to
// double), L2-normalise each row in float (utils.cc:47-50), round(512*d) saturated to
// uint8 (utils.cc:65-76).  Deviation: Eigen's rowwise().normalized() reduces the squared
// norm with SIMD partial sums; here the float sum is sequential (<= 1 ulp of the norm).
void orc_create_random_descriptors(int n, unsigned seed, uint8_t* out) {
  std::mt19937 prng(seed);
  std::vector<float> row(128);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < 128; ++j) {
      std::uniform_real_distribution<float> distribution(0.0f, 1.0f);
      row[j] = (float)std::pow(distribution(prng), 2);
    }
    float sq = 0.f;
    for (int j = 0; j < 128; ++j) sq += row[j] * row[j];
    const float norm = std::sqrt(sq);
    for (int j = 0; j < 128; ++j) {
      const float v = (sq > 0.f) ? row[j] / norm : row[j];
      const float scaled = std::round(512.0f * v);
      out[(size_t)i * 128 + j] = (uint8_t)std::min(std::max(scaled, 0.0f), 255.0f);
    }
  }
}

// utils.cc:47-76 on one float row (used by the ratio-test fixture of sift_test.cc:529-543).
void orc_l2_normalize_to_u8(const float* in, uint8_t* out) {
  float sq = 0.f;
  for (int j = 0; j < 128; ++j) sq += in[j] * in[j];
  const float norm = std::sqrt(sq);
  for (int j = 0; j < 128; ++j) {
    const float v = (sq > 0.f) ? in[j] / norm : in[j];
    const float scaled = std::round(512.0f * v);
    out[j] = (uint8_t)std::min(std::max(scaled, 0.0f), 255.0f);
  }
}

// CPU baseline driver: the reference runs `num_threads` SiftCPUFeatureMatcher threads, each
// pulling one pair at a time from a shared queue (src/feature/matching.cc:336-357,640-644).
// desc: n_images pointers; pairs: [n_pairs][2].  counts[n_pairs] receives #matches.
// Returns wall seconds.
double orc_match_pairs_mt(const uint8_t* const* desc, const int* n_desc, const uint32_t* pairs,
                          long n_pairs, float max_ratio, float max_distance, int cross_check,
                          int n_threads, int* counts, uint32_t* checksum) {
  std::atomic<long> next(0);
  std::atomic<uint32_t> cs(0);
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) {
    th.emplace_back([&]() {
      std::vector<uint32_t> m;
      for (;;) {
        const long p = next.fetch_add(1);
        if (p >= n_pairs) break;
        const uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
        const int n = MatchSiftFeaturesCPU(desc[i1], n_desc[i1], desc[i2], n_desc[i2], max_ratio,
                                           max_distance, cross_check, &m);
        if (counts) counts[p] = n;
        uint32_t h = 0;
        for (size_t k = 0; k < m.size(); ++k) h = h * 1000003u + m[k];
        cs.fetch_add(h * (uint32_t)(p + 1));
      }
    });
  }
  for (auto& x : th) x.join();
  const auto t1 = std::chrono::steady_clock::now();
  if (checksum) *checksum = cs.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
