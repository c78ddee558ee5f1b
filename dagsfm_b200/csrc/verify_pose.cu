// Relative pose of verified pairs: the part of
// TwoViewGeometry::EstimateWithRelativePose (src/estimators/two_view_geometry.cc:232-290) that follows
// EstimateCalibrated -- DAGSfM's TwoViewGeometry::Estimate runs it for every pair whose two cameras
// have a prior focal length (:113-126) and stores qvec / tvec / tri_angle with the pair.
//   DecomposeEssentialMatrix / PoseFromEssentialMatrix    src/base/essential_matrix.cc:41-88
//   DecomposeHomographyMatrix / PoseFromHomographyMatrix  src/base/homography_matrix.cc:44-197
//   CheckCheirality                                       src/base/pose.cc:225-248
//   TriangulatePoint, CalculateTriangulationAnglesWithPM  src/base/triangulation.cc:38-51,183-215
//   CalculateDepth                                        src/base/projection.cc:193-197
//   Camera::CalibrationMatrix                             src/base/camera.cc:75-94
//   RotationMatrixToQuaternion = Eigen::Quaterniond(R)    src/base/pose.cc:70-73
//   Median                                                src/util/math.h:212-229
// One warp per pair.  The (at most four) pose candidates are computed redundantly by every lane
// (3x3 SVD / closed form); the per-inlier work -- a 4x4 DLT triangulation by Jacobi SVD, two depth tests
// -- is spread over the lanes, counted with ballots; the surviving points' triangulation angles are
// compacted into a per-pair slice of a scratch array and their median is found by a 63-pass bitwise
// radix selection on the (non-negative) IEEE bit patterns, so no sort and no size limit.
// FP64, --fmad=false: the arithmetic is the oracle's, operation for operation.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "../../include/dagsfm_b200.h"
#include "verify_common.cuh"
#include "verify_solvers.cuh"
#include "camera_models.cuh"

namespace b2 {
namespace vp {

using vf::View;
constexpr unsigned kFull = 0xffffffffu;

__device__ inline double det3(const double* a) {
  return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}

// essential_matrix.cc:41-61.  R1, R2 row-major, t unit.
__device__ __noinline__ void decompose_essential(const double* E, double* R1, double* R2, double* t) {
  double G[9], V[9], sig[3];
  for (int i = 0; i < 9; ++i) G[i] = E[i];
  vf::jacobi_svd<3>(View<1>{G}, 3, 3, View<1>{V}, sig);  // G = E V (columns = sigma_k u_k), sorted
  double U[9], Vt[9];
  for (int k = 0; k < 2; ++k) {
    double n = 0;
    for (int r = 0; r < 3; ++r) n += G[3 * r + k] * G[3 * r + k];
    n = sqrt(n);
    for (int r = 0; r < 3; ++r) U[3 * r + k] = n > 0 ? G[3 * r + k] / n : (r == k ? 1.0 : 0.0);
  }
  // third left singular vector (sigma ~ 0): completes the right-handed frame
  U[2] = U[3] * U[7] - U[6] * U[4];
  U[5] = U[6] * U[1] - U[0] * U[7];
  U[8] = U[0] * U[4] - U[3] * U[1];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Vt[3 * r + c] = V[3 * c + r];
  if (det3(U) < 0)
    for (int i = 0; i < 9; ++i) U[i] = -U[i];
  if (det3(Vt) < 0)
    for (int i = 0; i < 9; ++i) Vt[i] = -Vt[i];
  const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double T[9];
  vf::mat3_mul(U, W, T);
  vf::mat3_mul(T, Vt, R1);
  vf::mat3_mul(U, Wt, T);
  vf::mat3_mul(T, Vt, R2);
  const double n = sqrt(U[2] * U[2] + U[5] * U[5] + U[8] * U[8]);
  t[0] = U[2] / n;
  t[1] = U[5] / n;
  t[2] = U[8] / n;
}

__device__ inline int sign_of(double v) { return (0.0 < v) - (v < 0.0); }
__device__ inline double opposite_of_minor(const double* m, int row, int col) {
  const int col1 = col == 0 ? 1 : 0, col2 = col == 2 ? 1 : 2, row1 = row == 0 ? 1 : 0, row2 = row == 2 ? 1 : 2;
  return m[3 * row1 + col2] * m[3 * row2 + col1] - m[3 * row1 + col1] * m[3 * row2 + col2];
}
__device__ inline void calibration_matrix(const b2_camera& c, double* K) {
  for (int i = 0; i < 9; ++i) K[i] = 0;
  K[8] = 1;
  if (cam::two_focal(c.model)) { K[0] = c.params[0]; K[4] = c.params[1]; K[2] = c.params[2]; K[5] = c.params[3]; }
  else { K[0] = K[4] = c.params[0]; K[2] = c.params[1]; K[5] = c.params[2]; }
}

// homography_matrix.cc:65-168.  Returns the number of candidates (1 for a pure rotation, else 4).
__device__ __noinline__ int decompose_homography(const double* H, const double* K1, const double* K2, double* R /*4x9*/,
                                                  double* t /*4x3*/) {
  double K2i[9], T[9], Hn[9];
  vf::mat3_inverse(K2, K2i);
  vf::mat3_mul(K2i, H, T);
  vf::mat3_mul(T, K1, Hn);
  {
    double G[9], V[9], sig[3];
    for (int i = 0; i < 9; ++i) G[i] = Hn[i];
    vf::jacobi_svd<3>(View<1>{G}, 3, 3, View<1>{V}, sig);
    for (int i = 0; i < 9; ++i) Hn[i] /= sig[1];
  }
  double Ht[9], S[9];
  vf::mat3_transpose(Hn, Ht);
  vf::mat3_mul(Ht, Hn, S);
  S[0] -= 1; S[4] -= 1; S[8] -= 1;
  double inf_norm = 0;
  for (int i = 0; i < 9; ++i) inf_norm = fmax(inf_norm, fabs(S[i]));
  if (inf_norm < 1e-3) {
    for (int i = 0; i < 9; ++i) R[i] = Hn[i];
    t[0] = t[1] = t[2] = 0;
    return 1;
  }
  const double M00 = opposite_of_minor(S, 0, 0), M11 = opposite_of_minor(S, 1, 1), M22 = opposite_of_minor(S, 2, 2);
  const double rtM00 = sqrt(M00), rtM11 = sqrt(M11), rtM22 = sqrt(M22);
  const double M01 = opposite_of_minor(S, 0, 1), M12 = opposite_of_minor(S, 1, 2), M02 = opposite_of_minor(S, 0, 2);
  const int e12 = sign_of(M12), e02 = sign_of(M02), e01 = sign_of(M01);
  const double nS[3] = {fabs(S[0]), fabs(S[4]), fabs(S[8])};
  int idx = 0;
  for (int k = 1; k < 3; ++k)
    if (nS[k] > nS[idx]) idx = k;
  double np1[3], np2[3];
  if (idx == 0) {
    np1[0] = S[0]; np2[0] = S[0];
    np1[1] = S[1] + rtM22; np2[1] = S[1] - rtM22;
    np1[2] = S[2] + e12 * rtM11; np2[2] = S[2] - e12 * rtM11;
  } else if (idx == 1) {
    np1[0] = S[1] + rtM22; np2[0] = S[1] - rtM22;
    np1[1] = S[4]; np2[1] = S[4];
    np1[2] = S[5] - e02 * rtM00; np2[2] = S[5] + e02 * rtM00;
  } else {
    np1[0] = S[2] + e01 * rtM11; np2[0] = S[2] - e01 * rtM11;
    np1[1] = S[5] + rtM00; np2[1] = S[5] - rtM00;
    np1[2] = S[8]; np2[2] = S[8];
  }
  const double traceS = S[0] + S[4] + S[8];
  const double v = 2.0 * sqrt(1.0 + traceS - M00 - M11 - M22);
  const double ESii = sign_of(S[4 * idx]);
  const double r_2 = 2 + traceS + v, nt_2 = 2 + traceS - v;
  const double r = sqrt(r_2), n_t = sqrt(nt_2);
  double n1[3], n2[3];
  {
    const double a = sqrt(np1[0] * np1[0] + np1[1] * np1[1] + np1[2] * np1[2]);
    const double b = sqrt(np2[0] * np2[0] + np2[1] * np2[1] + np2[2] * np2[2]);
    for (int k = 0; k < 3; ++k) { n1[k] = np1[k] / a; n2[k] = np2[k] / b; }
  }
  const double half_nt = 0.5 * n_t, esii_t_r = ESii * r;
  double t1s[3], t2s[3];
  for (int k = 0; k < 3; ++k) {
    t1s[k] = half_nt * (esii_t_r * n2[k] - n_t * n1[k]);
    t2s[k] = half_nt * (esii_t_r * n1[k] - n_t * n2[k]);
  }
  double B[9], R1[9], R2[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) B[3 * a + b] = (a == b ? 1.0 : 0.0) - (2.0 / v) * t1s[a] * n1[b];
  vf::mat3_mul(Hn, B, R1);
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) B[3 * a + b] = (a == b ? 1.0 : 0.0) - (2.0 / v) * t2s[a] * n2[b];
  vf::mat3_mul(Hn, B, R2);
  for (int i = 0; i < 9; ++i) { R[i] = R1[i]; R[9 + i] = R1[i]; R[18 + i] = R2[i]; R[27 + i] = R2[i]; }
  for (int a = 0; a < 3; ++a) {
    const double u = R1[3 * a] * t1s[0] + R1[3 * a + 1] * t1s[1] + R1[3 * a + 2] * t1s[2];
    const double w = R2[3 * a] * t2s[0] + R2[3 * a + 1] * t2s[1] + R2[3 * a + 2] * t2s[2];
    t[a] = u; t[3 + a] = -u; t[6 + a] = w; t[9 + a] = -w;
  }
  return 4;
}

// triangulation.cc:38-51 with proj_matrix1 = [I | 0], proj_matrix2 = [R | t]
__device__ __noinline__ void triangulate_point(const double* R, const double* t, double2 p1, double2 p2, double* X) {
  double A[16], V[16], sig[4];
  const double P1[3][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
  double P2[3][4];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) P2[r][c] = R[3 * r + c];
    P2[r][3] = t[r];
  }
  for (int c = 0; c < 4; ++c) {
    A[c] = p1.x * P1[2][c] - P1[0][c];
    A[4 + c] = p1.y * P1[2][c] - P1[1][c];
    A[8 + c] = p2.x * P2[2][c] - P2[0][c];
    A[12 + c] = p2.y * P2[2][c] - P2[1][c];
  }
  vf::jacobi_svd<4>(View<1>{A}, 4, 4, View<1>{V}, sig);
  for (int r = 0; r < 3; ++r) X[r] = V[4 * r + 3] / V[15];
}

// one point of CheckCheirality (pose.cc:225-248); max_depth and |third column of [R | t]| are per candidate
__device__ inline bool cheirality(const double* R, const double* t, double max_depth, double n2, double2 p1, double2 p2,
                                  double* X) {
  triangulate_point(R, t, p1, p2, X);
  const double kMinDepth = DBL_EPSILON;
  const double depth1 = X[2] * 1.0;
  if (depth1 > kMinDepth && depth1 < max_depth) {
    const double depth2 = (R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2]) * n2;
    if (depth2 > kMinDepth && depth2 < max_depth) return true;
  }
  return false;
}

// Eigen::Quaterniond(rot_mat) -> (w, x, y, z)
__device__ inline void rotation_to_quaternion(const double* m, double* q) {
  const double tr = m[0] + m[4] + m[8];
  if (tr > 0) {
    double t = sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}

// k-th smallest (0-based) of the n non-negative doubles a[0..n): bitwise radix selection, whole warp.
__device__ inline double warp_select(const double* a, int n, int k, int lane) {
  unsigned long long prefix = 0, mask = 0;
  for (int bit = 62; bit >= 0; --bit) {
    const unsigned long long b = 1ull << bit;
    int cnt0 = 0;
    for (int i = lane; i < n; i += 32) {
      const unsigned long long key = (unsigned long long)__double_as_longlong(a[i]);
      cnt0 += ((key & mask) == prefix && !(key & b)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt0 += __shfl_xor_sync(kFull, cnt0, o);
    if (k >= cnt0) {
      k -= cnt0;
      prefix |= b;
    }
    mask |= b;
  }
  return __longlong_as_double((long long)prefix);
}

struct PoseArgs {
  const b2_camera* cams;
  const int64_t* img_off;
  int32_t n_images;
  const double2* nxy;
  int64_t n_pairs;
  const uint32_t* pairs;
  const int64_t* match_off;
  const b2_two_view_result* results;
  const uint32_t* inliers;
  b2_relative_pose* poses;
  double* angles;  // [match_off[n_pairs]] scratch: the pair's slice starts at match_off[p]
  int* err;
};

__global__ void __launch_bounds__(128) relative_pose_kernel(PoseArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < a.n_pairs; p += n_warps) {
    const b2_two_view_result& res = a.results[p];
    const uint32_t i1 = a.pairs[2 * p], i2 = a.pairs[2 * p + 1];
    b2_relative_pose out;
    out.qvec[0] = out.qvec[1] = out.qvec[2] = out.qvec[3] = 0;  // TwoViewGeometry() leaves qvec = tvec = 0 (two_view_geometry.h:159-166)
    out.tvec[0] = out.tvec[1] = out.tvec[2] = 0;
    out.tri_angle = 0;
    out.config = res.config;
    out.n_points3D = 0;
    bool run = i1 < (uint32_t)a.n_images && i2 < (uint32_t)a.n_images;
    if (!run && lane == 0) atomicExch(a.err, 1);
    // TwoViewGeometry::Estimate: only pairs of two prior-focal-length cameras take the relative-pose path; a
    // DEGENERATE pair has no inliers (the reference decomposes an unset H into NaNs nobody reads); MULTIPLE
    // comes from EstimateMultiple, which never computes poses.
    if (run) run = a.cams[i1].has_prior_focal_length && a.cams[i2].has_prior_focal_length;
    const int cfg = res.config;
    // 4 / 5 (PLANAR / PANORAMIC) only occur when TwoViewGeometry::EstimateRelativePose (:169-230) is re-run on a stored geometry
    if (cfg != 2 && cfg != 3 && cfg != 4 && cfg != 5 && cfg != 6 && cfg != 7) run = false;
    if (!run) {
      if (lane == 0) a.poses[p] = out;
      continue;
    }
    const int n = res.n_inliers;
    const uint32_t* inl = a.inliers + 2 * a.match_off[p];
    const double2* nx1 = a.nxy + a.img_off[i1];
    const double2* nx2 = a.nxy + a.img_off[i2];
    {  // the inlier list must fit the pair's slot and name keypoints of the two images (device-chained callers
       // cannot be checked on the host): an inconsistent pair gets the default pose and raises the error flag
      bool bad = n < 0 || (int64_t)n > a.match_off[p + 1] - a.match_off[p];
      if (!bad) {
        const uint64_t np1 = (uint64_t)(a.img_off[i1 + 1] - a.img_off[i1]), np2 = (uint64_t)(a.img_off[i2 + 1] - a.img_off[i2]);
        for (int i = lane; i < n; i += 32) bad |= inl[2 * i] >= np1 || inl[2 * i + 1] >= np2;
      }
      if (__any_sync(kFull, bad)) {
        if (lane == 0) {
          atomicExch(a.err, 2);
          a.poses[p] = out;
        }
        continue;
      }
    }
    double Rc[36], tc[12];
    int nc;
    if (cfg == 2 || cfg == 3) {
      double R1[9], R2[9], t[3];
      decompose_essential(res.E, R1, R2, t);
      for (int i = 0; i < 9; ++i) { Rc[i] = R1[i]; Rc[9 + i] = R2[i]; Rc[18 + i] = R1[i]; Rc[27 + i] = R2[i]; }
      for (int i = 0; i < 3; ++i) { tc[i] = t[i]; tc[3 + i] = t[i]; tc[6 + i] = -t[i]; tc[9 + i] = -t[i]; }
      nc = 4;
    } else {
      double K1[9], K2[9];
      calibration_matrix(a.cams[i1], K1);
      calibration_matrix(a.cams[i2], K2);
      nc = decompose_homography(res.H, K1, K2, Rc, tc);
    }
    // the candidate with the most points in front of both cameras; ties keep the later one (">=")
    int best = 0, best_count = 0;
    for (int c = 0; c < nc; ++c) {
      const double* R = Rc + 9 * c;
      const double* t = tc + 3 * c;
      double rt[3];
      for (int k = 0; k < 3; ++k) rt[k] = R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2];
      const double max_depth = 1000.0f * sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2]);
      const double n2 = sqrt(R[2] * R[2] + R[5] * R[5] + R[8] * R[8]);
      int cnt = 0;
      for (int i = lane; i < n; i += 32) {
        double X[3];
        cnt += cheirality(R, t, max_depth, n2, nx1[inl[2 * i]], nx2[inl[2 * i + 1]], X) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
      if (cnt >= best_count) { best = c; best_count = cnt; }
    }
    const double* R = Rc + 9 * best;
    const double* t = tc + 3 * best;
    rotation_to_quaternion(R, out.qvec);
    out.tvec[0] = t[0]; out.tvec[1] = t[1]; out.tvec[2] = t[2];
    out.n_points3D = best_count;
    if (best_count > 0) {
      // triangulation angles of the surviving points, compacted in match order
      double rt[3], c2[3];
      for (int k = 0; k < 3; ++k) rt[k] = R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2];
      const double max_depth = 1000.0f * sqrt(rt[0] * rt[0] + rt[1] * rt[1] + rt[2] * rt[2]);
      const double n2 = sqrt(R[2] * R[2] + R[5] * R[5] + R[8] * R[8]);
      for (int k = 0; k < 3; ++k) c2[k] = -rt[k];
      const double baseline2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
      double* ang = a.angles + a.match_off[p];
      int base = 0;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        double angle = 0;
        if (i < n) {
          double X[3];
          ok = cheirality(R, t, max_depth, n2, nx1[inl[2 * i]], nx2[inl[2 * i + 1]], X);
          if (ok) {
            const double ray1 = sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
            const double d0 = X[0] - c2[0], d1 = X[1] - c2[1], d2 = X[2] - c2[2];
            const double ray2 = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            const double v = fabs(acos((ray1 * ray1 + ray2 * ray2 - baseline2) / (2 * ray1 * ray2)));
            angle = (v != v) ? 0.0 : fmin(v, M_PI - v);
          }
        }
        const unsigned m = __ballot_sync(kFull, ok);
        if (ok) ang[base + __popc(m & ((1u << lane) - 1u))] = angle;
        base += __popc(m);
      }
      __syncwarp();
      const int mid = best_count / 2;
      const double hi = warp_select(ang, best_count, mid, lane);
      out.tri_angle = (best_count % 2 == 0) ? (hi + warp_select(ang, best_count, mid - 1, lane)) / 2.0 : hi;
    }
    if (cfg == 6) {  // PLANAR_OR_PANORAMIC is resolved by the translation (two_view_geometry.cc:282-289)
      if (sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]) == 0) { out.config = 5; out.tri_angle = 0; }
      else out.config = 4;
    }
    if (lane == 0) a.poses[p] = out;
  }
}

}  // namespace vp

cudaError_t launch_relative_pose(const b2_camera* cams, const int64_t* img_off, int n_images, const double* nxy,
                                 int64_t n_pairs, const uint32_t* pairs, const int64_t* match_off,
                                 const b2_two_view_result* results, const uint32_t* inliers, b2_relative_pose* poses,
                                 double* angles, int* err, int n_sm, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  vp::PoseArgs a;
  a.cams = cams; a.img_off = img_off; a.n_images = n_images; a.nxy = (const double2*)nxy;
  a.n_pairs = n_pairs; a.pairs = pairs; a.match_off = match_off; a.results = results; a.inliers = inliers;
  a.poses = poses; a.angles = angles; a.err = err;
  const int64_t blocks = (n_pairs + 3) / 4;
  const int grid = (int)(blocks < (int64_t)n_sm * 8 ? blocks : (int64_t)n_sm * 8);
  vp::relative_pose_kernel<<<grid, 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace b2
