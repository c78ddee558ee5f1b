// Result metrics of a bundle adjustment:
//   CalculateSquaredReprojectionError          src/base/projection.cc:119-136
//   Reconstruction::ComputeMeanReprojectionError  src/base/reconstruction.cc:814-858  (also sets Point3D::Error)
//   QuaternionRotatePoint                      src/base/pose.cc:110-116 (Eigen::Quaterniond * Vector3d on the normalised qvec)
// One warp per point over its track (the CSR rows of the problem): squared error of every observation -- points not in
// front of the camera are skipped exactly as the reference skips its numeric_limits::max() marker --, the track's sum
// of error norms / track length = Point3D::SetError, and per-block partial sums for the mean (summed on the host in
// block order: deterministic).  All eleven camera models through camera_models.cuh.  FP64, --fmad=false.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "ba_common.cuh"
#include "camera_models.cuh"

namespace b2 {
namespace bam {

__global__ void __launch_bounds__(256)
reprojection_error_kernel(int n_pts, const int64_t* __restrict__ pt_start, const int32_t* __restrict__ obs_img,
                          const double2* __restrict__ obs_xy, const int32_t* __restrict__ img_cam,
                          const int32_t* __restrict__ cam_model, const double* __restrict__ cam_params, int cam_stride,
                          const double* __restrict__ qvec, const double* __restrict__ tvec, const double* __restrict__ xyz,
                          double* __restrict__ point_error, double* __restrict__ partial_sum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  double s = 0;
  int L = 0;
  if (p < n_pts) {
    const int64_t o0 = pt_start[p];
    L = (int)(pt_start[p + 1] - o0);
    const double X0 = xyz[3 * p], X1 = xyz[3 * p + 1], X2 = xyz[3 * p + 2];
    for (int a = lane; a < L; a += 32) {
      const int i = obs_img[o0 + a], cm = img_cam[i];
      // NormalizeQuaternion (pose.cc:82-91), then Eigen's q * v: v + w * (2 q_v x v) + q_v x (2 q_v x v)
      double w = qvec[4 * i], x = qvec[4 * i + 1], y = qvec[4 * i + 2], z = qvec[4 * i + 3];
      const double n = sqrt(w * w + x * x + y * y + z * z);
      if (n == 0) { w = 1.0; } else { w /= n; x /= n; y /= n; z /= n; }
      const double u0 = 2.0 * (y * X2 - z * X1), u1 = 2.0 * (z * X0 - x * X2), u2 = 2.0 * (x * X1 - y * X0);
      const double p0 = X0 + w * u0 + (y * u2 - z * u1) + tvec[3 * i];
      const double p1 = X1 + w * u1 + (z * u0 - x * u2) + tvec[3 * i + 1];
      const double p2 = X2 + w * u2 + (x * u1 - y * u0) + tvec[3 * i + 2];
      if (p2 < DBL_EPSILON) continue;  // "point is in front of camera" fails: the reference's max() marker, skipped by the caller
      double px, py;
      cam::world_to_image(cam_model[cm], cam_params + (int64_t)cam_stride * cm, p0 / p2, p1 / p2, &px, &py);
      const double2 xy = obs_xy[o0 + a];
      const double dx = px - xy.x, dy = py - xy.y;
      s += sqrt(dx * dx + dy * dy);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  if (lane == 0) {
    if (p < n_pts && point_error) point_error[p] = L > 0 ? s / L : 0.0;  // Point3D::SetError(reproj_error_sum / Track().Length())
    ws[warp] = (p < n_pts) ? s : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += ws[k];
    partial_sum[blockIdx.x] = t;
  }
}

}  // namespace bam

cudaError_t bam_launch_reprojection_errors(int n_pts, const int64_t* pt_start, const int32_t* obs_img, const double* obs_xy,
                                           const int32_t* img_cam, const int32_t* cam_model, const double* cam_params,
                                           int cam_stride, const double* qvec, const double* tvec, const double* xyz,
                                           double* point_error, double* partial_sum, int* n_blocks, cudaStream_t s) {
  *n_blocks = (int)(((int64_t)n_pts * 32 + 255) / 256);
  if (n_pts == 0) return cudaSuccess;
  bam::reprojection_error_kernel<<<*n_blocks, 256, 0, s>>>(n_pts, pt_start, obs_img, (const double2*)obs_xy, img_cam, cam_model,
                                                          cam_params, cam_stride, qvec, tvec, xyz, point_error, partial_sum);
  return cudaGetLastError();
}

}  // namespace b2
