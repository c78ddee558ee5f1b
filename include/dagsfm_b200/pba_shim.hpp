// dagsfm_b200/pba_shim.hpp -- the subset of the vendored PBA interface the reference drives
// (`pba::ParallelBA`, `pba::CameraT`, `pba::Point3D`, `pba::Point2D`, `pba::ConfigBA`) on top of
// b2_ba_solve, so that ParallelBundleAdjuster::Solve (src/optim/bundle_adjustment.cc:548-628) and
// its SetUp/TearDown (:654-772) compile unchanged against this header instead of lib/PBA/pba.h.
//
//   reference                                         here
//   pba::ParallelBA pba(device, num_threads)          gpu index = device if >= 0, else 0
//   SetNextBundleMode / EnableRadialDistortion        BUNDLE_FULL + PBA_PROJECTION_DISTORTION only
//   SetFixedIntrinsics(bool)                          refine_focal_length = refine_extra_params = !fixed
//   GetInternalConfig()->__lm_max_iteration           b2_ba_options::max_num_iterations
//   SetCameraData / SetPointData / SetProjection      host arrays, updated in place by the solve
//   RunBundleAdjustment()                             b2_ba_solve; returns LM iterations, -1 on error
//   GetInitialMSE / GetFinalMSE / GetIterationsLM     from b2_ba_summary (MSE = sum r^2 / #projections)
//
// PBA's camera is K[R|t] with K = diag(f, f, 1), measurements already centred on the principal
// point and one projection-distortion coefficient: x = f (1 + k r^2) (X/Z, Y/Z) -- exactly the
// reference's SIMPLE_RADIAL model with cx = cy = 0 (camera_models.h:714-757), which is how
// ParallelBundleAdjuster::AddImagesToProblem fills it (:694-700).  Intrinsics are never shared.
// The solve itself is the double-precision LM + exact Schur step of b2_ba_solve (PBA: float storage,
// preconditioned CG); parameters are converted float <-> double at the boundary like PBA_CPU_DOUBLE.
//
// Header-only; link libdagsfm_b200.so.
#ifndef DAGSFM_B200_PBA_SHIM_HPP_
#define DAGSFM_B200_PBA_SHIM_HPP_

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {
namespace pba {

template <class FT>
struct CameraT_ {
  typedef FT float_t;
  float_t f;
  float_t t[3];
  float_t m[3][3];
  float_t radial;
  int distortion_type;
  float_t constant_camera;  // 0 variable, 1 constant, 2 fixed intrinsics (PBA's encoding)

  CameraT_() : f(0), radial(0), distortion_type(0), constant_camera(0) {
    for (int i = 0; i < 3; ++i) {
      t[i] = 0;
      for (int j = 0; j < 3; ++j) m[i][j] = (i == j) ? float_t(1) : float_t(0);
    }
  }
  void SetConstantCamera() { constant_camera = 1; }
  void SetVariableCamera() { constant_camera = 0; }
  void SetFixedIntrinsic() { constant_camera = 2; }
  template <class F> void SetFocalLength(F v) { f = (float_t)v; }
  float_t GetFocalLength() const { return f; }
  template <class F> void SetProjectionDistortion(F r) { radial = (float_t)r; distortion_type = 1; }
  float_t GetProjectionDistortion() const { return distortion_type == 1 ? radial : 0; }
  template <class F> void SetMatrixRotation(const F* r) {
    for (int i = 0, k = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) m[i][j] = (float_t)r[k++];
  }
  template <class F> void GetMatrixRotation(F* r) const {
    for (int i = 0, k = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r[k++] = (F)m[i][j];
  }
  template <class F> void SetTranslation(const F T[3]) { for (int i = 0; i < 3; ++i) t[i] = (float_t)T[i]; }
  template <class F> void GetTranslation(F T[3]) const { for (int i = 0; i < 3; ++i) T[i] = (F)t[i]; }
};
typedef CameraT_<float> CameraT;

template <class FT>
struct Point3D_ {
  FT xyz[3];
  FT reserved;
  template <class F> void SetPoint(const F* p) { for (int i = 0; i < 3; ++i) xyz[i] = (FT)p[i]; }
  template <class F> void SetPoint(F x, F y, F z) { xyz[0] = (FT)x; xyz[1] = (FT)y; xyz[2] = (FT)z; }
  template <class F> void GetPoint(F* p) const { for (int i = 0; i < 3; ++i) p[i] = (F)xyz[i]; }
};
typedef Point3D_<float> Point3D;

struct Point2D {
  float x, y;
  Point2D() : x(0), y(0) {}
  template <class F> Point2D(F X, F Y) : x((float)X), y((float)Y) {}
  template <class F> void SetPoint2D(F X, F Y) { x = (float)X; y = (float)Y; }
  template <class F> void GetPoint2D(F& X, F& Y) const { X = (F)x; Y = (F)y; }
};

// The ConfigBA members the reference touches (bundle_adjustment.cc:583-589, :603-612).
struct ConfigBA {
  int __lm_max_iteration = 50;
  int __cg_min_iteration = 10;
  int __verbose_level = 2;
  float __lm_delta_threshold = 1e-6f;
  float __lm_gradient_threshold = 1e-10f;
  float __lm_mse_threshold = 0.25f;
  float __initial_mse = 0, __final_mse = 0;
  int __num_lm_iteration = 0, __pba_return_code = 0;
  float GetInitialMSE() { return __initial_mse; }
  float GetFinalMSE() { return __final_mse; }
  int GetIterationsLM() { return __num_lm_iteration; }
  int GetBundleReturnCode() { return __pba_return_code; }
};

class ParallelBA {
 public:
  enum DeviceT { PBA_INVALID_DEVICE = -4, PBA_CPU_DOUBLE = -3, PBA_CPU_FLOAT = -2, PBA_CUDA_DEVICE_DEFAULT = -1, PBA_CUDA_DEVICE0 = 0 };
  enum DistortionT { PBA_MEASUREMENT_DISTORTION = -1, PBA_NO_DISTORTION = 0, PBA_PROJECTION_DISTORTION = 1 };
  enum BundleModeT { BUNDLE_FULL = 0, BUNDLE_ONLY_MOTION = 1, BUNDLE_ONLY_STRUCTURE = 2 };

  explicit ParallelBA(DeviceT device = PBA_CUDA_DEVICE_DEFAULT, const int /*num_threads*/ = -1)
      : gpu_(device >= 0 ? (int)device : 0) {}
  ~ParallelBA() {
    if (h_) b2_ba_destroy(h_);
  }
  ParallelBA(const ParallelBA&) = delete;
  ParallelBA& operator=(const ParallelBA&) = delete;

  ConfigBA* GetInternalConfig() { return &cfg_; }
  void SetFixedIntrinsics(bool fixed) { fixed_intrinsics_ = fixed; }
  void SetFocalLengthFixed(bool fixed) { fixed_intrinsics_ = fixed; }
  void EnableRadialDistortion(DistortionT type) { distortion_ = type; }
  void SetNextBundleMode(BundleModeT mode = BUNDLE_FULL) { mode_ = mode; }
  void SetNextTimeBudget(int) {}
  void ReserveStorage(std::size_t, std::size_t, std::size_t) {}
  void SetCameraData(std::size_t ncam, CameraT* cams) { ncam_ = ncam; cams_ = cams; }
  void SetPointData(std::size_t npoint, Point3D* pts) { npt_ = npoint; pts_ = pts; }
  void SetProjection(std::size_t nproj, const Point2D* imgpts, const int* point_idx, const int* cam_idx) {
    nproj_ = nproj; proj_ = imgpts; pidx_ = point_idx; cidx_ = cam_idx;
  }
  float GetMeanSquaredError() { return cfg_.__final_mse; }

  // Returns the number of LM iterations (as PBA does), -1 if the problem cannot be solved here.
  int RunBundleAdjustment() {
    cfg_.__pba_return_code = 'E';
    if (mode_ != BUNDLE_FULL || distortion_ == PBA_MEASUREMENT_DISTORTION) return -1;
    if (!cams_ || !pts_ || !proj_ || !pidx_ || !cidx_ || ncam_ == 0 || npt_ == 0 || nproj_ == 0) return -1;
    const int32_t nc = (int32_t)ncam_, np = (int32_t)npt_;
    std::vector<double> q(4 * ncam_), t(3 * ncam_), par(4 * ncam_), xyz(3 * npt_), oxy(2 * nproj_);
    // PBA_NO_DISTORTION: K = diag(f, f, 1) only = SIMPLE_PINHOLE (0); otherwise SIMPLE_RADIAL (2)
    std::vector<int32_t> icam(ncam_), model(ncam_, distortion_ == PBA_PROJECTION_DISTORTION ? 2 : 0), oimg(nproj_), opt(nproj_);
    std::vector<uint8_t> cpose(ncam_), ctvec(ncam_, 0), ccam(ncam_), cpt(npt_, 0);
    for (int32_t i = 0; i < nc; ++i) {
      const CameraT& c = cams_[i];
      double R[9];
      c.GetMatrixRotation(R);
      rotation_to_quaternion(R, &q[4 * i]);
      for (int k = 0; k < 3; ++k) t[3 * i + k] = c.t[k];
      par[4 * i] = c.f; par[4 * i + 1] = 0; par[4 * i + 2] = 0;
      par[4 * i + 3] = (distortion_ == PBA_PROJECTION_DISTORTION) ? (double)c.GetProjectionDistortion() : 0.0;
      icam[i] = i;
      cpose[i] = c.constant_camera == 1;
      ccam[i] = (c.constant_camera == 1 || c.constant_camera == 2) ? 1 : 0;
    }
    for (int32_t j = 0; j < np; ++j)
      for (int k = 0; k < 3; ++k) xyz[3 * (std::size_t)j + k] = pts_[j].xyz[k];
    for (std::size_t k = 0; k < nproj_; ++k) {
      if (cidx_[k] < 0 || cidx_[k] >= nc || pidx_[k] < 0 || pidx_[k] >= np) return -1;
      if (k > 0 && pidx_[k] < pidx_[k - 1]) return -1;  // PBA too wants the tracks stored contiguously (:655-657)
      oimg[k] = cidx_[k]; opt[k] = pidx_[k];
      oxy[2 * k] = proj_[k].x; oxy[2 * k + 1] = proj_[k].y;
    }
    b2_ba_problem p = {};  // camera_params_stride = 0 means 4 doubles per camera
    p.n_images = nc; p.n_cameras = nc; p.n_points = np; p.n_obs = (int64_t)nproj_;
    p.qvec = q.data(); p.tvec = t.data(); p.image_camera = icam.data();
    p.const_pose = cpose.data(); p.const_tvec = ctvec.data();
    p.camera_model = model.data(); p.camera_params = par.data(); p.const_camera = ccam.data();
    p.xyz = xyz.data(); p.const_point = cpt.data();
    p.obs_image = oimg.data(); p.obs_point = opt.data(); p.obs_xy = oxy.data();
    b2_ba_options o;
    b2_ba_default_options(&o);
    o.max_num_iterations = cfg_.__lm_max_iteration;
    o.refine_focal_length = o.refine_extra_params = fixed_intrinsics_ ? 0 : 1;
    o.refine_principal_point = 0;
    if (!h_ && b2_ba_create(gpu_, &h_) != B2_OK) return -1;
    b2_ba_summary s;
    if (b2_ba_solve(h_, &p, &o, &s) != B2_OK) return -1;
    for (int32_t i = 0; i < nc; ++i) {
      CameraT& c = cams_[i];
      double R[9];
      quaternion_to_rotation(&q[4 * i], R);
      c.SetMatrixRotation(R);
      for (int k = 0; k < 3; ++k) c.t[k] = (float)t[3 * i + k];
      c.f = (float)par[4 * i];
      if (c.distortion_type == 1) c.radial = (float)par[4 * i + 3];
    }
    for (int32_t j = 0; j < np; ++j)
      for (int k = 0; k < 3; ++k) pts_[j].xyz[k] = (float)xyz[3 * (std::size_t)j + k];
    cfg_.__initial_mse = (float)(2.0 * s.initial_cost / (double)nproj_);
    cfg_.__final_mse = (float)(2.0 * s.final_cost / (double)nproj_);
    cfg_.__num_lm_iteration = s.num_iterations;
    cfg_.__pba_return_code = s.termination_type == 0 ? 'G' : 'M';
    return s.num_iterations;
  }

 private:
  // row-major R <-> (w, x, y, z), the convention of base/pose.cc (QuaternionToRotationMatrix)
  static void quaternion_to_rotation(const double* q, double* R) {
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
  }
  static void rotation_to_quaternion(const double* R, double* q) {
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
      const double s = 2 * std::sqrt(tr + 1.0);
      q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
      const double s = 2 * std::sqrt(1.0 + R[0] - R[4] - R[8]);
      q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
      const double s = 2 * std::sqrt(1.0 + R[4] - R[0] - R[8]);
      q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
    } else {
      const double s = 2 * std::sqrt(1.0 + R[8] - R[0] - R[4]);
      q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
    }
  }

  int gpu_;
  b2_ba* h_ = nullptr;
  ConfigBA cfg_;
  bool fixed_intrinsics_ = false;
  DistortionT distortion_ = PBA_NO_DISTORTION;
  BundleModeT mode_ = BUNDLE_FULL;
  std::size_t ncam_ = 0, npt_ = 0, nproj_ = 0;
  CameraT* cams_ = nullptr;
  Point3D* pts_ = nullptr;
  const Point2D* proj_ = nullptr;
  const int *pidx_ = nullptr, *cidx_ = nullptr;
};

// lib/PBA/pba.h:146-152: the factory a caller that loads PBA dynamically resolves, and the version probe (pba.cpp:128-132
// returns 105).  Here they are ordinary inline functions of the adaptor: nothing in the reference's src/ uses the dlopen route.
inline ParallelBA* NewParallelBA(ParallelBA::DeviceT device = ParallelBA::PBA_CUDA_DEVICE_DEFAULT) { return new ParallelBA(device); }
typedef ParallelBA* (*NEWPARALLELBAPROC)(ParallelBA::DeviceT);
inline int ParallelBA_GetVersion() { return 105; }

}  // namespace pba
}  // namespace dagsfm_b200
#endif  // DAGSFM_B200_PBA_SHIM_HPP_
