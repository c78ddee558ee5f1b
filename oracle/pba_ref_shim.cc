// Thin extern "C" driver around the REFERENCE's vendored PBA (lib/PBA, Changchang Wu's
// multicore bundle adjustment), compiled from the reference tree where it lies
// (-I$(REF)/lib/PBA) into oracle/_ref/libpba_ref.so.  No reference source is copied here.
// Configuration mirrors ParallelBundleAdjuster::Solve / AddImagesToProblem
// (reference src/optim/bundle_adjustment.cc:556-626,690-772): PBA_CPU_DOUBLE,
// PBA_PROJECTION_DISTORTION, thresholds /100, cg_min_iteration 10, measurements centred at
// the principal point, rotation passed as a row-major matrix.
// TEST INFRASTRUCTURE ONLY (CPU timing baseline + convergence sanity check: PBA's interface
// structs are float32, DataInterface.h:398-402, so it is not a 1e-6 px parity oracle).
#include <chrono>
#include <cstring>
#include <vector>

#ifdef B2_PBA_SHIM
// The same driver compiled against include/dagsfm_b200/pba_shim.hpp instead of the reference's
// pba.h (tests/test_pba_shim.py): shows that code written for PBA's interface drives b2_ba_solve.
#include "dagsfm_b200/pba_shim.hpp"
namespace pba = dagsfm_b200::pba;
#else
#include "pba.h"
#endif

extern "C" int pba_ref_run(int n_cam, const double* focal, const double* radial, const double* R_rowmajor /*[n][9]*/,
                           const double* t /*[n][3]*/, const unsigned char* cam_const, int n_pts, double* xyz,
                           long n_proj, const double* xy_centered /*[n_proj][2]*/, const int* pt_idx,
                           const int* cam_idx, int n_threads, int max_iter, double* out_focal, double* out_radial,
                           double* out_R, double* out_t, double* stats /*[4]: initial mse, final mse, LM iters, seconds*/) {
  std::vector<pba::CameraT> cams(n_cam);
  for (int i = 0; i < n_cam; ++i) {
    cams[i].SetFocalLength(focal[i]);
    cams[i].SetProjectionDistortion(radial[i]);
    cams[i].SetMatrixRotation(R_rowmajor + 9 * i);
    cams[i].SetTranslation(t + 3 * i);
    if (cam_const[i]) cams[i].SetConstantCamera(); else cams[i].SetVariableCamera();
  }
  std::vector<pba::Point3D> pts(n_pts);
  for (int p = 0; p < n_pts; ++p) pts[p].SetPoint(xyz + 3 * p);
  std::vector<pba::Point2D> meas(n_proj);
  for (long k = 0; k < n_proj; ++k) meas[k].SetPoint2D(xy_centered[2 * k], xy_centered[2 * k + 1]);

  pba::ParallelBA pba(pba::ParallelBA::PBA_CPU_DOUBLE, n_threads);
  pba.SetNextBundleMode(pba::ParallelBA::BUNDLE_FULL);
  pba.EnableRadialDistortion(pba::ParallelBA::PBA_PROJECTION_DISTORTION);
  pba.SetFixedIntrinsics(false);
  pba::ConfigBA* cfg = pba.GetInternalConfig();
  cfg->__lm_delta_threshold /= 100.0f;
  cfg->__lm_gradient_threshold /= 100.0f;
  cfg->__lm_mse_threshold = 0.0f;
  cfg->__cg_min_iteration = 10;
  cfg->__verbose_level = 0;
  cfg->__lm_max_iteration = max_iter;
  pba.SetCameraData(cams.size(), cams.data());
  pba.SetPointData(pts.size(), pts.data());
  pba.SetProjection(meas.size(), meas.data(), pt_idx, cam_idx);
  const auto t0 = std::chrono::steady_clock::now();
  pba.RunBundleAdjustment();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  stats[0] = cfg->GetInitialMSE();
  stats[1] = cfg->GetFinalMSE();
  stats[2] = cfg->GetIterationsLM();
  stats[3] = secs;
  for (int i = 0; i < n_cam; ++i) {
    out_focal[i] = cams[i].GetFocalLength();
    out_radial[i] = cams[i].GetProjectionDistortion();
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out_R[9 * i + 3 * r + c] = cams[i].m[r][c];
    for (int k = 0; k < 3; ++k) out_t[3 * i + k] = cams[i].t[k];
  }
  for (int p = 0; p < n_pts; ++p) { float v[3]; pts[p].GetPoint(v); xyz[3 * p] = v[0]; xyz[3 * p + 1] = v[1]; xyz[3 * p + 2] = v[2]; }
  return 0;
}
