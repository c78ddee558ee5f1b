// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
// product path (dagsfm_b200/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// CPU restatement (FP64) of the reference's bundle adjustment as it is configured for
// the final global BA:
//   src/base/cost_functions.h:44-88     BundleAdjustmentCostFunction (variable pose)
//   src/base/cost_functions.h:90-158    BundleAdjustmentConstantPoseCostFunction
//   src/base/camera_models.h:714-757    SIMPLE_RADIAL WorldToImage (+ SIMPLE_PINHOLE, PINHOLE)
//   src/optim/bundle_adjustment.cc:338-526 AddImageToProblem / ParameterizeCameras / Points
//     (QuaternionParameterization on qvec, SubsetParameterization on tvec and intrinsics,
//      constant cameras / poses / points)
//   src/optim/bundle_adjustment.cc:258-310 Solve -> ceres::Solve
//   src/controllers/distributed_mapper_controller.cpp:522-542 GlobalBundleAdjustment options
// The arithmetic of the solve lives in Ceres Solver -- a third-party dependency that is NOT
// under /root/reference (find_package(Ceres), CMakeLists.txt:87; docker/Dockerfile:35 pins
// 1.14.0) and is not installed here.  Restated from its published algorithm and defaults
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, schur_complement_solver.cc,
// rotation.h, local_parameterization.cc of Ceres 1.14): Levenberg-Marquardt trust region,
// Jacobi column scaling 1/(1+|col|), diagonal clamp [1e-6, 1e32], radius 1e4 / update rule,
// min_relative_decrease 1e-3, exact Schur-complement step (DENSE/SPARSE_SCHUR).
// The Jacobian is analytic (Ceres uses autodiff of the same formulas); tests check it
// against central differences.
// ITERATIVE_SCHUR + SCHUR_JACOBI (selected above 1000 images, bundle_adjustment.cc:274-284) is restated further down
// from Ceres 1.14's implicit_schur_complement.cc / schur_jacobi_preconditioner.cc / conjugate_gradients_solver.cc /
// iterative_schur_complement_solver.cc; like the exact path it has no golden vectors in the reference.
// PARITY UNPINNED: the reference's tests pin only structure (num_residuals_reduced,
// num_effective_parameters_reduced, which blocks move: bundle_adjustment_test.cc:186-645)
// and four residual values (cost_functions_test.cc:41-98) -- replayed in
// tests/test_oracle_ba.py -- not the numeric result of a solve.  The vendored PBA built
// into oracle/_ref is the reference-owned convergence check / CPU timing baseline.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ba {

// ----------------------------------------------------------------- camera models
// src/base/camera_models.h:117-129; parameter layouts and the focal / principal point / extra index sets of every
// model's Initialize*Idxs.  kMaxParams = 12 (FULL_OPENCV, THIN_PRISM_FISHEYE).
static const int kMaxParams = 12;
static const int kNC = 6 + kMaxParams;  // camera-side columns per observation: rotation 3, translation 3, intrinsics <= 12
static int NumParams(int model) {
  static const int n[11] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
  return (model >= 0 && model <= 10) ? n[model] : 0;
}
static bool TwoFocal(int model) { return model == 1 || model == 4 || model == 5 || model == 6 || model == 7 || model == 10; }
// 0 focal, 1 principal point, 2 extra, -1 unused
static void ParamKinds(int model, int kind[kMaxParams]) {
  const int nf = TwoFocal(model) ? 2 : 1, n = NumParams(model);
  for (int k = 0; k < kMaxParams; ++k) kind[k] = k >= n ? -1 : k < nf ? 0 : k < nf + 2 ? 1 : 2;
}

// Forward-mode dual numbers: what Ceres' AutoDiffCostFunction does with Jets for the models without hand-derived
// formulas below (cost_functions.h:50-55).  d[0], d[1]: d/du, d/dv; d[2 + k]: d/dparams[k].
struct Dual {
  double a;
  double d[2 + kMaxParams];
  Dual() : a(0) { for (double& x : d) x = 0; }
  Dual(double c) : a(c) { for (double& x : d) x = 0; }
};
static Dual operator+(const Dual& x, const Dual& y) { Dual r; r.a = x.a + y.a; for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.d[i] + y.d[i]; return r; }
static Dual operator-(const Dual& x, const Dual& y) { Dual r; r.a = x.a - y.a; for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.d[i] - y.d[i]; return r; }
static Dual operator*(const Dual& x, const Dual& y) { Dual r; r.a = x.a * y.a; for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.a * y.d[i] + x.d[i] * y.a; return r; }
static Dual operator/(const Dual& x, const Dual& y) {
  Dual r;
  const double inv = 1.0 / y.a;
  r.a = x.a * inv;
  for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = (x.d[i] - r.a * y.d[i]) * inv;
  return r;
}
static Dual dsqrt(const Dual& x) { Dual r; r.a = std::sqrt(x.a); for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.d[i] / (2.0 * r.a); return r; }
static Dual datan(const Dual& x) { Dual r; r.a = std::atan(x.a); for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.d[i] / (1.0 + x.a * x.a); return r; }
static Dual dtan(const Dual& x) { Dual r; r.a = std::tan(x.a); for (int i = 0; i < 2 + kMaxParams; ++i) r.d[i] = x.d[i] * (1.0 + r.a * r.a); return r; }
static double dsqrt(double x) { return std::sqrt(x); }
static double datan(double x) { return std::atan(x); }
static double dtan(double x) { return std::tan(x); }
static double Val(double x) { return x; }
static double Val(const Dual& x) { return x.a; }

// CameraModel::WorldToImage (camera_models.h) for T = double or Dual, every model
template <typename T>
static void WorldToImageT(int model, const T* params, const T& u, const T& v, T* x, T* y) {
  const double eps = std::numeric_limits<double>::epsilon();
  const bool two = TwoFocal(model);
  const T* e = params + (two ? 4 : 3);
  T a = u, b = v;
  if (model == 7) {  // :1105-1175
    const T omega = e[0];
    const T radius2 = u * u + v * v, omega2 = omega * omega;
    T factor;
    if (Val(omega2) < 1e-4) {
      factor = (omega2 * radius2) / T(3) - omega2 / T(12) + T(1);
    } else if (Val(radius2) < 1e-4) {
      const T tan_half_omega = dtan(omega / T(2));
      factor = (T(-2) * tan_half_omega * (T(4) * radius2 * tan_half_omega * tan_half_omega - T(3))) / (T(3) * omega);
    } else {
      const T radius = dsqrt(radius2);
      const T numerator = datan(radius * T(2) * dtan(omega / T(2)));
      factor = numerator / (radius * omega);
    }
    a = u * factor;
    b = v * factor;
  } else if (model >= 2) {
    if (model == 10) {  // :1406-1422
      const T r = dsqrt(u * u + v * v);
      if (Val(r) > eps) {
        const T theta = datan(r);
        a = theta * u / r;
        b = theta * v / r;
      }
    }
    const T u2 = a * a, uv = a * b, v2 = b * b, r2 = u2 + v2;
    T du = T(0), dv = T(0);
    if (model == 2) {  // :747-757
      const T radial = e[0] * r2;
      du = a * radial; dv = b * radial;
    } else if (model == 3) {  // :816-828
      const T radial = e[0] * r2 + e[1] * r2 * r2;
      du = a * radial; dv = b * radial;
    } else if (model == 4) {  // :888-903
      const T radial = e[0] * r2 + e[1] * r2 * r2;
      du = a * radial + T(2) * e[2] * uv + e[3] * (r2 + T(2) * u2);
      dv = b * radial + T(2) * e[3] * uv + e[2] * (r2 + T(2) * v2);
    } else if (model == 5 || model == 8 || model == 9) {  // :963-986, :1272-1290, :1348-1368
      const T r = dsqrt(a * a + b * b);
      if (Val(r) > eps) {
        const T theta = datan(r), theta2 = theta * theta, theta4 = theta2 * theta2;
        T thetad;
        if (model == 5) thetad = theta * (T(1) + e[0] * theta2 + e[1] * theta4 + e[2] * (theta4 * theta2) + e[3] * (theta4 * theta4));
        else if (model == 8) thetad = theta * (T(1) + e[0] * theta2);
        else thetad = theta * (T(1) + e[0] * theta2 + e[1] * theta4);
        du = a * thetad / r - a;
        dv = b * thetad / r - b;
      }
    } else if (model == 6) {  // :1058-1080
      const T r4 = r2 * r2, r6 = r4 * r2;
      const T radial = (T(1) + e[0] * r2 + e[1] * r4 + e[4] * r6) / (T(1) + e[5] * r2 + e[6] * r4 + e[7] * r6);
      du = a * radial + T(2) * e[2] * uv + e[3] * (r2 + T(2) * u2) - a;
      dv = b * radial + T(2) * e[3] * uv + e[2] * (r2 + T(2) * v2) - b;
    } else if (model == 10) {  // :1460-1482
      const T r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
      const T radial = e[0] * r2 + e[1] * r4 + e[4] * r6 + e[5] * r8;
      du = a * radial + T(2) * e[2] * uv + e[3] * (r2 + T(2) * u2) + e[6] * r2;
      dv = b * radial + T(2) * e[3] * uv + e[2] * (r2 + T(2) * v2) + e[7] * r2;
    }
    a = a + du;
    b = b + dv;
  }
  if (two) { *x = params[0] * a + params[2]; *y = params[1] * b + params[3]; }
  else { *x = params[0] * a + params[1]; *y = params[0] * b + params[2]; }
}

// residual + Jacobians of one observation.
// q (w,x,y,z), t, X, params -> r[2]; Jq 2x3 (local, after the 4x3 parameterization Jacobian),
// Jt 2x3, JX 2x3, Jk 2x12 (row stride kMaxParams).
static void Evaluate(int model, const double* q, const double* t, const double* X, const double* k, const double* obs,
                     double* r, double* Jq, double* Jt, double* JX, double* Jk) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  // ceres::UnitQuaternionRotatePoint
  const double t2 = w * x, t3 = w * y, t4 = w * z, t5 = -x * x, t6 = x * y, t7 = x * z, t8 = -y * y, t9 = y * z,
               t1 = -z * z;
  double p[3];
  p[0] = 2 * ((t8 + t1) * X[0] + (t6 - t4) * X[1] + (t3 + t7) * X[2]) + X[0];
  p[1] = 2 * ((t4 + t6) * X[0] + (t5 + t1) * X[1] + (t9 - t2) * X[2]) + X[1];
  p[2] = 2 * ((t7 - t3) * X[0] + (t2 + t9) * X[1] + (t5 + t8) * X[2]) + X[2];
  p[0] += t[0];
  p[1] += t[1];
  p[2] += t[2];
  const double u = p[0] / p[2], v = p[1] / p[2];
  double xi, yi;          // image point
  double dxdu, dxdv, dydu, dydv;
  double dk[2][kMaxParams];
  for (int a = 0; a < 2; ++a) for (int b = 0; b < kMaxParams; ++b) dk[a][b] = 0;
  if (model == 0) {
    xi = k[0] * u + k[1]; yi = k[0] * v + k[2];
    dxdu = k[0]; dxdv = 0; dydu = 0; dydv = k[0];
    dk[0][0] = u; dk[0][1] = 1; dk[1][0] = v; dk[1][2] = 1;
  } else if (model == 1) {
    xi = k[0] * u + k[2]; yi = k[1] * v + k[3];
    dxdu = k[0]; dxdv = 0; dydu = 0; dydv = k[1];
    dk[0][0] = u; dk[0][2] = 1; dk[1][1] = v; dk[1][3] = 1;
  } else if (model == 2) {
    const double u2 = u * u, v2 = v * v, r2 = u2 + v2, radial = k[3] * r2;
    const double du = u * radial, dv = v * radial;
    const double xd = u + du, yd = v + dv;
    xi = k[0] * xd + k[1]; yi = k[0] * yd + k[2];
    dxdu = k[0] * (1 + radial + 2 * k[3] * u2); dxdv = k[0] * (2 * k[3] * u * v);
    dydu = k[0] * (2 * k[3] * u * v); dydv = k[0] * (1 + radial + 2 * k[3] * v2);
    dk[0][0] = xd; dk[0][1] = 1; dk[0][3] = k[0] * u * r2;
    dk[1][0] = yd; dk[1][2] = 1; dk[1][3] = k[0] * v * r2;
  } else {  // the other eight models: WorldToImage on dual numbers (Ceres differentiates the same template with Jets)
    Dual pj[kMaxParams], uj(u), vj(v), xj, yj;
    const int np = NumParams(model);
    for (int a = 0; a < np; ++a) { pj[a] = Dual(k[a]); pj[a].d[2 + a] = 1.0; }
    uj.d[0] = 1.0;
    vj.d[1] = 1.0;
    WorldToImageT<Dual>(model, pj, uj, vj, &xj, &yj);
    xi = xj.a; yi = yj.a;
    dxdu = xj.d[0]; dxdv = xj.d[1]; dydu = yj.d[0]; dydv = yj.d[1];
    for (int a = 0; a < np; ++a) { dk[0][a] = xj.d[2 + a]; dk[1][a] = yj.d[2 + a]; }
  }
  r[0] = xi - obs[0];
  r[1] = yi - obs[1];
  if (!Jq) return;
  // d(u,v)/dp
  const double ip2 = 1.0 / p[2];
  const double dudp[3] = {ip2, 0, -p[0] * ip2 * ip2}, dvdp[3] = {0, ip2, -p[1] * ip2 * ip2};
  double drdp[2][3];
  for (int c = 0; c < 3; ++c) {
    drdp[0][c] = dxdu * dudp[c] + dxdv * dvdp[c];
    drdp[1][c] = dydu * dudp[c] + dydv * dvdp[c];
  }
  // dp/dX = 2M + I (the rotation matrix of the formula)
  const double R[3][3] = {{2 * (t8 + t1) + 1, 2 * (t6 - t4), 2 * (t3 + t7)},
                          {2 * (t4 + t6), 2 * (t5 + t1) + 1, 2 * (t9 - t2)},
                          {2 * (t7 - t3), 2 * (t2 + t9), 2 * (t5 + t8) + 1}};
  for (int i = 0; i < 2; ++i)
    for (int c = 0; c < 3; ++c) {
      JX[3 * i + c] = drdp[i][0] * R[0][c] + drdp[i][1] * R[1][c] + drdp[i][2] * R[2][c];
      Jt[3 * i + c] = drdp[i][c];
    }
  // dp/dq (3x4): 2 * dM/dq_k * X
  const double X0 = X[0], X1 = X[1], X2 = X[2];
  const double dpdq[3][4] = {
      {2 * (-z * X1 + y * X2), 2 * (y * X1 + z * X2), 2 * (-2 * y * X0 + x * X1 + w * X2), 2 * (-2 * z * X0 - w * X1 + x * X2)},
      {2 * (z * X0 - x * X2), 2 * (y * X0 - 2 * x * X1 - w * X2), 2 * (x * X0 + z * X2), 2 * (w * X0 - 2 * z * X1 + y * X2)},
      {2 * (-y * X0 + x * X1), 2 * (z * X0 + w * X1 - 2 * x * X2), 2 * (-w * X0 + z * X1 - 2 * y * X2), 2 * (x * X0 + y * X1)}};
  // QuaternionParameterization::ComputeJacobian (4x3)
  const double JL[4][3] = {{-x, -y, -z}, {w, z, -y}, {-z, w, x}, {y, -x, w}};
  for (int i = 0; i < 2; ++i) {
    double drdq[4];
    for (int a = 0; a < 4; ++a) drdq[a] = drdp[i][0] * dpdq[0][a] + drdp[i][1] * dpdq[1][a] + drdp[i][2] * dpdq[2][a];
    for (int c = 0; c < 3; ++c)
      Jq[3 * i + c] = drdq[0] * JL[0][c] + drdq[1] * JL[1][c] + drdq[2] * JL[2][c] + drdq[3] * JL[3][c];
    for (int a = 0; a < kMaxParams; ++a) Jk[kMaxParams * i + a] = dk[i][a];
  }
}

// QuaternionParameterization::Plus
static void QuatPlus(const double* x, const double* delta, double* out) {
  const double nd = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    const double s = std::sin(nd) / nd;
    const double qd[4] = {std::cos(nd), s * delta[0], s * delta[1], s * delta[2]};
    // QuaternionProduct(q_delta, x)
    out[0] = qd[0] * x[0] - qd[1] * x[1] - qd[2] * x[2] - qd[3] * x[3];
    out[1] = qd[0] * x[1] + qd[1] * x[0] + qd[2] * x[3] - qd[3] * x[2];
    out[2] = qd[0] * x[2] - qd[1] * x[3] + qd[2] * x[0] + qd[3] * x[1];
    out[3] = qd[0] * x[3] + qd[1] * x[2] - qd[2] * x[1] + qd[3] * x[0];
  } else {
    for (int i = 0; i < 4; ++i) out[i] = x[i];
  }
}

struct Problem {
  int n_img, n_cam, n_pts;
  long n_obs;
  double *qvec, *tvec;        // [n_img*4], [n_img*3]
  const int* img_cam;         // camera of each image
  const uint8_t* pose_const;  // 1 = constant pose
  const uint8_t* tvec_const;  // bitmask of constant tvec components
  const int* cam_model;
  double* cam_params;         // [n_cam*cam_stride]
  int cam_stride = 4;         // doubles per camera in cam_params (>= NumParams of every model used)
  const uint8_t* cam_const;   // 1 = whole camera constant
  int refine_focal, refine_principal, refine_extra;
  double* xyz;
  const uint8_t* pt_const;
  const int *obs_img, *obs_pt;  // sorted by point
  const double* obs_xy;
};
struct Options {
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  int n_threads;
  int loss_type = 0;        // BundleAdjustmentOptions::LossFunctionType: 0 TRIVIAL, 1 SOFT_L1, 2 CAUCHY
  double loss_scale = 1.0;  // loss_function_scale
  // 0 = exact Schur step (DENSE_SCHUR / SPARSE_SCHUR), 1 = ITERATIVE_SCHUR + SCHUR_JACOBI, which
  // BundleAdjuster::Solve selects above 1000 images (bundle_adjustment.cc:274-284)
  int linear_solver = 0;
  int max_linear_solver_iterations = 100;  // GlobalBundleAdjustment(): distributed_mapper_controller.cpp:529
};

// ceres::LossFunction::Evaluate (Ceres 1.14 loss_function.cc) for the three types
// BundleAdjustmentOptions::CreateLossFunction can build (bundle_adjustment.cc:53-68):
// rho[0] = rho(s), rho[1] = rho'(s), rho[2] = rho''(s) with s = |r|^2 of one observation.
static void LossEvaluate(int type, double a, double s, double rho[3]) {
  if (type == 1) {  // SoftLOneLoss(a): b = a^2, c = 1 / b
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double tmp = std::sqrt(sum);
    rho[0] = 2.0 * b * (tmp - 1.0);
    rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
    rho[2] = -(c * rho[1]) / (2.0 * sum);
  } else if (type == 2) {  // CauchyLoss(a)
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}
// ceres::internal::Corrector (corrector.cc): for rho'' <= 0 -- always the case for SoftLOne and
// Cauchy -- residual and Jacobian of the block are both scaled by sqrt(rho'); otherwise the
// Triggs correction with alpha = 1 - sqrt(1 + 2 s rho''/rho').  Returns the residual scaling and
// writes sqrt(rho') and alpha / s.
static double Corrector(double s, const double rho[3], double* sqrt_rho1, double* alpha_sq_norm) {
  *sqrt_rho1 = std::sqrt(rho[1]);
  if (s == 0.0 || rho[2] <= 0.0) {
    *alpha_sq_norm = 0.0;
    return *sqrt_rho1;
  }
  const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
  const double alpha = 1.0 - std::sqrt(D);
  *alpha_sq_norm = alpha / s;
  return *sqrt_rho1 / (1 - alpha);
}
struct Summary {
  double initial_cost, final_cost;
  int num_successful_steps, num_unsuccessful_steps, termination;  // 0 convergence, 1 no convergence, 2 failure
  int num_residuals, num_effective_parameters;
  double seconds;
  long num_linear_iterations = 0;  // CG iterations over all LM iterations (ITERATIVE_SCHUR)
};

struct Layout {
  std::vector<int> pose_col;  // [n_img*6], -1 = constant
  std::vector<int> intr_col;  // [n_cam*kMaxParams]
  std::vector<int> pt_col;    // [n_pts] first of 3, -1 = constant (relative to the point block)
  int n_cam_cols = 0, n_pt_var = 0;
};

static Layout MakeLayout(const Problem& P) {
  Layout L;
  L.pose_col.assign((size_t)P.n_img * 6, -1);
  L.intr_col.assign((size_t)P.n_cam * kMaxParams, -1);
  L.pt_col.assign(P.n_pts, -1);
  int c = 0;
  // which cameras / images / points actually appear
  std::vector<char> img_used(P.n_img, 0), cam_used(P.n_cam, 0), pt_used(P.n_pts, 0);
  for (long o = 0; o < P.n_obs; ++o) { img_used[P.obs_img[o]] = 1; cam_used[P.img_cam[P.obs_img[o]]] = 1; pt_used[P.obs_pt[o]] = 1; }
  for (int i = 0; i < P.n_img; ++i) {
    if (!img_used[i] || P.pose_const[i]) continue;
    for (int k = 0; k < 3; ++k) L.pose_col[6 * i + k] = c++;
    for (int k = 0; k < 3; ++k)
      if (!(P.tvec_const[i] & (1 << k))) L.pose_col[6 * i + 3 + k] = c++;
  }
  for (int cm = 0; cm < P.n_cam; ++cm) {
    if (!cam_used[cm] || P.cam_const[cm]) continue;
    int kind[kMaxParams];
    ParamKinds(P.cam_model[cm], kind);
    for (int k = 0; k < NumParams(P.cam_model[cm]); ++k) {
      const bool var = (kind[k] == 0 && P.refine_focal) || (kind[k] == 1 && P.refine_principal) || (kind[k] == 2 && P.refine_extra);
      if (var) L.intr_col[kMaxParams * cm + k] = c++;
    }
  }
  L.n_cam_cols = c;
  int pv = 0;
  for (int p = 0; p < P.n_pts; ++p)
    if (pt_used[p] && !P.pt_const[p]) L.pt_col[p] = pv++;
  L.n_pt_var = pv;
  return L;
}

// In-place Cholesky (lower) of the dense SPD matrix A (n x n row-major); returns false if not PD.
static bool CholeskySolve(std::vector<double>& A, int n, std::vector<double>& b) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    for (int j = k0; j < k0 + kb; ++j) {  // factor the diagonal block and the panel below it column by column
      double d = A[(size_t)j * n + j];
      for (int k = k0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
      if (!(d > 0)) return false;
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
#pragma omp parallel for schedule(static)
      for (int i = j + 1; i < n; ++i) {
        double s = A[(size_t)i * n + j];
        for (int k = k0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
        A[(size_t)i * n + j] = s / d;
      }
    }
    const int r0 = k0 + kb;  // trailing update A22 -= L21 L21^T (lower part)
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = r0; i < n; ++i) {
      const double* li = &A[(size_t)i * n + k0];
      for (int j = r0; j <= i; ++j) {
        const double* lj = &A[(size_t)j * n + k0];
        double s = 0;
        for (int k = 0; k < kb; ++k) s += li[k] * lj[k];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

struct ObsJ { double r[2]; double Jc[2][kNC]; double Jp[2][3]; int col[kNC]; double rho0; };

struct Solver {
  Problem P;
  Options O;
  Layout L;
  std::vector<long> pt_start;      // CSR by point
  std::vector<double> scale_c, scale_p;  // Jacobi column scaling
  std::vector<ObsJ> J;

  double EvalCost(const double* q, const double* t, const double* kp, const double* X) const {
    double cost = 0;
#pragma omp parallel for reduction(+ : cost) schedule(static)
    for (long o = 0; o < P.n_obs; ++o) {
      const int i = P.obs_img[o], p = P.obs_pt[o], c = P.img_cam[i];
      double r[2];
      Evaluate(P.cam_model[c], q + 4 * i, t + 3 * i, X + 3 * p, kp + P.cam_stride * c, P.obs_xy + 2 * o, r, nullptr, nullptr, nullptr, nullptr);
      if (O.loss_type == 0) {
        cost += r[0] * r[0] + r[1] * r[1];
      } else {
        double rho[3];
        LossEvaluate(O.loss_type, O.loss_scale, r[0] * r[0] + r[1] * r[1], rho);
        cost += rho[0];
      }
    }
    return 0.5 * cost;
  }
  void EvalJacobian(bool scaled) {
#pragma omp parallel for schedule(static)
    for (long o = 0; o < P.n_obs; ++o) {
      const int i = P.obs_img[o], p = P.obs_pt[o], c = P.img_cam[i];
      double Jq[6], Jt[6], JX[6], Jk[2 * kMaxParams];
      ObsJ& e = J[o];
      Evaluate(P.cam_model[c], P.qvec + 4 * i, P.tvec + 3 * i, P.xyz + 3 * p, P.cam_params + P.cam_stride * c, P.obs_xy + 2 * o, e.r, Jq, Jt, JX, Jk);
      for (int k = 0; k < 6; ++k) e.col[k] = L.pose_col[6 * i + k];
      for (int k = 0; k < kMaxParams; ++k) e.col[6 + k] = L.intr_col[kMaxParams * c + k];
      e.rho0 = e.r[0] * e.r[0] + e.r[1] * e.r[1];
      double jscale = 1.0;
      if (O.loss_type != 0) {  // ResidualBlock::Evaluate: correct the Jacobians, then the residuals
        double rho[3], sqrt_rho1, alpha_sq_norm;
        const double sq = e.rho0;
        LossEvaluate(O.loss_type, O.loss_scale, sq, rho);
        const double rscale = Corrector(sq, rho, &sqrt_rho1, &alpha_sq_norm);
        // rho'' <= 0 for both robust types, so alpha_sq_norm == 0: a pure scaling by sqrt(rho')
        jscale = sqrt_rho1;
        e.r[0] *= rscale;
        e.r[1] *= rscale;
        e.rho0 = rho[0];
      }
      for (int a = 0; a < 2; ++a) {
        for (int k = 0; k < 3; ++k) { e.Jc[a][k] = jscale * Jq[3 * a + k]; e.Jc[a][3 + k] = jscale * Jt[3 * a + k]; e.Jp[a][k] = jscale * JX[3 * a + k]; }
        for (int k = 0; k < kMaxParams; ++k) e.Jc[a][6 + k] = jscale * Jk[kMaxParams * a + k];
        for (int k = 0; k < kNC; ++k) {
          if (e.col[k] < 0) e.Jc[a][k] = 0;
          else if (scaled) e.Jc[a][k] *= scale_c[e.col[k]];
        }
        if (L.pt_col[p] < 0) { e.Jp[a][0] = e.Jp[a][1] = e.Jp[a][2] = 0; }
        else if (scaled) for (int k = 0; k < 3; ++k) e.Jp[a][k] *= scale_p[3 * L.pt_col[p] + k];
      }
    }
  }
};


// ------------------------------------------------------------------ ITERATIVE_SCHUR
// Ceres 1.14 (external; restated from its published sources): iterative_schur_complement_solver.cc
// runs ConjugateGradientsSolver on the ImplicitSchurComplement operator
//   S x = (F'F + D_f^2) x - F'E (E'E + D_e^2)^-1 E'F x        (implicit_schur_complement.cc)
// (E = point columns, F = camera-side columns of the Jacobian, never forming S), preconditioned by
// SCHUR_JACOBI = the inverse of the block diagonal of S, one block per camera-side PARAMETER BLOCK
// (schur_jacobi_preconditioner.cc).  colmap adds qvec, tvec and the camera parameters as three
// separate blocks (bundle_adjustment.cc:383-418), so the blocks are 3x3 (rotation, local), <=3x3
// (variable tvec components) and <=4x4 (variable intrinsics, possibly shared by many images).
// The start vector is zero, r_tolerance = -1 (off) and q_tolerance = eta = 0.1
// (levenberg_marquardt_strategy.cc, Solver::Options::eta default), residual_reset_period = 10.
struct BlockDiag {
  std::vector<int> first, size;  // per column: first column and size of its parameter block
  std::vector<double> M;         // [D][kMaxParams]: row `col`, entries (col, first + j)
};
static BlockDiag MakeBlocks(const Problem& P, const Layout& L) {
  BlockDiag B;
  const int D = L.n_cam_cols;
  B.first.assign(std::max(D, 1), 0);
  B.size.assign(std::max(D, 1), 0);
  auto mark = [&](const int* cols, int n) {
    int f = -1, cnt = 0;
    for (int k = 0; k < n; ++k) if (cols[k] >= 0) { if (f < 0) f = cols[k]; ++cnt; }
    for (int k = 0; k < n; ++k) if (cols[k] >= 0) { B.first[cols[k]] = f; B.size[cols[k]] = cnt; }
  };
  for (int i = 0; i < P.n_img; ++i) { mark(&L.pose_col[6 * i], 3); mark(&L.pose_col[6 * i + 3], 3); }
  for (int c = 0; c < P.n_cam; ++c) mark(&L.intr_col[kMaxParams * c], kMaxParams);
  B.M.assign((size_t)std::max(D, 1) * kMaxParams, 0.0);
  return B;
}
// BlockRandomAccessDiagonalMatrix::Invert: block.llt().solve(Identity); returns false if a block is not PD
static bool InvertBlocks(BlockDiag& B, int D) {
  for (int f = 0; f < D; f += B.size[f]) {
    const int n = B.size[f];
    double A[kMaxParams][kMaxParams], Lc[kMaxParams][kMaxParams] = {{0}}, Inv[kMaxParams][kMaxParams];
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) A[r][c] = B.M[(size_t)(f + r) * kMaxParams + c];
    for (int j = 0; j < n; ++j) {
      double d = A[j][j];
      for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
      if (!(d > 0)) return false;
      Lc[j][j] = std::sqrt(d);
      for (int i = j + 1; i < n; ++i) {
        double v = A[i][j];
        for (int k = 0; k < j; ++k) v -= Lc[i][k] * Lc[j][k];
        Lc[i][j] = v / Lc[j][j];
      }
    }
    for (int c = 0; c < n; ++c) {  // solve L L^T x = e_c
      double y[kMaxParams];
      for (int i = 0; i < n; ++i) {
        double v = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < i; ++k) v -= Lc[i][k] * y[k];
        y[i] = v / Lc[i][i];
      }
      for (int i = n - 1; i >= 0; --i) {
        double v = y[i];
        for (int k = i + 1; k < n; ++k) v -= Lc[k][i] * Inv[k][c];
        Inv[i][c] = v / Lc[i][i];
      }
    }
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) B.M[(size_t)(f + r) * kMaxParams + c] = Inv[r][c];
  }
  return true;
}
static void ApplyBlocks(const BlockDiag& B, int D, const double* r, double* z) {
  for (int j = 0; j < D; ++j) {
    double v = 0;
    for (int k = 0; k < B.size[j]; ++k) v += B.M[(size_t)j * kMaxParams + k] * r[B.first[j] + k];
    z[j] = v;
  }
}

static void Solve(Problem P, Options O, Summary* S) {
  const auto t_start = std::chrono::steady_clock::now();
  Solver sv;
  sv.P = P;
  sv.O = O;
  sv.L = MakeLayout(P);
  const Layout& L = sv.L;
  const int D = L.n_cam_cols;
  const int NP = L.n_pt_var;
  // image.NormalizeQvec() (bundle_adjustment.cc:345)
  for (int i = 0; i < P.n_img; ++i) {
    double* q = P.qvec + 4 * i;
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n > 0) for (int k = 0; k < 4; ++k) q[k] /= n;
  }
  sv.pt_start.assign(P.n_pts + 1, 0);
  for (long o = 0; o < P.n_obs; ++o) sv.pt_start[P.obs_pt[o] + 1]++;
  for (int p = 0; p < P.n_pts; ++p) sv.pt_start[p + 1] += sv.pt_start[p];
  sv.J.resize(P.n_obs);
  sv.scale_c.assign(std::max(D, 1), 1.0);
  sv.scale_p.assign(std::max(3 * NP, 1), 1.0);
  {  // Ceres' reduced program drops residual blocks whose parameter blocks are all constant
    long n = 0;
    for (long o = 0; o < P.n_obs; ++o) {
      const int i = P.obs_img[o], c = P.img_cam[i];
      bool free_block = L.pt_col[P.obs_pt[o]] >= 0;
      for (int k = 0; k < 6 && !free_block; ++k) free_block = L.pose_col[6 * i + k] >= 0;
      for (int k = 0; k < kMaxParams && !free_block; ++k) free_block = L.intr_col[kMaxParams * c + k] >= 0;
      n += free_block ? 1 : 0;
    }
    S->num_residuals = (int)(2 * n);
  }
  S->num_effective_parameters = D + 3 * NP;
  S->num_successful_steps = S->num_unsuccessful_steps = 0;
  S->termination = 1;
  if (P.n_obs == 0) { S->initial_cost = S->final_cost = 0; S->seconds = 0; return; }

  // initial evaluation + Jacobi scaling (scale = 1 / (1 + column norm))
  sv.EvalJacobian(false);
  {
    std::vector<double> nc(std::max(D, 1), 0.0), np(std::max(3 * NP, 1), 0.0);
    for (long o = 0; o < P.n_obs; ++o) {
      const ObsJ& e = sv.J[o];
      const int pc = L.pt_col[P.obs_pt[o]];
      for (int a = 0; a < 2; ++a) {
        for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) nc[e.col[k]] += e.Jc[a][k] * e.Jc[a][k];
        if (pc >= 0) for (int k = 0; k < 3; ++k) np[3 * pc + k] += e.Jp[a][k] * e.Jp[a][k];
      }
    }
    for (int j = 0; j < D; ++j) sv.scale_c[j] = 1.0 / (1.0 + std::sqrt(nc[j]));
    for (int j = 0; j < 3 * NP; ++j) sv.scale_p[j] = 1.0 / (1.0 + std::sqrt(np[j]));
  }
  double cost = 0;
  for (long o = 0; o < P.n_obs; ++o) cost += sv.J[o].rho0;  // |r|^2, or rho(|r|^2) under a robust loss
  cost *= 0.5;
  S->initial_cost = cost;
  sv.EvalJacobian(true);

  double radius = 1e4, decrease_factor = 2.0;
  const double min_radius = 1e-32, max_radius = 1e16, min_rel_dec = 1e-3, min_diag = 1e-6, max_diag = 1e32;
  std::vector<double> Smat, gc(std::max(D, 1)), dc(std::max(D, 1)), diag_c(std::max(D, 1)), diag_p(std::max(3 * NP, 1));
  std::vector<double> Vinv((size_t)std::max(NP, 1) * 9), gp((size_t)std::max(3 * NP, 1)), dp((size_t)std::max(3 * NP, 1));
  std::vector<double> qn(P.n_img * 4), tn(P.n_img * 3), kn(P.n_cam * P.cam_stride), Xn(P.n_pts * 3);
  bool need_grad_check = true;

  for (int iter = 0; iter < O.max_num_iterations; ++iter) {
    // diag(J^T J) and gradient (scaled Jacobian)
    std::fill(diag_c.begin(), diag_c.end(), 0.0);
    std::fill(diag_p.begin(), diag_p.end(), 0.0);
    std::fill(gc.begin(), gc.end(), 0.0);
    std::fill(gp.begin(), gp.end(), 0.0);
    for (long o = 0; o < P.n_obs; ++o) {
      const ObsJ& e = sv.J[o];
      const int pc = L.pt_col[P.obs_pt[o]];
      for (int a = 0; a < 2; ++a) {
        for (int k = 0; k < kNC; ++k)
          if (e.col[k] >= 0) { diag_c[e.col[k]] += e.Jc[a][k] * e.Jc[a][k]; gc[e.col[k]] += e.Jc[a][k] * e.r[a]; }
        if (pc >= 0) for (int k = 0; k < 3; ++k) { diag_p[3 * pc + k] += e.Jp[a][k] * e.Jp[a][k]; gp[3 * pc + k] += e.Jp[a][k] * e.r[a]; }
      }
    }
    if (need_grad_check) {  // gradient_max_norm of the unscaled problem
      double gmax = 0;
      // |x - Plus(x, -g)|_inf (trust_region_minimizer.cc): plain |g_j| under identity / subset parameterisations, the
      // displacement of the quaternion under QuaternionParameterization::Plus for a rotation block
      std::vector<char> is_rot((size_t)std::max(D, 1), 0);
      for (int i = 0; i < P.n_img; ++i) {
        const int c = L.pose_col[6 * i];
        if (c < 0) continue;
        is_rot[c] = is_rot[c + 1] = is_rot[c + 2] = 1;
        const double d[3] = {-gc[c] / sv.scale_c[c], -gc[c + 1] / sv.scale_c[c + 1], -gc[c + 2] / sv.scale_c[c + 2]};
        double y[4];
        QuatPlus(P.qvec + 4 * i, d, y);
        for (int a = 0; a < 4; ++a) gmax = std::max(gmax, std::abs(P.qvec[4 * i + a] - y[a]));
      }
      for (int j = 0; j < D; ++j)
        if (!is_rot[j]) gmax = std::max(gmax, std::abs(gc[j] / sv.scale_c[j]));
      for (int j = 0; j < 3 * NP; ++j) gmax = std::max(gmax, std::abs(gp[j] / sv.scale_p[j]));
      if (gmax <= O.gradient_tolerance) { S->termination = 0; break; }
      need_grad_check = false;
    }
    // LM diagonal D^2 = clamp(diag) / radius
    std::vector<double> lm_c(D), lm_p(3 * NP);
    for (int j = 0; j < D; ++j) lm_c[j] = std::min(std::max(diag_c[j], min_diag), max_diag) / radius;
    for (int j = 0; j < 3 * NP; ++j) lm_p[j] = std::min(std::max(diag_p[j], min_diag), max_diag) / radius;
    bool ok = true;
    if (O.linear_solver == 1) {
      // ---- ITERATIVE_SCHUR: (E'E + D_e^2)^-1 per point, then CG on the implicit operator
#pragma omp parallel for schedule(static)
      for (int p = 0; p < P.n_pts; ++p) {
        const int pc = L.pt_col[p];
        if (pc < 0) continue;
        double V[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (long o = sv.pt_start[p]; o < sv.pt_start[p + 1]; ++o) {
          const ObsJ& e = sv.J[o];
          for (int a = 0; a < 2; ++a)
            for (int k = 0; k < 3; ++k)
              for (int l = 0; l < 3; ++l) V[k][l] += e.Jp[a][k] * e.Jp[a][l];
        }
        for (int k = 0; k < 3; ++k) V[k][k] += lm_p[3 * pc + k];
        const double c00 = V[1][1] * V[2][2] - V[1][2] * V[2][1], c01 = V[1][2] * V[2][0] - V[1][0] * V[2][2],
                     c02 = V[1][0] * V[2][1] - V[1][1] * V[2][0];
        const double det = V[0][0] * c00 + V[0][1] * c01 + V[0][2] * c02, id = 1.0 / det;
        double* Vi = &Vinv[(size_t)pc * 9];
        Vi[0] = c00 * id; Vi[1] = (V[0][2] * V[2][1] - V[0][1] * V[2][2]) * id; Vi[2] = (V[0][1] * V[1][2] - V[0][2] * V[1][1]) * id;
        Vi[3] = c01 * id; Vi[4] = (V[0][0] * V[2][2] - V[0][2] * V[2][0]) * id; Vi[5] = (V[0][2] * V[1][0] - V[0][0] * V[1][2]) * id;
        Vi[6] = c02 * id; Vi[7] = (V[0][1] * V[2][0] - V[0][0] * V[2][1]) * id; Vi[8] = (V[0][0] * V[1][1] - V[0][1] * V[1][0]) * id;
      }
      // S x in residual space (ImplicitSchurComplement::RightMultiply): y1 = F x; y3 = -(E'E)^-1 E' y1;
      // y1 += E y3; out = D_f^2 x + F' y1
      auto SchurMul = [&](const std::vector<double>& x, std::vector<double>& out) {
        for (int j = 0; j < D; ++j) out[j] = lm_c[j] * x[j];
        for (int p = 0; p < P.n_pts; ++p) {
          const int pc = L.pt_col[p];
          const long o0 = sv.pt_start[p], o1 = sv.pt_start[p + 1];
          double y2[3] = {0, 0, 0}, z[3] = {0, 0, 0};
          if (pc >= 0) {
            for (long o = o0; o < o1; ++o) {
              const ObsJ& e = sv.J[o];
              for (int a = 0; a < 2; ++a) {
                double u = 0;
                for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) u += e.Jc[a][k] * x[e.col[k]];
                for (int k = 0; k < 3; ++k) y2[k] += e.Jp[a][k] * u;
              }
            }
            const double* Vi = &Vinv[(size_t)pc * 9];
            for (int k = 0; k < 3; ++k) z[k] = Vi[3 * k] * y2[0] + Vi[3 * k + 1] * y2[1] + Vi[3 * k + 2] * y2[2];
          }
          for (long o = o0; o < o1; ++o) {
            const ObsJ& e = sv.J[o];
            for (int a = 0; a < 2; ++a) {
              double u = 0;
              for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) u += e.Jc[a][k] * x[e.col[k]];
              u -= e.Jp[a][0] * z[0] + e.Jp[a][1] * z[1] + e.Jp[a][2] * z[2];
              for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) out[e.col[k]] += e.Jc[a][k] * u;
            }
          }
        }
      };
      // rhs = F'(b - E (E'E)^-1 E'b)   (ImplicitSchurComplement::UpdateRhs; b = the residual vector here)
      std::vector<double> rhs(std::max(D, 1), 0.0);
      for (int p = 0; p < P.n_pts; ++p) {
        const int pc = L.pt_col[p];
        double tp[3] = {0, 0, 0};
        if (pc >= 0) {
          const double* Vi = &Vinv[(size_t)pc * 9];
          for (int k = 0; k < 3; ++k) tp[k] = Vi[3 * k] * gp[3 * pc] + Vi[3 * k + 1] * gp[3 * pc + 1] + Vi[3 * k + 2] * gp[3 * pc + 2];
        }
        for (long o = sv.pt_start[p]; o < sv.pt_start[p + 1]; ++o) {
          const ObsJ& e = sv.J[o];
          for (int a = 0; a < 2; ++a) {
            const double u = e.r[a] - (e.Jp[a][0] * tp[0] + e.Jp[a][1] * tp[1] + e.Jp[a][2] * tp[2]);
            for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) rhs[e.col[k]] += e.Jc[a][k] * u;
          }
        }
      }
      // SCHUR_JACOBI: block diagonal of S, inverted
      BlockDiag B = MakeBlocks(P, L);
      for (int j = 0; j < D; ++j) B.M[(size_t)j * kMaxParams + (j - B.first[j])] += lm_c[j];
      for (int p = 0; p < P.n_pts; ++p) {
        const int pc = L.pt_col[p];
        const long o0 = sv.pt_start[p], o1 = sv.pt_start[p + 1];
        const double* Vi = pc >= 0 ? &Vinv[(size_t)pc * 9] : nullptr;
        for (long oa = o0; oa < o1; ++oa) {
          const ObsJ& ea = sv.J[oa];
          for (int k = 0; k < kNC; ++k) {  // U = F'F restricted to the blocks
            if (ea.col[k] < 0) continue;
            for (int l = 0; l < kNC; ++l)
              if (ea.col[l] >= 0 && B.first[ea.col[l]] == B.first[ea.col[k]])
                B.M[(size_t)ea.col[k] * kMaxParams + (ea.col[l] - B.first[ea.col[l]])] += ea.Jc[0][k] * ea.Jc[0][l] + ea.Jc[1][k] * ea.Jc[1][l];
          }
          if (!Vi) continue;
          double Ya[3 * kNC];
          for (int k = 0; k < kNC; ++k) {
            double w[3];
            for (int l = 0; l < 3; ++l) w[l] = ea.Jc[0][k] * ea.Jp[0][l] + ea.Jc[1][k] * ea.Jp[1][l];
            for (int l = 0; l < 3; ++l) Ya[3 * k + l] = w[0] * Vi[l] + w[1] * Vi[3 + l] + w[2] * Vi[6 + l];
          }
          for (long ob = o0; ob < o1; ++ob) {
            const ObsJ& eb = sv.J[ob];
            for (int k = 0; k < kNC; ++k) {
              if (ea.col[k] < 0) continue;
              for (int l = 0; l < kNC; ++l) {
                if (eb.col[l] < 0 || B.first[eb.col[l]] != B.first[ea.col[k]]) continue;
                double wb[3];
                for (int m = 0; m < 3; ++m) wb[m] = eb.Jc[0][l] * eb.Jp[0][m] + eb.Jc[1][l] * eb.Jp[1][m];
                B.M[(size_t)ea.col[k] * kMaxParams + (eb.col[l] - B.first[eb.col[l]])] -= Ya[3 * k] * wb[0] + Ya[3 * k + 1] * wb[1] + Ya[3 * k + 2] * wb[2];
              }
            }
          }
        }
      }
      ok = InvertBlocks(B, D);
      // ConjugateGradientsSolver::Solve (conjugate_gradients_solver.cc), x0 = 0
      std::vector<double> x(std::max(D, 1), 0.0);
      if (ok && D > 0) {
        std::vector<double> r(D), pvec(D), z(D), tmp(D);
        double norm_b = 0;
        for (int j = 0; j < D; ++j) norm_b += rhs[j] * rhs[j];
        norm_b = std::sqrt(norm_b);
        if (norm_b != 0.0) {
          const double q_tolerance = 0.1, tol_r = -1.0 * norm_b;
          const int residual_reset_period = 10, min_num_iterations = 0;
          SchurMul(x, tmp);
          for (int j = 0; j < D; ++j) r[j] = rhs[j] - tmp[j];
          double rho = 1.0, Q0 = 0;
          for (int j = 0; j < D; ++j) Q0 += x[j] * (rhs[j] + r[j]);
          Q0 = -1.0 * Q0;
          for (int it = 1;; ++it) {
            S->num_linear_iterations++;
            ApplyBlocks(B, D, r.data(), z.data());
            const double last_rho = rho;
            rho = 0;
            for (int j = 0; j < D; ++j) rho += r[j] * z[j];
            if (rho == 0.0 || std::isinf(rho)) { ok = false; break; }  // LINEAR_SOLVER_FAILURE
            if (it == 1) {
              pvec = z;
            } else {
              const double beta = rho / last_rho;
              if (beta == 0.0 || std::isinf(beta)) { ok = false; break; }
              for (int j = 0; j < D; ++j) pvec[j] = z[j] + beta * pvec[j];
            }
            std::vector<double>& q = z;
            SchurMul(pvec, q);
            double pq = 0;
            for (int j = 0; j < D; ++j) pq += pvec[j] * q[j];
            if (pq <= 0 || std::isinf(pq)) break;  // NO_CONVERGENCE: the current x is still used
            const double alpha = rho / pq;
            if (std::isinf(alpha)) { ok = false; break; }
            for (int j = 0; j < D; ++j) x[j] = x[j] + alpha * pvec[j];
            if (it % residual_reset_period == 0) {
              SchurMul(x, tmp);
              for (int j = 0; j < D; ++j) r[j] = rhs[j] - tmp[j];
            } else {
              for (int j = 0; j < D; ++j) r[j] = r[j] - alpha * q[j];
            }
            double Q1 = 0;
            for (int j = 0; j < D; ++j) Q1 += x[j] * (rhs[j] + r[j]);
            Q1 = -1.0 * Q1;
            const double zeta = it * (Q1 - Q0) / Q1;
            if (zeta < q_tolerance && it >= min_num_iterations) break;
            Q0 = Q1;
            double norm_r = 0;
            for (int j = 0; j < D; ++j) norm_r += r[j] * r[j];
            norm_r = std::sqrt(norm_r);
            if (norm_r <= tol_r && it >= min_num_iterations) break;
            if (it >= O.max_linear_solver_iterations) break;
          }
        }
        for (int j = 0; j < D; ++j) if (!std::isfinite(x[j])) ok = false;  // IsArrayValid
      }
      for (int j = 0; j < D; ++j) dc[j] = -x[j];
    } else {
    // Schur complement: S = U + D_c - sum_p W V^-1 W^T ; rhs = g_c - sum_p W V^-1 g_p
    Smat.assign((size_t)D * D, 0.0);
    std::vector<double> rhs(gc.begin(), gc.begin() + D);
    for (long o = 0; o < P.n_obs; ++o) {  // U blocks
      const ObsJ& e = sv.J[o];
      for (int k = 0; k < kNC; ++k) {
        if (e.col[k] < 0) continue;
        for (int l = 0; l < kNC; ++l) {
          if (e.col[l] < 0) continue;
          Smat[(size_t)e.col[k] * D + e.col[l]] += e.Jc[0][k] * e.Jc[0][l] + e.Jc[1][k] * e.Jc[1][l];
        }
      }
    }
    for (int j = 0; j < D; ++j) Smat[(size_t)j * D + j] += lm_c[j];
#pragma omp parallel for schedule(dynamic, 64)
    for (int p = 0; p < P.n_pts; ++p) {
      const int pc = L.pt_col[p];
      if (pc < 0) continue;
      double V[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (long o = sv.pt_start[p]; o < sv.pt_start[p + 1]; ++o) {
        const ObsJ& e = sv.J[o];
        for (int a = 0; a < 2; ++a)
          for (int k = 0; k < 3; ++k)
            for (int l = 0; l < 3; ++l) V[k][l] += e.Jp[a][k] * e.Jp[a][l];
      }
      for (int k = 0; k < 3; ++k) V[k][k] += lm_p[3 * pc + k];
      // inverse of the symmetric 3x3
      const double c00 = V[1][1] * V[2][2] - V[1][2] * V[2][1], c01 = V[1][2] * V[2][0] - V[1][0] * V[2][2],
                   c02 = V[1][0] * V[2][1] - V[1][1] * V[2][0];
      const double det = V[0][0] * c00 + V[0][1] * c01 + V[0][2] * c02, id = 1.0 / det;
      double* Vi = &Vinv[(size_t)pc * 9];
      Vi[0] = c00 * id; Vi[1] = (V[0][2] * V[2][1] - V[0][1] * V[2][2]) * id; Vi[2] = (V[0][1] * V[1][2] - V[0][2] * V[1][1]) * id;
      Vi[3] = c01 * id; Vi[4] = (V[0][0] * V[2][2] - V[0][2] * V[2][0]) * id; Vi[5] = (V[0][2] * V[1][0] - V[0][0] * V[1][2]) * id;
      Vi[6] = c02 * id; Vi[7] = (V[0][1] * V[2][0] - V[0][0] * V[2][1]) * id; Vi[8] = (V[0][0] * V[1][1] - V[0][1] * V[1][0]) * id;
      double tp[3];
      for (int k = 0; k < 3; ++k) tp[k] = Vi[3 * k] * gp[3 * pc] + Vi[3 * k + 1] * gp[3 * pc + 1] + Vi[3 * k + 2] * gp[3 * pc + 2];
      const long o0 = sv.pt_start[p], nL = sv.pt_start[p + 1] - o0;
      std::vector<double> W((size_t)nL * 3 * kNC), Y((size_t)nL * 3 * kNC);
      for (long a = 0; a < nL; ++a) {
        const ObsJ& e = sv.J[o0 + a];
        double* Wa = &W[a * 3 * kNC];
        double* Ya = &Y[a * 3 * kNC];
        for (int k = 0; k < kNC; ++k)
          for (int l = 0; l < 3; ++l) Wa[3 * k + l] = e.Jc[0][k] * e.Jp[0][l] + e.Jc[1][k] * e.Jp[1][l];
        for (int k = 0; k < kNC; ++k)
          for (int l = 0; l < 3; ++l) Ya[3 * k + l] = Wa[3 * k] * Vi[l] + Wa[3 * k + 1] * Vi[3 + l] + Wa[3 * k + 2] * Vi[6 + l];
        for (int k = 0; k < kNC; ++k) {
          if (e.col[k] < 0) continue;
          const double v = Wa[3 * k] * tp[0] + Wa[3 * k + 1] * tp[1] + Wa[3 * k + 2] * tp[2];
#pragma omp atomic
          rhs[e.col[k]] -= v;
        }
      }
      for (long a = 0; a < nL; ++a) {
        const ObsJ& ea = sv.J[o0 + a];
        const double* Ya = &Y[a * 3 * kNC];
        for (long b = 0; b < nL; ++b) {
          const ObsJ& eb = sv.J[o0 + b];
          const double* Wb = &W[b * 3 * kNC];
          for (int k = 0; k < kNC; ++k) {
            if (ea.col[k] < 0) continue;
            for (int l = 0; l < kNC; ++l) {
              if (eb.col[l] < 0) continue;
              const double v = Ya[3 * k] * Wb[3 * l] + Ya[3 * k + 1] * Wb[3 * l + 1] + Ya[3 * k + 2] * Wb[3 * l + 2];
#pragma omp atomic
              Smat[(size_t)ea.col[k] * D + eb.col[l]] -= v;
            }
          }
        }
      }
    }
    // solve S y = rhs ; step_c = -y
    if (D > 0) {
      dc.assign(rhs.begin(), rhs.end());
      ok = CholeskySolve(Smat, D, dc);
      for (int j = 0; j < D; ++j) dc[j] = -dc[j];
    }
    }  // exact Schur step
    if (ok) {
      // back-substitution: dp = -V^-1 (g_p + W^T dc)
#pragma omp parallel for schedule(static)
      for (int p = 0; p < P.n_pts; ++p) {
        const int pc = L.pt_col[p];
        if (pc < 0) continue;
        double s[3] = {gp[3 * pc], gp[3 * pc + 1], gp[3 * pc + 2]};
        for (long o = sv.pt_start[p]; o < sv.pt_start[p + 1]; ++o) {
          const ObsJ& e = sv.J[o];
          for (int a = 0; a < 2; ++a) {
            double jd = 0;
            for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) jd += e.Jc[a][k] * dc[e.col[k]];
            for (int k = 0; k < 3; ++k) s[k] += e.Jp[a][k] * jd;
          }
        }
        const double* Vi = &Vinv[(size_t)pc * 9];
        for (int k = 0; k < 3; ++k) dp[3 * pc + k] = -(Vi[3 * k] * s[0] + Vi[3 * k + 1] * s[1] + Vi[3 * k + 2] * s[2]);
      }
    }
    double model_cost_change = 0;
    if (ok) {
      for (long o = 0; o < P.n_obs; ++o) {
        const ObsJ& e = sv.J[o];
        const int pc = L.pt_col[P.obs_pt[o]];
        for (int a = 0; a < 2; ++a) {
          double m = 0;
          for (int k = 0; k < kNC; ++k) if (e.col[k] >= 0) m += e.Jc[a][k] * dc[e.col[k]];
          if (pc >= 0) for (int k = 0; k < 3; ++k) m += e.Jp[a][k] * dp[3 * pc + k];
          model_cost_change -= m * (e.r[a] + m / 2.0);
        }
      }
      if (!(model_cost_change > 0)) ok = false;
    }
    if (!ok) {  // invalid step: shrink the trust region
      S->num_unsuccessful_steps++;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      // TrustRegionMinimizer::MinTrustRegionRadiusReached: CONVERGENCE ("Minimum trust region radius reached"), checked
      // after MaxSolverIterationsReached in FinalizeIterationAndCheckIfMinimizerCanContinue
      if (radius <= min_radius) { S->termination = (iter + 1 >= O.max_num_iterations) ? 1 : 0; break; }
      continue;
    }
    // candidate x_plus_delta (undo the Jacobi scaling)
    double step_sq = 0, x_sq = 0;
    qn.assign(P.qvec, P.qvec + P.n_img * 4);
    tn.assign(P.tvec, P.tvec + P.n_img * 3);
    kn.assign(P.cam_params, P.cam_params + P.n_cam * P.cam_stride);
    Xn.assign(P.xyz, P.xyz + P.n_pts * 3);
    for (int i = 0; i < P.n_img; ++i) {
      double d[3] = {0, 0, 0};
      bool any = false;
      for (int k = 0; k < 3; ++k) { const int c = L.pose_col[6 * i + k]; if (c >= 0) { d[k] = dc[c] * sv.scale_c[c]; any = true; step_sq += d[k] * d[k]; } }
      if (any) QuatPlus(P.qvec + 4 * i, d, &qn[4 * i]);
      for (int k = 0; k < 3; ++k) { const int c = L.pose_col[6 * i + 3 + k]; if (c >= 0) { const double v = dc[c] * sv.scale_c[c]; tn[3 * i + k] += v; step_sq += v * v; } }
    }
    for (int cm = 0; cm < P.n_cam; ++cm)
      for (int k = 0; k < kMaxParams; ++k) { const int c = L.intr_col[kMaxParams * cm + k]; if (c >= 0) { const double v = dc[c] * sv.scale_c[c]; kn[P.cam_stride * cm + k] += v; step_sq += v * v; } }
    for (int p = 0; p < P.n_pts; ++p) {
      const int pc = L.pt_col[p];
      if (pc < 0) continue;
      for (int k = 0; k < 3; ++k) { const double v = dp[3 * pc + k] * sv.scale_p[3 * pc + k]; Xn[3 * p + k] += v; step_sq += v * v; }
    }
    for (int i = 0; i < P.n_img * 4; ++i) x_sq += P.qvec[i] * P.qvec[i];
    for (int i = 0; i < P.n_img * 3; ++i) x_sq += P.tvec[i] * P.tvec[i];
    for (int i = 0; i < P.n_cam * P.cam_stride; ++i) x_sq += P.cam_params[i] * P.cam_params[i];
    for (int i = 0; i < P.n_pts * 3; ++i) x_sq += P.xyz[i] * P.xyz[i];
    if (std::sqrt(step_sq) <= O.parameter_tolerance * (std::sqrt(x_sq) + O.parameter_tolerance)) { S->termination = 0; break; }
    const double new_cost = sv.EvalCost(qn.data(), tn.data(), kn.data(), Xn.data());
    const double cost_change = cost - new_cost;
    // TrustRegionMinimizer::Minimize (Ceres 1.12+): FunctionToleranceReached is tested on every valid step BEFORE the
    // step is judged -- |cost_change| <= function_tolerance * cost ends the solve with CONVERGENCE and the candidate is
    // not applied (with the reference's function_tolerance = 0 this fires on an exactly unchanged cost)
    if (std::abs(cost_change) <= O.function_tolerance * cost) { S->termination = 0; break; }
    const double rho = cost_change / model_cost_change;
    if (rho > min_rel_dec) {  // successful step
      memcpy(P.qvec, qn.data(), qn.size() * 8);
      memcpy(P.tvec, tn.data(), tn.size() * 8);
      memcpy(P.cam_params, kn.data(), kn.size() * 8);
      memcpy(P.xyz, Xn.data(), Xn.size() * 8);
      S->num_successful_steps++;
      const double t = 2.0 * rho - 1.0;
      radius = radius / std::max(1.0 / 3.0, 1.0 - t * t * t);
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      cost = new_cost;
      sv.EvalJacobian(true);
      need_grad_check = true;
    } else {
      S->num_unsuccessful_steps++;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      // TrustRegionMinimizer::MinTrustRegionRadiusReached: CONVERGENCE ("Minimum trust region radius reached"), checked
      // after MaxSolverIterationsReached in FinalizeIterationAndCheckIfMinimizerCanContinue
      if (radius <= min_radius) { S->termination = (iter + 1 >= O.max_num_iterations) ? 1 : 0; break; }
    }
  }
  S->final_cost = cost;
  S->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
}

// B8: Reconstruction::ComputeMeanReprojectionError (base/reconstruction.cc:814-858) with
// CalculateSquaredReprojectionError (base/projection.cc:119-136) and QuaternionRotatePoint (base/pose.cc:110-116:
// Eigen::Quaterniond * Vector3d on the normalised qvec) over the tracks of the problem (its CSR rows), points in
// ascending id order as the sequential sums of the reference run.
static double MeanReprojectionError(const Problem& P, double* point_errors) {
  std::vector<long> pt_start(P.n_pts + 1, 0);
  for (long o = 0; o < P.n_obs; ++o) pt_start[P.obs_pt[o] + 1]++;
  for (int p = 0; p < P.n_pts; ++p) pt_start[p + 1] += pt_start[p];
  long total_reprojected_points = 0;
  double mean_reproj_error = 0.0;
  for (int p = 0; p < P.n_pts; ++p) {
    double reproj_error_sum = 0.0;
    const double* X = P.xyz + 3 * p;
    for (long o = pt_start[p]; o < pt_start[p + 1]; ++o) {
      const int i = P.obs_img[o], c = P.img_cam[i];
      double q[4] = {P.qvec[4 * i], P.qvec[4 * i + 1], P.qvec[4 * i + 2], P.qvec[4 * i + 3]};
      const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
      if (n == 0) q[0] = 1.0;  // NormalizeQuaternion
      else for (double& v : q) v /= n;
      // Eigen's QuaternionBase::_transformVector: uv = 2 * vec x v; v + w * uv + vec x uv
      const double uv[3] = {2.0 * (q[2] * X[2] - q[3] * X[1]), 2.0 * (q[3] * X[0] - q[1] * X[2]), 2.0 * (q[1] * X[1] - q[2] * X[0])};
      const double pw[3] = {X[0] + q[0] * uv[0] + (q[2] * uv[2] - q[3] * uv[1]) + P.tvec[3 * i],
                            X[1] + q[0] * uv[1] + (q[3] * uv[0] - q[1] * uv[2]) + P.tvec[3 * i + 1],
                            X[2] + q[0] * uv[2] + (q[1] * uv[1] - q[2] * uv[0]) + P.tvec[3 * i + 2]};
      if (pw[2] < std::numeric_limits<double>::epsilon()) continue;  // squared error = max(): skipped
      double x, y;
      const double u = pw[0] / pw[2], v = pw[1] / pw[2];
      WorldToImageT<double>(P.cam_model[c], P.cam_params + (long)P.cam_stride * c, u, v, &x, &y);
      const double dx = x - P.obs_xy[2 * o], dy = y - P.obs_xy[2 * o + 1];
      reproj_error_sum += std::sqrt(dx * dx + dy * dy);
    }
    const long L = pt_start[p + 1] - pt_start[p];
    if (point_errors) point_errors[p] = L > 0 ? reproj_error_sum / L : 0.0;
    total_reprojected_points += L;
    mean_reproj_error += reproj_error_sum;
  }
  return total_reprojected_points ? mean_reproj_error / static_cast<double>(total_reprojected_points) : 0.0;
}

}  // namespace ba

extern "C" {

struct orc_ba_problem {
  int32_t n_img, n_cam, n_pts; int64_t n_obs;
  double* qvec; double* tvec; const int32_t* img_cam; const uint8_t* pose_const; const uint8_t* tvec_const;
  const int32_t* cam_model; double* cam_params; const uint8_t* cam_const;
  int32_t refine_focal, refine_principal, refine_extra;
  double* xyz; const uint8_t* pt_const;
  const int32_t* obs_img; const int32_t* obs_pt; const double* obs_xy;
  int32_t cam_stride;  // doubles per camera in cam_params; 0 = 4
};
struct orc_ba_options { int32_t max_num_iterations; double function_tolerance, gradient_tolerance, parameter_tolerance; int32_t n_threads; int32_t loss_type; double loss_scale; int32_t linear_solver, max_linear_solver_iterations; };
struct orc_ba_summary { double initial_cost, final_cost; int32_t num_successful_steps, num_unsuccessful_steps, termination, num_residuals, num_effective_parameters; double seconds; int64_t num_linear_iterations; };

void orc_ba_solve(const orc_ba_problem* p, const orc_ba_options* o, orc_ba_summary* s) {
  ba::Problem P;
  P.n_img = p->n_img; P.n_cam = p->n_cam; P.n_pts = p->n_pts; P.n_obs = (long)p->n_obs;
  P.qvec = p->qvec; P.tvec = p->tvec; P.img_cam = p->img_cam; P.pose_const = p->pose_const; P.tvec_const = p->tvec_const;
  P.cam_model = p->cam_model; P.cam_params = p->cam_params; P.cam_const = p->cam_const;
  P.refine_focal = p->refine_focal; P.refine_principal = p->refine_principal; P.refine_extra = p->refine_extra;
  P.xyz = p->xyz; P.pt_const = p->pt_const; P.obs_img = p->obs_img; P.obs_pt = p->obs_pt; P.obs_xy = p->obs_xy;
  P.cam_stride = p->cam_stride > 0 ? p->cam_stride : 4;
  ba::Options O;
  O.max_num_iterations = o->max_num_iterations; O.function_tolerance = o->function_tolerance;
  O.gradient_tolerance = o->gradient_tolerance; O.parameter_tolerance = o->parameter_tolerance; O.n_threads = o->n_threads;
  O.loss_type = o->loss_type; O.loss_scale = o->loss_scale;
  O.linear_solver = o->linear_solver; O.max_linear_solver_iterations = o->max_linear_solver_iterations;
  ba::Summary S;
  ba::Solve(P, O, &S);
  s->initial_cost = S.initial_cost; s->final_cost = S.final_cost;
  s->num_successful_steps = S.num_successful_steps; s->num_unsuccessful_steps = S.num_unsuccessful_steps;
  s->termination = S.termination; s->num_residuals = S.num_residuals; s->num_effective_parameters = S.num_effective_parameters;
  s->seconds = S.seconds;
  s->num_linear_iterations = S.num_linear_iterations;
}

double orc_ba_mean_reprojection_error(const orc_ba_problem* p, double* point_errors) {
  ba::Problem P;
  P.n_img = p->n_img; P.n_cam = p->n_cam; P.n_pts = p->n_pts; P.n_obs = (long)p->n_obs;
  P.qvec = p->qvec; P.tvec = p->tvec; P.img_cam = p->img_cam; P.cam_model = p->cam_model; P.cam_params = p->cam_params;
  P.xyz = p->xyz; P.obs_img = p->obs_img; P.obs_pt = p->obs_pt; P.obs_xy = p->obs_xy;
  P.cam_stride = p->cam_stride > 0 ? p->cam_stride : 4;
  return ba::MeanReprojectionError(P, point_errors);
}

// residual + local Jacobians of one observation (tests: cost_functions_test.cc goldens, finite differences)
void orc_ba_evaluate(int model, const double* q, const double* t, const double* X, const double* k, const double* obs,
                     double* r, double* Jq, double* Jt, double* JX, double* Jk) {
  ba::Evaluate(model, q, t, X, k, obs, r, Jq, Jt, JX, Jk);
}
// loss value / derivatives and the corrector's scalings (tests: tests/test_oracle_ba_loss.py, host_ba_loss.cc)
void orc_ba_loss(int type, double a, double s, double* rho, double* residual_scaling, double* sqrt_rho1, double* alpha_sq_norm) {
  ba::LossEvaluate(type, a, s, rho);
  *residual_scaling = ba::Corrector(s, rho, sqrt_rho1, alpha_sq_norm);
}
void orc_ba_quat_plus(const double* x, const double* d, double* out) { ba::QuatPlus(x, d, out); }

}  // extern "C"
