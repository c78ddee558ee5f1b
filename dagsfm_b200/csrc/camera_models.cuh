// The eleven camera models of the reference (src/base/camera_models.h, ids :117-129) for the device:
// ImageToWorld incl. IterativeUndistortion (:547-590), WorldToImage, ImageToWorldThreshold (:535-543) and the
// layout helpers behind Camera::CalibrationMatrix (src/base/camera.cc:75-94).  Host-compilable (B2_CAM_HD) so the
// CPU suite can run the same source against the oracle.  FP64; compiled with --fmad=false where it is used.
//   0 SIMPLE_PINHOLE f,cx,cy | 1 PINHOLE fx,fy,cx,cy | 2 SIMPLE_RADIAL f,cx,cy,k | 3 RADIAL f,cx,cy,k1,k2
//   4 OPENCV fx,fy,cx,cy,k1,k2,p1,p2 | 5 OPENCV_FISHEYE fx,fy,cx,cy,k1..k4
//   6 FULL_OPENCV fx,fy,cx,cy,k1,k2,p1,p2,k3..k6 | 7 FOV fx,fy,cx,cy,omega
//   8 SIMPLE_RADIAL_FISHEYE f,cx,cy,k | 9 RADIAL_FISHEYE f,cx,cy,k1,k2
//   10 THIN_PRISM_FISHEYE fx,fy,cx,cy,k1,k2,p1,p2,k3,k4,sx1,sy1
#pragma once
#include <cfloat>
#include <cmath>

#if defined(__CUDACC__)
#define B2_CAM_HD __host__ __device__ inline
#else
#define B2_CAM_HD inline
#endif

namespace b2 {
namespace cam {

constexpr int kNumModels = 11;
B2_CAM_HD int num_params(int model) {
  switch (model) {
    case 0: return 3;
    case 1: case 2: case 8: return 4;
    case 3: case 7: case 9: return 5;
    case 4: case 5: return 8;
    case 6: case 10: return 12;
    default: return -1;
  }
}
// focal_length_idxs has two entries (fx, fy at 0, 1; principal point at 2, 3) or one (f at 0; cx, cy at 1, 2)
B2_CAM_HD bool two_focal(int model) { return model == 1 || (model >= 4 && model <= 7) || model == 10; }
B2_CAM_HD int first_extra(int model) { return two_focal(model) ? 4 : 3; }
B2_CAM_HD double mean_focal_length(int model, const double* p) { return two_focal(model) ? (p[0] + p[1]) / 2 : p[0]; }

// CameraModel::Distortion(extra_params, u, v, &du, &dv) of the models that undistort iteratively
B2_CAM_HD void distortion(int model, const double* e, double u, double v, double* du, double* dv) {
  const double u2 = u * u, v2 = v * v, r2 = u2 + v2;
  switch (model) {
    case 2: {  // :747-757
      const double radial = e[0] * r2;
      *du = u * radial; *dv = v * radial;
      return;
    }
    case 3: {  // :816-828
      const double radial = e[0] * r2 + e[1] * r2 * r2;
      *du = u * radial; *dv = v * radial;
      return;
    }
    case 4: {  // :888-903
      const double uv = u * v, radial = e[0] * r2 + e[1] * r2 * r2;
      *du = u * radial + 2.0 * e[2] * uv + e[3] * (r2 + 2.0 * u2);
      *dv = v * radial + 2.0 * e[3] * uv + e[2] * (r2 + 2.0 * v2);
      return;
    }
    case 5: case 8: case 9: {  // :963-986, :1272-1290, :1348-1368
      const double r = sqrt(u * u + v * v);
      if (r > DBL_EPSILON) {
        const double theta = atan(r), theta2 = theta * theta, theta4 = theta2 * theta2;
        double thetad;
        if (model == 5) {
          const double theta6 = theta4 * theta2, theta8 = theta4 * theta4;
          thetad = theta * (1.0 + e[0] * theta2 + e[1] * theta4 + e[2] * theta6 + e[3] * theta8);
        } else if (model == 8) {
          thetad = theta * (1.0 + e[0] * theta2);
        } else {
          thetad = theta * (1.0 + e[0] * theta2 + e[1] * theta4);
        }
        *du = u * thetad / r - u;
        *dv = v * thetad / r - v;
      } else {
        *du = 0.0; *dv = 0.0;
      }
      return;
    }
    case 6: {  // :1058-1080
      const double uv = u * v, r4 = r2 * r2, r6 = r4 * r2;
      const double radial = (1.0 + e[0] * r2 + e[1] * r4 + e[4] * r6) / (1.0 + e[5] * r2 + e[6] * r4 + e[7] * r6);
      *du = u * radial + 2.0 * e[2] * uv + e[3] * (r2 + 2.0 * u2) - u;
      *dv = v * radial + 2.0 * e[3] * uv + e[2] * (r2 + 2.0 * v2) - v;
      return;
    }
    case 10: {  // :1460-1482
      const double uv = u * v, r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
      const double radial = e[0] * r2 + e[1] * r4 + e[4] * r6 + e[5] * r8;
      *du = u * radial + 2.0 * e[2] * uv + e[3] * (r2 + 2.0 * u2) + e[6] * r2;
      *dv = v * radial + 2.0 * e[3] * uv + e[2] * (r2 + 2.0 * v2) + e[7] * r2;
      return;
    }
    default:
      *du = 0.0; *dv = 0.0;
  }
}

// FOVCameraModel::Distortion / Undistortion (:1137-1210) map the point itself
B2_CAM_HD void fov_map(double omega, double u, double v, bool undistort, double* ou, double* ov) {
  const double kEpsilon = 1e-4, radius2 = u * u + v * v, omega2 = omega * omega;
  double factor;
  if (omega2 < kEpsilon) {
    factor = (omega2 * radius2) / 3.0 - omega2 / 12.0 + 1.0;
  } else if (radius2 < kEpsilon) {
    const double tan_half_omega = tan(omega / 2.0);
    factor = undistort ? (omega * (omega * omega * radius2 + 3.0)) / (6.0 * tan_half_omega)
                       : (-2.0 * tan_half_omega * (4.0 * radius2 * tan_half_omega * tan_half_omega - 3.0)) / (3.0 * omega);
  } else {
    const double radius = sqrt(radius2);
    factor = undistort ? tan(radius * omega) / (radius * 2.0 * tan(omega / 2.0))
                       : atan(radius * 2.0 * tan(omega / 2.0)) / (radius * omega);
  }
  *ou = u * factor;
  *ov = v * factor;
}

// BaseCameraModel::IterativeUndistortion (:547-590): Newton with a central-difference Jacobian, <= 100 steps
B2_CAM_HD void iterative_undistortion(int model, const double* e, double* u, double* v) {
  const double x0_0 = *u, x0_1 = *v;
  double x_0 = *u, x_1 = *v;
  for (int i = 0; i < 100; ++i) {
    const double step0 = fmax(DBL_EPSILON, fabs(1e-6 * x_0));
    const double step1 = fmax(DBL_EPSILON, fabs(1e-6 * x_1));
    double dx0, dx1, b00, b01, f00, f01, b10, b11, f10, f11;
    distortion(model, e, x_0, x_1, &dx0, &dx1);
    distortion(model, e, x_0 - step0, x_1, &b00, &b01);
    distortion(model, e, x_0 + step0, x_1, &f00, &f01);
    distortion(model, e, x_0, x_1 - step1, &b10, &b11);
    distortion(model, e, x_0, x_1 + step1, &f10, &f11);
    const double J00 = 1 + (f00 - b00) / (2 * step0);
    const double J01 = (f10 - b10) / (2 * step1);
    const double J10 = (f01 - b01) / (2 * step0);
    const double J11 = 1 + (f11 - b11) / (2 * step1);
    const double invdet = 1.0 / (J00 * J11 - J01 * J10);
    const double r0 = x_0 + dx0 - x0_0, r1 = x_1 + dx1 - x0_1;
    const double s0 = (J11 * invdet) * r0 + (-J01 * invdet) * r1;
    const double s1 = (-J10 * invdet) * r0 + (J00 * invdet) * r1;
    x_0 -= s0;
    x_1 -= s1;
    if (s0 * s0 + s1 * s1 < 1e-10) break;
  }
  *u = x_0;
  *v = x_1;
}

// CameraModel::ImageToWorld of every model
B2_CAM_HD void image_to_world(int model, const double* p, double x, double y, double* u, double* v) {
  double a, b;
  if (two_focal(model)) { a = (x - p[2]) / p[0]; b = (y - p[3]) / p[1]; }
  else { a = (x - p[1]) / p[0]; b = (y - p[2]) / p[0]; }
  if (model == 7) {
    fov_map(p[4], a, b, true, &a, &b);
  } else if (model >= 2) {
    iterative_undistortion(model, p + first_extra(model), &a, &b);
    if (model == 10) {  // :1452-1458
      const double theta = sqrt(a * a + b * b);
      const double theta_cos_theta = theta * cos(theta);
      if (theta_cos_theta > DBL_EPSILON) {
        const double scale = sin(theta) / theta_cos_theta;
        a *= scale;
        b *= scale;
      }
    }
  }
  *u = a;
  *v = b;
}

// CameraModel::WorldToImage of every model
B2_CAM_HD void world_to_image(int model, const double* p, double u, double v, double* x, double* y) {
  double a = u, b = v;
  if (model == 7) {
    fov_map(p[4], u, v, false, &a, &b);
  } else if (model >= 2) {
    if (model == 10) {  // :1406-1435: equidistant mapping before the distortion
      const double r = sqrt(u * u + v * v);
      if (r > DBL_EPSILON) {
        const double theta = atan(r);
        a = theta * u / r;
        b = theta * v / r;
      }
    }
    double du, dv;
    distortion(model, p + first_extra(model), a, b, &du, &dv);
    a = a + du;
    b = b + dv;
  }
  if (two_focal(model)) { *x = p[0] * a + p[2]; *y = p[1] * b + p[3]; }
  else { *x = p[0] * a + p[1]; *y = p[0] * b + p[2]; }
}

}  // namespace cam
}  // namespace b2
