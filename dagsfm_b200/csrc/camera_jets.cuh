// Forward-mode derivatives of CameraModel::WorldToImage for the bundle adjuster's general camera path.
// The reference differentiates BundleAdjustmentCostFunction (src/base/cost_functions.h:57-84) with Ceres' Jets
// (autodiff); for SIMPLE_PINHOLE / PINHOLE / SIMPLE_RADIAL the adjuster uses hand-derived formulas
// (ba_kernels.cu: evaluate), for the other eight models (src/base/camera_models.h:784-1482) it evaluates the
// model once on dual numbers: value + the partial derivatives with respect to (u, v, params[0..K)).
// Host-compilable (tests compare with central differences and with the oracle's own dual-number evaluation).
#pragma once
#include <cfloat>
#include <cmath>

#include "camera_models.cuh"

namespace b2 {
namespace cam {

template <int N>
struct Jet {
  double a;     // value
  double v[N];  // partial derivatives
};
template <int N> B2_CAM_HD Jet<N> jet_const(double c) {
  Jet<N> r;
  r.a = c;
  for (int i = 0; i < N; ++i) r.v[i] = 0.0;
  return r;
}
template <int N> B2_CAM_HD Jet<N> jet_var(double c, int k) {
  Jet<N> r = jet_const<N>(c);
  r.v[k] = 1.0;
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a + y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i];
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a - y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i];
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  r.a = x.a * y.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a;
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r;
  const double inv = 1.0 / y.a;
  r.a = x.a * inv;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv;  // ceres/jet.h: (f'g - f g'/g)/g
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator+(const Jet<N>& x, double c) { Jet<N> r = x; r.a += c; return r; }
template <int N> B2_CAM_HD Jet<N> operator+(double c, const Jet<N>& x) { return x + c; }
template <int N> B2_CAM_HD Jet<N> operator-(const Jet<N>& x, double c) { Jet<N> r = x; r.a -= c; return r; }
template <int N> B2_CAM_HD Jet<N> operator*(const Jet<N>& x, double c) {
  Jet<N> r;
  r.a = x.a * c;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * c;
  return r;
}
template <int N> B2_CAM_HD Jet<N> operator*(double c, const Jet<N>& x) { return x * c; }
template <int N> B2_CAM_HD Jet<N> jsqrt(const Jet<N>& x) {
  Jet<N> r;
  r.a = sqrt(x.a);
  const double d = 1.0 / (2.0 * r.a);
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d;
  return r;
}
template <int N> B2_CAM_HD Jet<N> jatan(const Jet<N>& x) {
  Jet<N> r;
  r.a = atan(x.a);
  const double d = 1.0 / (1.0 + x.a * x.a);
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d;
  return r;
}
template <int N> B2_CAM_HD Jet<N> jtan(const Jet<N>& x) {
  Jet<N> r;
  r.a = tan(x.a);
  const double d = 1.0 + r.a * r.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d;
  return r;
}

// WorldToImage of model `model` on Jets over (u, v, params): x, y with d/d(u, v) in v[0], v[1] and d/d params[k] in
// v[2 + k].  N must be >= 2 + num_params(model).  Branches depend on values only, exactly as in the reference's
// templated model functions.
template <int N>
B2_CAM_HD void world_to_image_jet(int model, const double* params, double u0, double v0, Jet<N>* x, Jet<N>* y) {
  const int K = num_params(model);
  Jet<N> p[12];
  for (int k = 0; k < 12; ++k) p[k] = (k < K) ? jet_var<N>(params[k], 2 + k) : jet_const<N>(0.0);
  const Jet<N> u = jet_var<N>(u0, 0), v = jet_var<N>(v0, 1);
  const int e = first_extra(model);
  Jet<N> a = u, b = v;
  if (model == 7) {  // FOVCameraModel::Distortion (:1137-1175)
    const Jet<N> omega = p[4];
    const Jet<N> radius2 = u * u + v * v, omega2 = omega * omega;
    Jet<N> factor;
    if (omega2.a < 1e-4) {
      factor = (omega2 * radius2) * (1.0 / 3.0) - omega2 * (1.0 / 12.0) + 1.0;
    } else if (radius2.a < 1e-4) {
      const Jet<N> th = jtan(omega * 0.5);
      factor = (th * -2.0 * (radius2 * 4.0 * th * th - 3.0)) / (omega * 3.0);
    } else {
      const Jet<N> radius = jsqrt(radius2);
      factor = jatan(radius * 2.0 * jtan(omega * 0.5)) / (radius * omega);
    }
    a = u * factor;
    b = v * factor;
  } else if (model >= 2) {
    if (model == 10) {  // ThinPrismFisheye: equidistant mapping first (:1412-1422)
      const Jet<N> r = jsqrt(u * u + v * v);
      if (r.a > DBL_EPSILON) {
        const Jet<N> theta = jatan(r);
        a = theta * u / r;
        b = theta * v / r;
      }
    }
    const Jet<N> u2 = a * a, v2 = b * b, r2 = u2 + v2;
    Jet<N> du = jet_const<N>(0.0), dv = jet_const<N>(0.0);
    if (model == 2) {
      const Jet<N> radial = p[e] * r2;
      du = a * radial; dv = b * radial;
    } else if (model == 3) {
      const Jet<N> radial = p[e] * r2 + p[e + 1] * r2 * r2;
      du = a * radial; dv = b * radial;
    } else if (model == 4) {
      const Jet<N> uv = a * b, radial = p[e] * r2 + p[e + 1] * r2 * r2;
      du = a * radial + p[e + 2] * 2.0 * uv + p[e + 3] * (r2 + u2 * 2.0);
      dv = b * radial + p[e + 3] * 2.0 * uv + p[e + 2] * (r2 + v2 * 2.0);
    } else if (model == 5 || model == 8 || model == 9) {
      const Jet<N> r = jsqrt(a * a + b * b);
      if (r.a > DBL_EPSILON) {
        const Jet<N> theta = jatan(r), theta2 = theta * theta, theta4 = theta2 * theta2;
        Jet<N> poly;
        if (model == 5) poly = p[e] * theta2 + p[e + 1] * theta4 + p[e + 2] * (theta4 * theta2) + p[e + 3] * (theta4 * theta4) + 1.0;
        else if (model == 8) poly = p[e] * theta2 + 1.0;
        else poly = p[e] * theta2 + p[e + 1] * theta4 + 1.0;
        const Jet<N> thetad = theta * poly;
        du = a * thetad / r - a;
        dv = b * thetad / r - b;
      }
    } else if (model == 6) {
      const Jet<N> uv = a * b, r4 = r2 * r2, r6 = r4 * r2;
      const Jet<N> radial = (p[e] * r2 + p[e + 1] * r4 + p[e + 4] * r6 + 1.0) / (p[e + 5] * r2 + p[e + 6] * r4 + p[e + 7] * r6 + 1.0);
      du = a * radial + p[e + 2] * 2.0 * uv + p[e + 3] * (r2 + u2 * 2.0) - a;
      dv = b * radial + p[e + 3] * 2.0 * uv + p[e + 2] * (r2 + v2 * 2.0) - b;
    } else if (model == 10) {
      const Jet<N> uv = a * b, r4 = r2 * r2, r6 = r4 * r2, r8 = r6 * r2;
      const Jet<N> radial = p[e] * r2 + p[e + 1] * r4 + p[e + 4] * r6 + p[e + 5] * r8;
      du = a * radial + p[e + 2] * 2.0 * uv + p[e + 3] * (r2 + u2 * 2.0) + p[e + 6] * r2;
      dv = b * radial + p[e + 3] * 2.0 * uv + p[e + 2] * (r2 + v2 * 2.0) + p[e + 7] * r2;
    }
    a = a + du;
    b = b + dv;
  }
  if (two_focal(model)) { *x = p[0] * a + p[2]; *y = p[1] * b + p[3]; }
  else { *x = p[0] * a + p[1]; *y = p[0] * b + p[2]; }
}

}  // namespace cam
}  // namespace b2
