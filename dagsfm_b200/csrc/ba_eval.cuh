// The reprojection residual of one observation and its analytic Jacobian blocks -- shared by every bundle-adjustment
// kernel (ba_kernels.cu: staged ObsJac path; ba_fused.cu: kernels that recompute the blocks from the 24-byte observation
// record instead of streaming them).  Reference arithmetic: BundleAdjustmentCostFunction::operator()
// (src/base/cost_functions.h:57-84), SimpleRadialCameraModel::WorldToImage (src/base/camera_models.h:714-757),
// ceres::UnitQuaternionRotatePoint + QuaternionParameterization (Ceres 1.14, external).
#pragma once
#include <cuda_runtime.h>

#include "camera_jets.cuh"

namespace b2 {
namespace bak {

// One observation: residual and Jacobian blocks (only the residual if Jc == nullptr).  KI = intrinsics slots: 4 =
// the production instantiation (SIMPLE_PINHOLE / PINHOLE / SIMPLE_RADIAL, hand-derived formulas); 12 adds the other
// eight models of camera_models.h, differentiated on dual numbers (camera_jets.cuh) as Ceres' autodiff does.
template <int KI>
__device__ __forceinline__ void evaluate(int model, const double* q, const double* t, const double* X,
                                         const double* k, double ox, double oy, double* r, double* Jc /*2 x (6+KI)*/,
                                         double* Jp /*2x3*/) {
  constexpr int NC = 6 + KI;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double t2 = w * x, t3 = w * y, t4 = w * z, t5 = -x * x, t6 = x * y, t7 = x * z, t8 = -y * y, t9 = y * z,
               t1 = -z * z;
  double p0 = 2 * ((t8 + t1) * X[0] + (t6 - t4) * X[1] + (t3 + t7) * X[2]) + X[0];
  double p1 = 2 * ((t4 + t6) * X[0] + (t5 + t1) * X[1] + (t9 - t2) * X[2]) + X[1];
  double p2 = 2 * ((t7 - t3) * X[0] + (t2 + t9) * X[1] + (t5 + t8) * X[2]) + X[2];
  p0 += t[0];
  p1 += t[1];
  p2 += t[2];
  const double u = p0 / p2, v = p1 / p2;
  double xi, yi, dxdu, dxdv, dydu, dydv;
  double dk0[KI] = {0, 0, 0, 0}, dk1[KI] = {0, 0, 0, 0};  // remaining slots (KI = 12) are zero-initialised too
  if (model == 0) {
    xi = k[0] * u + k[1]; yi = k[0] * v + k[2];
    dxdu = k[0]; dxdv = 0; dydu = 0; dydv = k[0];
    dk0[0] = u; dk0[1] = 1; dk1[0] = v; dk1[2] = 1;
  } else if (model == 1) {
    xi = k[0] * u + k[2]; yi = k[1] * v + k[3];
    dxdu = k[0]; dxdv = 0; dydu = 0; dydv = k[1];
    dk0[0] = u; dk0[2] = 1; dk1[1] = v; dk1[3] = 1;
  } else if (KI == 4 || model == 2) {
    const double u2 = u * u, v2 = v * v, r2 = u2 + v2, radial = k[3] * r2;
    const double du = u * radial, dv = v * radial;
    const double xd = u + du, yd = v + dv;
    xi = k[0] * xd + k[1]; yi = k[0] * yd + k[2];
    dxdu = k[0] * (1 + radial + 2 * k[3] * u2); dxdv = k[0] * (2 * k[3] * u * v);
    dydu = k[0] * (2 * k[3] * u * v); dydv = k[0] * (1 + radial + 2 * k[3] * v2);
    dk0[0] = xd; dk0[1] = 1; dk0[3] = k[0] * u * r2;
    dk1[0] = yd; dk1[2] = 1; dk1[3] = k[0] * v * r2;
  } else {
    cam::Jet<2 + KI> xj, yj;
    cam::world_to_image_jet<2 + KI>(model, k, u, v, &xj, &yj);
    xi = xj.a; yi = yj.a;
    dxdu = xj.v[0]; dxdv = xj.v[1]; dydu = yj.v[0]; dydv = yj.v[1];
#pragma unroll
    for (int a = 0; a < KI; ++a) { dk0[a] = xj.v[2 + a]; dk1[a] = yj.v[2 + a]; }
  }
  r[0] = xi - ox;
  r[1] = yi - oy;
  if (!Jc) return;
  const double ip2 = 1.0 / p2;
  const double dudp[3] = {ip2, 0, -p0 * ip2 * ip2}, dvdp[3] = {0, ip2, -p1 * ip2 * ip2};
  double drdp[2][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    drdp[0][c] = dxdu * dudp[c] + dxdv * dvdp[c];
    drdp[1][c] = dydu * dudp[c] + dydv * dvdp[c];
  }
  const double R[3][3] = {{2 * (t8 + t1) + 1, 2 * (t6 - t4), 2 * (t3 + t7)},
                          {2 * (t4 + t6), 2 * (t5 + t1) + 1, 2 * (t9 - t2)},
                          {2 * (t7 - t3), 2 * (t2 + t9), 2 * (t5 + t8) + 1}};
  const double X0 = X[0], X1 = X[1], X2 = X[2];
  const double dpdq[3][4] = {
      {2 * (-z * X1 + y * X2), 2 * (y * X1 + z * X2), 2 * (-2 * y * X0 + x * X1 + w * X2), 2 * (-2 * z * X0 - w * X1 + x * X2)},
      {2 * (z * X0 - x * X2), 2 * (y * X0 - 2 * x * X1 - w * X2), 2 * (x * X0 + z * X2), 2 * (w * X0 - 2 * z * X1 + y * X2)},
      {2 * (-y * X0 + x * X1), 2 * (z * X0 + w * X1 - 2 * x * X2), 2 * (-w * X0 + z * X1 - 2 * y * X2), 2 * (x * X0 + y * X1)}};
  const double JL[4][3] = {{-x, -y, -z}, {w, z, -y}, {-z, w, x}, {y, -x, w}};
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    double drdq[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) drdq[a] = drdp[i][0] * dpdq[0][a] + drdp[i][1] * dpdq[1][a] + drdp[i][2] * dpdq[2][a];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      Jc[NC * i + c] = drdq[0] * JL[0][c] + drdq[1] * JL[1][c] + drdq[2] * JL[2][c] + drdq[3] * JL[3][c];
      Jc[NC * i + 3 + c] = drdp[i][c];
      Jp[3 * i + c] = drdp[i][0] * R[0][c] + drdp[i][1] * R[1][c] + drdp[i][2] * R[2][c];
    }
#pragma unroll
    for (int a = 0; a < KI; ++a) Jc[NC * i + 6 + a] = (i == 0) ? dk0[a] : dk1[a];
  }
}

}  // namespace bak
}  // namespace b2
