// Robust loss of one observation as the Jacobian kernel applies it (host-compilable: tests/cpp/host_ba_loss.cc
// checks it against the oracle's restatement of ceres::SoftLOneLoss / CauchyLoss + Corrector).
#pragma once
#include <cfloat>
#include <cmath>

#ifndef B2_HD
#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif
#endif

namespace b2 {
namespace bak {

template <int LOSS>
B2_HD void loss_eval(double a, double s, double* rho0, double* sqrt_rho1) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  if (LOSS == 1) {  // ceres::SoftLOneLoss
    const double tmp = sqrt(sum);
    *rho0 = 2.0 * b * (tmp - 1.0);
    *sqrt_rho1 = sqrt(fmax(DBL_MIN, 1.0 / tmp));
  } else {          // ceres::CauchyLoss
    const double inv = 1.0 / sum;
    *rho0 = b * log(sum);
    *sqrt_rho1 = sqrt(fmax(DBL_MIN, inv));
  }
}

}  // namespace bak
}  // namespace b2
