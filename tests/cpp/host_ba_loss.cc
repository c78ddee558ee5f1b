// Host build of the robust-loss function the Jacobian kernel calls (dagsfm_b200/csrc/ba_loss.cuh).
#include "../../dagsfm_b200/csrc/ba_loss.cuh"
extern "C" void host_ba_loss(int type, double a, double s, double* rho0, double* sqrt_rho1) {
  if (type == 1) b2::bak::loss_eval<1>(a, s, rho0, sqrt_rho1);
  else b2::bak::loss_eval<2>(a, s, rho0, sqrt_rho1);
}
