// dagsfm_b200/csrc/camera_jets.cuh compiled for the host (tests/test_camera_jets.py).
#include "camera_jets.cuh"
extern "C" {
// x, y and their derivatives: out[0..1] = (x, y); J[0..2*(2+K)) row-major 2 x (2 + 12): d/du, d/dv, d/dparams
void host_cam_world_to_image_jet(int model, const double* p, double u, double v, double* out, double* J) {
  b2::cam::Jet<14> x, y;
  b2::cam::world_to_image_jet<14>(model, p, u, v, &x, &y);
  out[0] = x.a; out[1] = y.a;
  for (int i = 0; i < 14; ++i) { J[i] = x.v[i]; J[14 + i] = y.v[i]; }
}
}
