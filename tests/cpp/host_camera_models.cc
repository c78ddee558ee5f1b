// dagsfm_b200/csrc/camera_models.cuh compiled for the host (tests/test_camera_models.py).
#include "camera_models.cuh"
extern "C" {
void host_cam_image_to_world(int model, const double* p, int n, const double* xy, double* out) {
  for (int i = 0; i < n; ++i) b2::cam::image_to_world(model, p, xy[2 * i], xy[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}
void host_cam_world_to_image(int model, const double* p, int n, const double* uv, double* out) {
  for (int i = 0; i < n; ++i) b2::cam::world_to_image(model, p, uv[2 * i], uv[2 * i + 1], &out[2 * i], &out[2 * i + 1]);
}
int host_cam_num_params(int model) { return b2::cam::num_params(model); }
double host_cam_mean_focal(int model, const double* p) { return b2::cam::mean_focal_length(model, p); }
}
