// Replays cases of the reference's src/optim/bundle_adjustment_test.cc on the C++ adaptor
// include/dagsfm_b200/bundle_adjustment_shim.hpp with a stand-in for colmap's Reconstruction (the members SetUp
// touches): TestConfigNumObservations, TestTwoView, TestTwoViewConstantCamera, TestPartiallyContainedTracks,
// TestPartiallyContainedTracksForceToOptimizePoint, TestConstantPoints, TestVariableImage -- the reduced residual /
// parameter counts of the solver summary and which blocks move.  argv[1] == "config" runs the host-only part.
// Built by tests/test_ba_shim.py against the product library (GPU) or the CUDA-emulator build (CPU).
#include <array>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <vector>

#include "dagsfm_b200/bundle_adjustment_shim.hpp"

using namespace dagsfm_b200;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

template <int N> struct Vec {
  std::array<double, N> v{};
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double operator()(int i) const { return v[i]; }
  bool operator==(const Vec& o) const { return v == o.v; }
  bool operator!=(const Vec& o) const { return v != o.v; }
};
struct TrackElement { image_t image_id; uint32_t point2D_idx; };
struct Trk {
  std::vector<TrackElement> els;
  size_t Length() const { return els.size(); }
  const std::vector<TrackElement>& Elements() const { return els; }
};
struct P2 {
  Vec<2> xy; point3D_t id = (point3D_t)-1;
  bool HasPoint3D() const { return id != (point3D_t)-1; }
  point3D_t Point3DId() const { return id; }
  const Vec<2>& XY() const { return xy; }
};
struct P3 {
  Vec<3> xyz; Trk track;
  Vec<3>& XYZ() { return xyz; }
  const Vec<3>& XYZ() const { return xyz; }
  const Trk& Track() const { return track; }
};
struct Cam {
  int model = 2; std::vector<double> params;
  int ModelId() const { return model; }
  size_t NumParams() const { return params.size(); }
  double* ParamsData() { return params.data(); }
  double Params(int i) const { return params[i]; }
};
struct Img {
  camera_t camera_id = 0; Vec<4> q; Vec<3> t; std::vector<P2> pts;
  camera_t CameraId() const { return camera_id; }
  void NormalizeQvec() { double n = 0; for (double x : q.v) n += x * x; n = std::sqrt(n); for (double& x : q.v) x /= n; }
  Vec<4>& Qvec() { return q; } const Vec<4>& Qvec() const { return q; }
  Vec<3>& Tvec() { return t; } const Vec<3>& Tvec() const { return t; }
  double Tvec(int i) const { return t.v[i]; }
  const std::vector<P2>& Points2D() const { return pts; }
  const P2& Point2D(uint32_t i) const { return pts[i]; }
};
struct Reconstruction {   // the slice of colmap::Reconstruction that BundleAdjuster::SetUp touches
  std::map<camera_t, Cam> cameras; std::map<image_t, Img> images; std::map<point3D_t, P3> points;
  Cam& Camera(camera_t id) { return cameras.at(id); }
  const Cam& Camera(camera_t id) const { return cameras.at(id); }
  Img& Image(image_t id) { return images.at(id); }
  const Img& Image(image_t id) const { return images.at(id); }
  P3& Point3D(point3D_t id) { return points.at(id); }
  const P3& Point3D(point3D_t id) const { return points.at(id); }
  const std::map<point3D_t, P3>& Points3D() const { return points; }
  void DeleteObservation(image_t image_id, uint32_t idx) {  // Reconstruction::DeleteObservation
    const point3D_t pid = images.at(image_id).pts[idx].id;
    images.at(image_id).pts[idx].id = (point3D_t)-1;
    auto& els = points.at(pid).track.els;
    for (size_t k = 0; k < els.size(); ++k) if (els[k].image_id == image_id && els[k].point2D_idx == idx) { els.erase(els.begin() + k); break; }
  }
};

// bundle_adjustment_test.cc:107-184 (point ids start at 1 as Reconstruction::AddPoint3D numbers them)
static void GenerateReconstruction(size_t num_images, size_t num_points, Reconstruction* r) {
  std::mt19937 prng(0);
  auto real = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(prng); };
  for (size_t i = 0; i < num_points; ++i) { P3 p; p.xyz.v = {real(-1, 1), real(-1, 1), real(-1, 1)}; r->points[i + 1] = p; }
  for (size_t i = 0; i < num_images; ++i) {
    Cam c; c.model = 2; c.params = {1.2 * 1000, 500, 500, 0};
    r->cameras[(camera_t)i] = c;
    Img im; im.camera_id = (camera_t)i; im.q.v = {1, 0, 0, 0}; im.t.v = {real(-1, 1), real(-1, 1), 10};
    uint32_t idx = 0;
    for (auto& kv : r->points) {
      const auto& X = kv.second.xyz.v;
      const double z = X[2] + im.t.v[2];
      P2 p2;
      p2.xy.v = {1200 * (X[0] + im.t.v[0]) / z + 500 + real(-2, 2), 1200 * (X[1] + im.t.v[1]) / z + 500 + real(-2, 2)};
      p2.id = kv.first;
      im.pts.push_back(p2);
      kv.second.track.els.push_back({(image_t)i, idx++});
    }
    r->images[(image_t)i] = im;
  }
}

static bool variable_camera(const Cam& c, const Cam& o) { return c.Params(0) != o.Params(0) && c.Params(1) == o.Params(1) && c.Params(2) == o.Params(2) && c.Params(3) != o.Params(3); }
static bool constant_camera(const Cam& c, const Cam& o) { return c.params == o.params; }
static bool variable_image(const Img& a, const Img& o) { return a.q != o.q && a.t != o.t; }
static bool constant_image(const Img& a, const Img& o) { return a.q == o.q && a.t == o.t; }

static int test_config() {  // TestConfigNumObservations (:186-209) and the container semantics (:bundle_adjustment_test.cc)
  Reconstruction r;
  GenerateReconstruction(4, 100, &r);
  BundleAdjustmentConfig config;
  config.AddImage(0); config.AddImage(1);
  CHECK(config.NumResiduals(r) == 400);
  config.AddVariablePoint(1);
  CHECK(config.NumResiduals(r) == 404);
  config.AddConstantPoint(2);
  CHECK(config.NumResiduals(r) == 408);
  config.AddImage(2);
  CHECK(config.NumResiduals(r) == 604);
  config.AddImage(3);
  CHECK(config.NumResiduals(r) == 800);
  CHECK(config.NumImages() == 4 && config.NumPoints() == 2 && config.NumVariablePoints() == 1 && config.NumConstantPoints() == 1);
  config.SetConstantPose(0);
  config.SetConstantTvec(1, {0});
  CHECK(config.HasConstantPose(0) && config.HasConstantTvec(1) && config.ConstantTvec(1) == std::vector<int>{0});
  CHECK(config.NumConstantPoses() == 1 && config.NumConstantTvecs() == 1);
  config.SetVariablePose(0); config.RemoveConstantTvec(1); config.RemoveImage(3); config.RemoveVariablePoint(1);
  CHECK(!config.HasConstantPose(0) && !config.HasConstantTvec(1) && config.NumImages() == 3 && !config.HasPoint(1) && config.HasPoint(2));
  BundleAdjustmentOptions o;   // defaults of bundle_adjustment.h:48-87
  CHECK(o.refine_focal_length && !o.refine_principal_point && o.refine_extra_params && o.refine_extrinsics);
  CHECK(o.solver_options.max_num_iterations == 100 && o.solver_options.max_linear_solver_iterations == 200 && o.solver_options.gradient_tolerance == 0.0);
  return 0;
}

static int test_solve() {
  BundleAdjustmentOptions options;
  options.print_summary = false;
  {  // TestTwoView (:211-247)
    Reconstruction r; GenerateReconstruction(2, 100, &r); const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantTvec(1, {0});
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 400 && ba.Summary().num_effective_parameters_reduced == 309);
    CHECK(variable_camera(r.Camera(0), orig.Camera(0)) && constant_image(r.Image(0), orig.Image(0)));
    CHECK(variable_camera(r.Camera(1), orig.Camera(1)) && variable_image(r.Image(1), orig.Image(1)) && r.Image(1).Tvec(0) == orig.Image(1).Tvec(0));
    for (const auto& kv : r.Points3D()) CHECK(kv.second.xyz != orig.Point3D(kv.first).xyz);
    CHECK(ba.Summary().final_cost < ba.Summary().initial_cost && ba.Summary().IsSolutionUsable() && !ba.Summary().iterative_schur);
  }
  {  // TestTwoViewConstantCamera (:249-284)
    Reconstruction r; GenerateReconstruction(2, 100, &r); const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantPose(1); config.SetConstantCamera(0);
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 400 && ba.Summary().num_effective_parameters_reduced == 302);
    CHECK(constant_camera(r.Camera(0), orig.Camera(0)) && constant_image(r.Image(0), orig.Image(0)));
    CHECK(variable_camera(r.Camera(1), orig.Camera(1)) && constant_image(r.Image(1), orig.Image(1)));
    for (const auto& kv : r.Points3D()) CHECK(kv.second.xyz != orig.Point3D(kv.first).xyz);
  }
  {  // TestPartiallyContainedTracks (:286-332)
    Reconstruction r; GenerateReconstruction(3, 100, &r);
    const point3D_t variable_id = r.Image(2).Point2D(0).Point3DId();
    r.DeleteObservation(2, 0);
    const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantPose(1);
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 400 && ba.Summary().num_effective_parameters_reduced == 7);
    CHECK(variable_camera(r.Camera(0), orig.Camera(0)) && variable_camera(r.Camera(1), orig.Camera(1)) && constant_camera(r.Camera(2), orig.Camera(2)));
    for (image_t i = 0; i < 3; ++i) CHECK(constant_image(r.Image(i), orig.Image(i)));
    for (const auto& kv : r.Points3D()) CHECK((kv.second.xyz != orig.Point3D(kv.first).xyz) == (kv.first == variable_id));
  }
  {  // TestPartiallyContainedTracksForceToOptimizePoint (:334-390)
    Reconstruction r; GenerateReconstruction(3, 100, &r);
    const point3D_t variable_id = r.Image(2).Point2D(0).Point3DId(), add_var = r.Image(2).Point2D(1).Point3DId(),
                    add_const = r.Image(2).Point2D(2).Point3DId();
    r.DeleteObservation(2, 0);
    const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantPose(1);
    config.AddVariablePoint(add_var); config.AddConstantPoint(add_const);
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 402 && ba.Summary().num_effective_parameters_reduced == 10);
    CHECK(constant_camera(r.Camera(2), orig.Camera(2)) && constant_image(r.Image(2), orig.Image(2)));
    for (const auto& kv : r.Points3D())
      CHECK((kv.second.xyz != orig.Point3D(kv.first).xyz) == (kv.first == variable_id || kv.first == add_var));
  }
  {  // TestConstantPoints (:392-434)
    Reconstruction r; GenerateReconstruction(2, 100, &r); const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantPose(1);
    config.AddConstantPoint(1); config.AddConstantPoint(2);
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 400 && ba.Summary().num_effective_parameters_reduced == 298);
    for (const auto& kv : r.Points3D()) CHECK((kv.second.xyz == orig.Point3D(kv.first).xyz) == (kv.first == 1 || kv.first == 2));
  }
  {  // TestVariableImage (:436-475)
    Reconstruction r; GenerateReconstruction(3, 100, &r); const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.AddImage(2); config.SetConstantPose(0); config.SetConstantTvec(1, {0});
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_residuals_reduced == 600 && ba.Summary().num_effective_parameters_reduced == 317);
    CHECK(constant_image(r.Image(0), orig.Image(0)) && variable_image(r.Image(2), orig.Image(2)));
  }
  {  // an OPENCV camera goes through the 12-slot layout; no observations -> Solve returns false (.cc:266-268)
    Reconstruction r; GenerateReconstruction(2, 60, &r);
    for (auto& kv : r.cameras) { kv.second.model = 4; kv.second.params = {1200, 1200, 500, 500, 0, 0, 0, 0}; }
    const Reconstruction orig = r;
    BundleAdjustmentConfig config; config.AddImage(0); config.AddImage(1); config.SetConstantPose(0); config.SetConstantTvec(1, {0});
    BundleAdjuster ba(options, config);
    CHECK(ba.Solve(&r));
    CHECK(ba.Summary().num_effective_parameters_reduced == 3 * 60 + 5 + 2 * 6 && ba.Summary().final_cost < ba.Summary().initial_cost);
    CHECK(r.Camera(0).Params(4) != orig.Camera(0).Params(4) && r.Camera(0).Params(2) == 500);
    BundleAdjustmentConfig empty;
    BundleAdjuster none(options, empty);
    Reconstruction r2; GenerateReconstruction(2, 10, &r2);
    CHECK(!none.Solve(&r2));
  }
  std::printf("ba shim ok\n");
  return 0;
}

int main(int argc, char** argv) {
  if (test_config()) return 1;
  std::printf("ba shim config tests ok\n");
  if (argc > 1 && !std::strcmp(argv[1], "config")) return 0;
  return test_solve();
}
