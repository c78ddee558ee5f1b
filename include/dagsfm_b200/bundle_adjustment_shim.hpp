// Header-only C++ adaptor that gives the bundle-adjustment C ABI (include/dagsfm_b200.h, BA section) the reference's
// own classes for this path, so that the reference's callers compile unchanged:
//
//   BundleAdjustmentOptions      src/optim/bundle_adjustment.h:48-103, .cc:53-73
//   BundleAdjustmentConfig       src/optim/bundle_adjustment.h:105-161, .cc:80-246
//   BundleAdjuster               src/optim/bundle_adjustment.h:165-197, .cc:253-526
//       ::Solve -> SetUp (AddImageToProblem, AddPointToProblem, ParameterizeCameras, ParameterizePoints) -> ceres::Solve
//
// Callers: DistributedMapperController::AdjustGlobalBundle (controllers/distributed_mapper_controller.cpp:836-933),
// BundleAdjustmentController::Run (controllers/bundle_adjustment.cc:69-103), IncrementalMapper::AdjustLocalBundle /
// AdjustGlobalBundle (sfm/incremental_mapper.cc:562-740).
//
// The reference cannot be built in this environment (Eigen / Ceres / glog absent), so the adaptor is a template over the
// members of Reconstruction / Image / Camera / Point2D / Point3D / Track it touches -- the ones SetUp touches:
//   reconstruction->Image(id): CameraId(), NormalizeQvec(), Qvec().data(), Tvec().data(), Points2D(), Point2D(idx)
//   Point2D: HasPoint3D(), Point3DId(), XY() (indexable with (0), (1))
//   reconstruction->Camera(id): ModelId(), NumParams(), ParamsData()
//   reconstruction->Point3D(id): XYZ().data(), Track().Length(), Track().Elements() -> {image_id, point2D_idx}
// Summary() returns the fields of ceres::Solver::Summary the reference reads (PrintSolverSummary, .cc:1148-1156).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {

using image_t = uint32_t;
using camera_t = uint32_t;
using point3D_t = uint64_t;

struct SolverOptions {  // the members of ceres::Solver::Options the reference sets or the adaptor honours
  double function_tolerance = 0.0;
  double gradient_tolerance = 0.0;
  double parameter_tolerance = 0.0;
  bool minimizer_progress_to_stdout = false;
  int max_num_iterations = 100;
  int max_linear_solver_iterations = 200;
  int max_num_consecutive_invalid_steps = 10;
  int max_consecutive_nonmonotonic_steps = 10;
  int num_threads = -1;
  int num_linear_solver_threads = -1;
};

struct BundleAdjustmentOptions {  // bundle_adjustment.h:48-103
  enum class LossFunctionType { TRIVIAL, SOFT_L1, CAUCHY };
  LossFunctionType loss_function_type = LossFunctionType::TRIVIAL;
  double loss_function_scale = 1.0;
  bool refine_focal_length = true;
  bool refine_principal_point = false;
  bool refine_extra_params = true;
  bool refine_extrinsics = true;
  bool print_summary = true;
  SolverOptions solver_options;
  int gpu_index = 0;  // not in the reference: the device that solves
  bool Check() const { return loss_function_scale >= 0; }
};

class BundleAdjustmentConfig {  // bundle_adjustment.h:105-161
 public:
  size_t NumImages() const { return image_ids_.size(); }
  size_t NumPoints() const { return variable_point3D_ids_.size() + constant_point3D_ids_.size(); }
  size_t NumConstantCameras() const { return constant_camera_ids_.size(); }
  size_t NumConstantPoses() const { return constant_poses_.size(); }
  size_t NumConstantTvecs() const { return constant_tvecs_.size(); }
  size_t NumVariablePoints() const { return variable_point3D_ids_.size(); }
  size_t NumConstantPoints() const { return constant_point3D_ids_.size(); }

  // .cc:207-246
  template <class Reconstruction>
  size_t NumResiduals(const Reconstruction& reconstruction) const {
    size_t num_observations = 0;
    for (const image_t image_id : image_ids_)
      for (const auto& point2D : reconstruction.Image(image_id).Points2D()) num_observations += point2D.HasPoint3D() ? 1 : 0;
    auto outside = [&](const point3D_t point3D_id) {
      size_t n = 0;
      for (const auto& el : reconstruction.Point3D(point3D_id).Track().Elements()) n += image_ids_.count(el.image_id) ? 0 : 1;
      return n;
    };
    for (const auto id : variable_point3D_ids_) num_observations += outside(id);
    for (const auto id : constant_point3D_ids_) num_observations += outside(id);
    return 2 * num_observations;
  }

  void AddImage(const image_t image_id) { image_ids_.insert(image_id); }
  bool HasImage(const image_t image_id) const { return image_ids_.count(image_id) > 0; }
  void RemoveImage(const image_t image_id) { image_ids_.erase(image_id); }

  void SetConstantCamera(const camera_t camera_id) { constant_camera_ids_.insert(camera_id); }
  void SetVariableCamera(const camera_t camera_id) { constant_camera_ids_.erase(camera_id); }
  bool IsConstantCamera(const camera_t camera_id) const { return constant_camera_ids_.count(camera_id) > 0; }

  void SetConstantPose(const image_t image_id) {
    Check(HasImage(image_id) && !HasConstantTvec(image_id), "SetConstantPose");  // .cc:150-154
    constant_poses_.insert(image_id);
  }
  void SetVariablePose(const image_t image_id) { constant_poses_.erase(image_id); }
  bool HasConstantPose(const image_t image_id) const { return constant_poses_.count(image_id) > 0; }

  void SetConstantTvec(const image_t image_id, const std::vector<int>& idxs) {
    Check(idxs.size() > 0 && idxs.size() <= 3 && HasImage(image_id) && !HasConstantPose(image_id), "SetConstantTvec");
    Check(std::set<int>(idxs.begin(), idxs.end()).size() == idxs.size(), "Tvec indices must not contain duplicates");
    constant_tvecs_[image_id] = idxs;
  }
  void RemoveConstantTvec(const image_t image_id) { constant_tvecs_.erase(image_id); }
  bool HasConstantTvec(const image_t image_id) const { return constant_tvecs_.count(image_id) > 0; }
  const std::vector<int>& ConstantTvec(const image_t image_id) const { return constant_tvecs_.at(image_id); }

  void AddVariablePoint(const point3D_t point3D_id) {
    Check(!HasConstantPoint(point3D_id), "AddVariablePoint");
    variable_point3D_ids_.insert(point3D_id);
  }
  void AddConstantPoint(const point3D_t point3D_id) {
    Check(!HasVariablePoint(point3D_id), "AddConstantPoint");
    constant_point3D_ids_.insert(point3D_id);
  }
  bool HasPoint(const point3D_t id) const { return HasVariablePoint(id) || HasConstantPoint(id); }
  bool HasVariablePoint(const point3D_t id) const { return variable_point3D_ids_.count(id) > 0; }
  bool HasConstantPoint(const point3D_t id) const { return constant_point3D_ids_.count(id) > 0; }
  void RemoveVariablePoint(const point3D_t id) { variable_point3D_ids_.erase(id); }
  void RemoveConstantPoint(const point3D_t id) { constant_point3D_ids_.erase(id); }

  const std::unordered_set<image_t>& Images() const { return image_ids_; }
  const std::unordered_set<point3D_t>& VariablePoints() const { return variable_point3D_ids_; }
  const std::unordered_set<point3D_t>& ConstantPoints() const { return constant_point3D_ids_; }

 private:
  static void Check(bool ok, const char* what) {  // glog CHECK in the reference: abort
    if (!ok) { std::fprintf(stderr, "Check failed: BundleAdjustmentConfig::%s\n", what); std::abort(); }
  }
  std::unordered_set<camera_t> constant_camera_ids_;
  std::unordered_set<image_t> image_ids_;
  std::unordered_set<point3D_t> variable_point3D_ids_;
  std::unordered_set<point3D_t> constant_point3D_ids_;
  std::unordered_set<image_t> constant_poses_;
  std::unordered_map<image_t, std::vector<int>> constant_tvecs_;
};

struct SolverSummary {  // ceres::Solver::Summary, the fields the reference reads
  double initial_cost = 0, final_cost = 0;
  int num_residuals_reduced = 0, num_effective_parameters_reduced = 0;
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  int termination_type = 1;  // 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE
  int num_linear_solver_iterations = 0;
  bool iterative_schur = false;
  double total_time_in_seconds = 0;
  bool IsSolutionUsable() const { return termination_type == 0 || termination_type == 1; }
};

class BundleAdjuster {
 public:
  BundleAdjuster(const BundleAdjustmentOptions& options, const BundleAdjustmentConfig& config)
      : options_(options), config_(config) {
    if (!options_.Check()) { std::fprintf(stderr, "Check failed: options_.Check()\n"); std::abort(); }
  }

  // bundle_adjustment.cc:258-310.  Returns false iff the problem has no residuals (or the device solve failed).
  template <class Reconstruction>
  bool Solve(Reconstruction* reconstruction) {
    if (used_) { std::fprintf(stderr, "Check failed: Cannot use the same BundleAdjuster multiple times\n"); std::abort(); }
    used_ = true;
    // ---- SetUp (.cc:316-332) into the flat arrays of b2_ba_problem; ids in ascending order
    std::vector<image_t> config_images(config_.Images().begin(), config_.Images().end());
    std::sort(config_images.begin(), config_images.end());
    struct Obs { point3D_t point3D_id; image_t image_id; double x, y; };
    std::vector<Obs> obs;
    std::map<point3D_t, size_t> point3D_num_observations;
    std::vector<image_t> prob_images;
    std::map<image_t, bool> constant_pose;
    std::set<camera_t> camera_ids;
    for (const image_t image_id : config_images) {  // AddImageToProblem (.cc:338-420)
      auto& image = reconstruction->Image(image_id);
      image.NormalizeQvec();
      const bool cp = !options_.refine_extrinsics || config_.HasConstantPose(image_id);
      size_t num_observations = 0;
      for (const auto& point2D : image.Points2D()) {
        if (!point2D.HasPoint3D()) continue;
        num_observations += 1;
        point3D_num_observations[point2D.Point3DId()] += 1;
        obs.push_back({(point3D_t)point2D.Point3DId(), image_id, point2D.XY()(0), point2D.XY()(1)});
      }
      prob_images.push_back(image_id);
      constant_pose[image_id] = cp;
      if (num_observations > 0) camera_ids.insert(image.CameraId());
    }
    std::vector<point3D_t> added(config_.VariablePoints().begin(), config_.VariablePoints().end());
    std::sort(added.begin(), added.end());
    std::vector<point3D_t> cpts(config_.ConstantPoints().begin(), config_.ConstantPoints().end());
    std::sort(cpts.begin(), cpts.end());
    added.insert(added.end(), cpts.begin(), cpts.end());
    for (const point3D_t point3D_id : added) {  // AddPointToProblem (.cc:422-472)
      auto& point3D = reconstruction->Point3D(point3D_id);
      if (point3D_num_observations[point3D_id] == point3D.Track().Length()) continue;
      for (const auto& track_el : point3D.Track().Elements()) {
        if (config_.HasImage(track_el.image_id)) continue;
        point3D_num_observations[point3D_id] += 1;
        auto& image = reconstruction->Image(track_el.image_id);
        if (camera_ids.count(image.CameraId()) == 0) {
          camera_ids.insert(image.CameraId());
          config_.SetConstantCamera(image.CameraId());
        }
        if (!constant_pose.count(track_el.image_id)) {
          prob_images.push_back(track_el.image_id);
          constant_pose[track_el.image_id] = true;  // BundleAdjustmentConstantPoseCostFunction
        }
        const auto& point2D = image.Point2D(track_el.point2D_idx);
        obs.push_back({point3D_id, (image_t)track_el.image_id, point2D.XY()(0), point2D.XY()(1)});
      }
    }
    if (obs.empty()) return false;  // problem_->NumResiduals() == 0 (.cc:266-268)

    std::map<image_t, int32_t> img_idx;
    for (size_t k = 0; k < prob_images.size(); ++k) img_idx[prob_images[k]] = (int32_t)k;
    std::set<camera_t> used_cams;
    for (const image_t i : prob_images) used_cams.insert(reconstruction->Image(i).CameraId());
    std::map<camera_t, int32_t> cam_idx;
    std::vector<camera_t> cams(used_cams.begin(), used_cams.end());
    for (size_t k = 0; k < cams.size(); ++k) cam_idx[cams[k]] = (int32_t)k;
    std::map<point3D_t, int32_t> pt_idx;
    std::vector<point3D_t> pts;
    for (const auto& el : point3D_num_observations) { pt_idx[el.first] = (int32_t)pts.size(); pts.push_back(el.first); }
    std::stable_sort(obs.begin(), obs.end(), [&](const Obs& a, const Obs& b) { return pt_idx[a.point3D_id] < pt_idx[b.point3D_id]; });

    const int n_img = (int)prob_images.size(), n_cam = (int)cams.size(), n_pts = (int)pts.size();
    int stride = 4;
    for (const camera_t c : cams) stride = std::max(stride, (int)reconstruction->Camera(c).NumParams());
    std::vector<double> qvec(4 * (size_t)n_img), tvec(3 * (size_t)n_img), cam_params((size_t)stride * n_cam, 0.0), xyz(3 * (size_t)n_pts);
    std::vector<int32_t> image_camera(n_img), camera_model(n_cam), obs_image(obs.size()), obs_point(obs.size());
    std::vector<uint8_t> const_pose(n_img), const_tvec(n_img, 0), const_camera(n_cam), const_point(n_pts);
    std::vector<double> obs_xy(2 * obs.size());
    for (int k = 0; k < n_img; ++k) {
      auto& image = reconstruction->Image(prob_images[k]);
      for (int j = 0; j < 4; ++j) qvec[4 * k + j] = image.Qvec().data()[j];
      for (int j = 0; j < 3; ++j) tvec[3 * k + j] = image.Tvec().data()[j];
      image_camera[k] = cam_idx[image.CameraId()];
      const_pose[k] = constant_pose[prob_images[k]] ? 1 : 0;
      if (!const_pose[k] && config_.HasConstantTvec(prob_images[k]))  // SubsetParameterization on tvec (.cc:409-415)
        for (const int idx : config_.ConstantTvec(prob_images[k])) const_tvec[k] |= (uint8_t)(1u << idx);
    }
    for (int k = 0; k < n_cam; ++k) {
      auto& camera = reconstruction->Camera(cams[k]);
      camera_model[k] = (int32_t)camera.ModelId();
      for (size_t j = 0; j < camera.NumParams(); ++j) cam_params[(size_t)stride * k + j] = camera.ParamsData()[j];
      // ParameterizeCameras (.cc:474-512); cameras that never entered camera_ids_ carry no residuals
      const_camera[k] = (config_.IsConstantCamera(cams[k]) || camera_ids.count(cams[k]) == 0) ? 1 : 0;
    }
    for (int k = 0; k < n_pts; ++k) {
      auto& point3D = reconstruction->Point3D(pts[k]);
      for (int j = 0; j < 3; ++j) xyz[3 * k + j] = point3D.XYZ().data()[j];
      // ParameterizePoints (.cc:514-526)
      const_point[k] = (point3D.Track().Length() > point3D_num_observations[pts[k]] || config_.HasConstantPoint(pts[k])) ? 1 : 0;
    }
    for (size_t o = 0; o < obs.size(); ++o) {
      obs_image[o] = img_idx[obs[o].image_id];
      obs_point[o] = pt_idx[obs[o].point3D_id];
      obs_xy[2 * o] = obs[o].x;
      obs_xy[2 * o + 1] = obs[o].y;
    }
    b2_ba_problem p = {};
    p.n_images = n_img; p.n_cameras = n_cam; p.n_points = n_pts; p.n_obs = (int64_t)obs.size();
    p.qvec = qvec.data(); p.tvec = tvec.data(); p.image_camera = image_camera.data();
    p.const_pose = const_pose.data(); p.const_tvec = const_tvec.data();
    p.camera_model = camera_model.data(); p.camera_params = cam_params.data(); p.const_camera = const_camera.data();
    p.xyz = xyz.data(); p.const_point = const_point.data();
    p.obs_image = obs_image.data(); p.obs_point = obs_point.data(); p.obs_xy = obs_xy.data();
    p.camera_params_stride = stride;

    b2_ba_options o;
    b2_ba_default_options(&o);
    o.max_num_iterations = options_.solver_options.max_num_iterations;
    o.function_tolerance = options_.solver_options.function_tolerance;
    o.gradient_tolerance = options_.solver_options.gradient_tolerance;
    o.parameter_tolerance = options_.solver_options.parameter_tolerance;
    o.max_linear_solver_iterations = options_.solver_options.max_linear_solver_iterations;
    o.refine_focal_length = options_.refine_focal_length ? 1 : 0;
    o.refine_principal_point = options_.refine_principal_point ? 1 : 0;
    o.refine_extra_params = options_.refine_extra_params ? 1 : 0;
    o.loss_function_type = (int32_t)options_.loss_function_type;
    o.loss_function_scale = options_.loss_function_scale;
    // the reference's "empirical choice" on config_.NumImages() (.cc:272-284)
    const size_t kMaxNumImagesDirectSparseSolver = 1000;
    o.linear_solver_type = config_.NumImages() <= kMaxNumImagesDirectSparseSolver ? 1 : 2;

    b2_ba* h = nullptr;
    b2_ba_summary s;
    if (b2_ba_create(options_.gpu_index, &h) != B2_OK || b2_ba_solve(h, &p, &o, &s) != B2_OK) {
      std::fprintf(stderr, "ERROR: bundle adjustment failed: %s\n", b2_last_error());
      if (h) b2_ba_destroy(h);
      return false;
    }
    b2_ba_destroy(h);
    // Ceres updates the parameter blocks in place (TearDown is empty, .cc:334-336)
    for (int k = 0; k < n_img; ++k) {
      auto& image = reconstruction->Image(prob_images[k]);
      for (int j = 0; j < 4; ++j) image.Qvec().data()[j] = qvec[4 * k + j];
      for (int j = 0; j < 3; ++j) image.Tvec().data()[j] = tvec[3 * k + j];
    }
    for (int k = 0; k < n_cam; ++k) {
      auto& camera = reconstruction->Camera(cams[k]);
      for (size_t j = 0; j < camera.NumParams(); ++j) camera.ParamsData()[j] = cam_params[(size_t)stride * k + j];
    }
    for (int k = 0; k < n_pts; ++k) {
      auto& point3D = reconstruction->Point3D(pts[k]);
      for (int j = 0; j < 3; ++j) point3D.XYZ().data()[j] = xyz[3 * k + j];
    }
    summary_.initial_cost = s.initial_cost; summary_.final_cost = s.final_cost;
    summary_.num_residuals_reduced = s.num_residuals_reduced;
    summary_.num_effective_parameters_reduced = s.num_effective_parameters_reduced;
    summary_.num_successful_steps = s.num_successful_steps; summary_.num_unsuccessful_steps = s.num_unsuccessful_steps;
    summary_.termination_type = s.termination_type;
    summary_.num_linear_solver_iterations = (int)s.num_linear_solver_iterations;
    summary_.iterative_schur = s.linear_solver_type_used == 2;
    summary_.total_time_in_seconds = s.solve_seconds;
    if (options_.print_summary) PrintSolverSummary(summary_);
    return true;
  }

  const SolverSummary& Summary() const { return summary_; }

  // bundle_adjustment.cc:1126-1161 (the lines that have a counterpart here)
  static void PrintSolverSummary(const SolverSummary& summary) {
    std::printf("    Residuals : %d\n   Parameters : %d\n   Iterations : %d\n", summary.num_residuals_reduced,
                summary.num_effective_parameters_reduced, summary.num_successful_steps + summary.num_unsuccessful_steps);
    std::printf("         Time : %g [s]\n Initial cost : %g [px]\n   Final cost : %g [px]\n", summary.total_time_in_seconds,
                std::sqrt(summary.initial_cost / summary.num_residuals_reduced),
                std::sqrt(summary.final_cost / summary.num_residuals_reduced));
    std::printf("  Termination : %s\n\n", summary.termination_type == 0 ? "Convergence" : summary.termination_type == 1 ? "No convergence" : "Failure");
  }

 private:
  const BundleAdjustmentOptions options_;
  BundleAdjustmentConfig config_;
  SolverSummary summary_;
  bool used_ = false;
};

}  // namespace dagsfm_b200
