/*
 * dagsfm_b200 -- C ABI of the B200-native hot path of AIBluefisher/DAGSfM.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and returns an
 * int status (B2_OK == 0).  Nothing throws across this boundary and no torch /
 * Eigen / STL type appears in a signature.  Each block cites the reference
 * interface (path:line under the reference tree) it replaces.
 *
 * There is NO CPU fallback behind these calls: if no CUDA device / sm_100a
 * kernel image is available they return B2_ERR_CUDA / B2_ERR_NO_DEVICE.
 *
 * Ownership and threading.  A handle (b2_matcher, b2_verifier, b2_ba, b2_retrieval) belongs to one GPU and owns its
 * device buffers and its CUDA stream; calls on one handle must not overlap in time, calls on DIFFERENT handles may come
 * from different host threads at the same time (handles share nothing but the device; the last-error text is per thread,
 * the launch counter atomic; tests/test_concurrency_gpu.py).  Every call returns after its device work has completed: a
 * host buffer or a caller-owned device buffer passed in may be reused as soon as the call returns, and device outputs are
 * ready for any stream.  Multi-GPU jobs use one process (or thread) and one set of handles per GPU.
 */
#ifndef DAGSFM_B200_H_
#define DAGSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
  B2_OK = 0,
  B2_ERR_INVALID = 1,   /* bad argument (the reference CHECK()-aborts here)     */
  B2_ERR_CUDA = 2,      /* a CUDA call failed; b2_last_error() has the text     */
  B2_ERR_NO_DEVICE = 3, /* no sm_100 device visible                            */
  B2_ERR_CAPACITY = 4,  /* caller's output buffer too small; nothing written    */
  B2_ERR_INTERNAL = 5   /* device-side consistency check tripped                */
};

/* Text of the last error raised on the calling thread ("" if none). */
const char* b2_last_error(void);
/* Library version, e.g. "dagsfm_b200 0.1 (sm_100a)". */
const char* b2_version(void);
/* Number of CUDA kernels this library has launched since load (all threads);
 * bench.py reports the delta over the timed region as `gpu_launches`. */
uint64_t b2_kernel_launch_count(void);
/* Roofline denominators MEASURED_PEAKS.json does not hold, measured on this box: FMA
 * throughput of the CUDA cores in FP64 / FP32, TFLOP/s (16 independent chains per thread, best of 4). */
int b2_measure_fp64_peak(int device, double* tflops);
int b2_measure_ffma_peak(int device, double* tflops);

/* ================================================================== MATCH ==
 * Replaces: MatchSiftFeaturesGPU (src/feature/sift.h:235-239,
 * src/feature/sift.cc:941-985) and the SiftMatchGPU object behind it
 * (lib/SiftGPU/SiftGPU.h:276-373: SetDescriptors / GetSiftMatch), with the
 * *semantics of the CPU path* MatchSiftFeaturesCPU (sift.cc:76-198,810-822),
 * which is what the reference's own test treats as the truth
 * (src/feature/sift_test.cc:448-578).
 *
 * Descriptors: row-major uint8, 128 bytes per feature (FeatureDescriptors,
 * src/feature/types.h:102-103).  Matches: {uint32 idx1, uint32 idx2}
 * (FeatureMatch, types.h:86-98), ascending idx1.
 */
typedef struct b2_matcher b2_matcher;

typedef struct b2_match_options {
  float max_ratio;        /* SiftMatchingOptions::max_ratio    (sift.h:131) default 0.8 */
  float max_distance;     /* SiftMatchingOptions::max_distance (sift.h:134) default 0.7 */
  int32_t cross_check;    /* SiftMatchingOptions::cross_check  (sift.h:137) default 1   */
  int32_t max_num_matches;/* SiftMatchingOptions::max_num_matches (sift.h:140) 32768: features of an image beyond it
                             take no part in any match (the upload clamp of SiftMatchCU.cpp:108) -- applied IN PLACE
                             to the image store / the two slots by the first call that needs it, on both seams */
} b2_match_options;

/* Fills the reference defaults (src/feature/sift.h:116-165). */
void b2_match_default_options(b2_match_options* opt);

/* CreateSiftGPUMatcher (sift.cc:877-939).  `device` = CUDA ordinal. */
int b2_match_create(int device, b2_matcher** out);
int b2_match_destroy(b2_matcher* m);

/* -- image store: the batched seam ----------------------------------------
 * Uploads the descriptors of `n_images` images into one HBM pool (each image
 * zero-padded to a multiple of 256 rows).  `desc[i]` -> n_desc[i] x 128 bytes
 * in HOST memory.  Replaces the per-pair SetDescriptors H2D copies
 * (SiftMatchCU.cpp:99-112) and the FeatureMatcherCache::GetDescriptors calls
 * of SiftGPUFeatureMatcher::Run (src/feature/matching.cc:376-427). */
int b2_match_set_images(b2_matcher* m, int32_t n_images,
                        const uint8_t* const* desc, const int32_t* n_desc);
/* Same, from one contiguous DEVICE buffer: image i occupies rows
 * [row_offset[i], row_offset[i]+n_desc[i]) of `desc_dev` (row = 128 bytes). */
int b2_match_set_images_device(b2_matcher* m, int32_t n_images,
                               const uint8_t* desc_dev,
                               const int64_t* row_offset, const int32_t* n_desc);

/* Matches `n_pairs` image pairs (indices into the image store), all on the
 * device, and copies the result to HOST memory:
 *   out_offsets[n_pairs+1]  prefix offsets into out_matches (in matches)
 *   out_matches[capacity]   {idx1, idx2}
 * Returns B2_ERR_CAPACITY (and the needed count in *n_total) if capacity is
 * too small.  Replaces the loop of SiftGPUFeatureMatcher::Run over
 * MatchSiftFeaturesGPU (matching.cc:392-424). */
int b2_match_pairs(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs /*[n][2]*/,
                   const b2_match_options* opt, int64_t* out_offsets,
                   uint32_t* out_matches /*[capacity][2]*/, int64_t capacity,
                   int64_t* n_total);

/* Device-resident variant (inputs already in HBM, results stay in HBM): only
 * the total match count and the per-pair counts' checksum come back.  Used by
 * the bench's `value` leg and by callers that chain into b2_verify_*.
 * `pairs_dev` is a DEVICE pointer to n_pairs x 2 uint32.  `out_offsets_dev`
 * (n_pairs+1 int64) and `out_matches_dev` (capacity x 2 uint32) are DEVICE
 * buffers owned by the caller. */
int b2_match_pairs_device(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs_dev,
                          const b2_match_options* opt, int64_t* out_offsets_dev,
                          uint32_t* out_matches_dev, int64_t capacity,
                          int64_t* n_total);

/* -- two-slot seam: SiftMatchGPU::SetDescriptors / GetSiftMatch ------------
 * (lib/SiftGPU/SiftGPU.h:312-326).  slot in {0,1}; `desc == NULL` keeps the
 * previous upload (sift.h:232-234).  b2_match_run returns the number of
 * matches in *n_out: at most one per feature of slot 0 (without cross-check several of them may
 * share a feature of slot 1), truncated to max_num_matches -- the reference returns -1 on a
 * device error (SiftMatchCU.cpp:193-196); here that is a non-zero status. */
int b2_match_set_descriptors(b2_matcher* m, int slot, int32_t n, const uint8_t* desc);
int b2_match_run(b2_matcher* m, const b2_match_options* opt,
                 uint32_t* out_matches /*[max_num_matches][2]*/, int32_t* n_out);

/* Test / profiling hooks (no reference counterpart). */
/* Seconds of device time (CUDA events on the matcher's stream) spent in the
 * tensor-core kernel and in all kernels during the last b2_match_pairs* call,
 * and the number of tensor-core launches. */
int b2_match_last_timing(b2_matcher* m, double* tc_kernel_s, double* all_kernels_s,
                         int64_t* tc_launches, int64_t* fixup_candidates);

/* -- guided matching ----------------------------------------------------------
 * Replaces: MatchGuidedSiftFeaturesGPU / MatchGuidedSiftFeaturesCPU (src/feature/sift.h:240-246,
 * sift.cc:824-875, :987-1066) as called by SiftGPUFeatureMatcher / GuidedSiftGPUFeatureMatcher
 * for a verified pair (src/feature/matching.cc:438-534): descriptor matching restricted
 * to keypoint pairs whose float32 residual under the pair's geometry is <= max_error^2
 * (Sampson error w.r.t. F for CALIBRATED / UNCALIBRATED, transfer error w.r.t. H for PLANAR /
 * PANORAMIC / PLANAR_OR_PANORAMIC); everything else as in b2_match_pairs.
 * A pair whose `config` has no guided filter yields no matches (the reference returns without
 * touching two_view_geometry->inlier_matches, sift.cc:866-868: the caller keeps what it had). */
typedef struct b2_guided_geometry {
  int32_t config;   /* TwoViewGeometry::ConfigurationType (two_view_geometry.h:48-58) */
  int32_t reserved;
  double F[9];      /* row-major, TwoViewGeometry::F */
  double H[9];      /* row-major, TwoViewGeometry::H */
} b2_guided_geometry;

/* Keypoint locations (FeatureKeypoint::x, ::y, src/feature/types.h:44-45) of the images of the
 * last b2_match_set_images* call: xy[i] -> n_keypoints[i] (x, y) float pairs, which must equal
 * that image's descriptor count (sift.cc:82-87). */
int b2_match_set_keypoints(b2_matcher* m, int32_t n_images, const float* const* xy,
                           const int32_t* n_keypoints);
/* geoms[p] belongs to pairs[p]; `max_error` = SiftMatchingOptions::max_error (pixels, default 4). */
int b2_match_guided_pairs(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs,
                          const b2_guided_geometry* geoms, double max_error, const b2_match_options* opt,
                          int64_t* out_offsets, uint32_t* out_matches, int64_t capacity, int64_t* n_total);

/* The guided stage chained onto the verifier on the device (GuidedSiftGPUFeatureMatcher::Run, matching.cc:493-530):
 * the geometry of pair p is (config, F, H) of results_dev[p] as b2_verify_pairs_device wrote it; a pair with fewer than
 * min_num_inliers inliers (:508-512) or a configuration without a guided filter gets an empty slice -- the reference
 * passes such pairs through with the verifier's inlier list.  pairs / results / offsets / matches in DEVICE memory. */
struct b2_two_view_result;
int b2_match_guided_pairs_device(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs_dev,
                                 const struct b2_two_view_result* results_dev, int32_t min_num_inliers, double max_error,
                                 const b2_match_options* opt, int64_t* out_offsets_dev, uint32_t* out_matches_dev,
                                 int64_t capacity, int64_t* n_total);

/* ================================================================= VERIFY ==
 * Replaces: TwoViewGeometry::Estimate (src/estimators/two_view_geometry.h:180-184,
 * .cc:113-126) as called by TwoViewGeometryVerifier::Run for every matched pair
 * (src/feature/matching.cc:571-608): LO-RANSAC (src/optim/loransac.h:92-233) over
 * E 5-point, F 7-point (LO 8-point) and H 4-point DLT with Sampson / transfer
 * residuals, the configuration decision of EstimateCalibrated / EstimateUncalibrated
 * (two_view_geometry.cc:292-489) and DetectWatermark (:491-555).
 * The relative pose / triangulation angle that DAGSfM's Estimate adds for pairs of two
 * prior-focal-length cameras (EstimateWithRelativePose, :232-290) is the separate
 * call b2_verify_relative_pose on the results of b2_verify_pairs.
 *
 * The reference's verifier threads never seed their PRNG (src/util/random.cc:46-49), so
 * it is itself run-to-run random; here every pair carries an explicit seed that
 * initialises a std::mt19937-identical stream consumed E -> F -> H -> watermark.
 */
typedef struct b2_verifier b2_verifier;

typedef struct b2_camera {          /* Camera (src/base/camera.h) reduced to what the path reads */
  int32_t model;                    /* camera_models.h:117-129: 0 SIMPLE_PINHOLE f,cx,cy | 1 PINHOLE fx,fy,cx,cy |
                                       2 SIMPLE_RADIAL f,cx,cy,k | 3 RADIAL f,cx,cy,k1,k2 | 4 OPENCV fx,fy,cx,cy,k1,k2,p1,p2 |
                                       5 OPENCV_FISHEYE fx,fy,cx,cy,k1..k4 | 6 FULL_OPENCV ..,k1,k2,p1,p2,k3..k6 |
                                       7 FOV fx,fy,cx,cy,omega | 8 SIMPLE_RADIAL_FISHEYE f,cx,cy,k |
                                       9 RADIAL_FISHEYE f,cx,cy,k1,k2 | 10 THIN_PRISM_FISHEYE ..,k1,k2,p1,p2,k3,k4,sx1,sy1 */
  int32_t width, height;
  int32_t has_prior_focal_length;   /* Camera::HasPriorFocalLength() */
  double params[12];
} b2_camera;

typedef struct b2_two_view_options { /* TwoViewGeometry::Options + RANSACOptions */
  int32_t min_num_inliers;           /* 15   (two_view_geometry.h:82, sift.h:158) */
  int32_t detect_watermark;          /* 1    (two_view_geometry.h:98)             */
  double min_E_F_inlier_ratio;       /* 0.95 */
  double max_H_inlier_ratio;         /* 0.8  */
  double watermark_min_inlier_ratio; /* 0.7  */
  double watermark_border_size;      /* 0.1  */
  double max_error;                  /* 4.0 px  (sift.h:143)  */
  double min_inlier_ratio;           /* 0.25    (sift.h:153)  */
  double confidence;                 /* 0.999   (sift.h:146)  */
  int64_t min_num_trials;            /* 30      (sift.h:150)  */
  int64_t max_num_trials;            /* 10000   (sift.h:151)  */
} b2_two_view_options;

typedef struct b2_two_view_result {  /* public fields of TwoViewGeometry (two_view_geometry.h:278-301) */
  int32_t config;                    /* ConfigurationType: 1 DEGENERATE 2 CALIBRATED 3 UNCALIBRATED
                                        6 PLANAR_OR_PANORAMIC 7 WATERMARK
                                        8 MULTIPLE (b2_verify_pairs_multiple only) */
  int32_t n_inliers;                 /* inlier_matches.size() */
  int32_t E_num_inliers, F_num_inliers, H_num_inliers;
  int32_t E_num_trials, F_num_trials, H_num_trials;
  double E[9], F[9], H[9];           /* row-major */
} b2_two_view_result;

void b2_two_view_default_options(b2_two_view_options* opt);
int b2_verify_create(int device, b2_verifier** out);
int b2_verify_destroy(b2_verifier* v);
/* Cameras + keypoint locations (x,y as double, i.e. FeatureKeypointsToPointsVector,
 * src/feature/utils.cc:38-45) of all images, HOST pointers; copied to HBM, where the
 * normalised coordinates Camera::ImageToWorld (incl. IterativeUndistortion,
 * src/base/camera_models.h:547-590) are computed once per keypoint. */
int b2_verify_set_images(b2_verifier* v, int32_t n_images, const b2_camera* cams,
                         const double* const* xy, const int32_t* n_pts);
/* Verifies n_pairs pairs.  HOST buffers: pairs [n][2]; match_offsets [n+1] and matches
 * [total][2] exactly as b2_match_pairs produces them; seeds [n]; results [n];
 * inlier_matches [total][2]: the inliers of pair p are written at match_offsets[p]
 * (results[p].n_inliers of them, in match order). */
int b2_verify_pairs(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs,
                    const int64_t* match_offsets, const uint32_t* matches,
                    const b2_two_view_options* opt, const uint32_t* seeds,
                    b2_two_view_result* results, uint32_t* inlier_matches);
/* TwoViewGeometry::EstimateMultiple (two_view_geometry.cc:128-167; TwoViewGeometryVerifier::Run with
 * options.multiple_models, matching.cc:595-598): rounds of b2_verify_pairs on the matches that remain
 * after removing the inliers of the previous round, until a round is DEGENERATE.  Exactly one
 * geometry -> that result; several -> config 8 (MULTIPLE), inlier lists concatenated in round order,
 * E/F/H zero.  `multiple_ignore_watermark` = TwoViewGeometry::Options::multiple_ignore_watermark
 * (default true).  Round r of pair p is seeded with seeds[p] + r * 0x9E3779B9 (mod 2^32). */
int b2_verify_pairs_multiple(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs,
                             const int64_t* match_offsets, const uint32_t* matches,
                             const b2_two_view_options* opt, int32_t multiple_ignore_watermark,
                             const uint32_t* seeds, b2_two_view_result* results, uint32_t* inlier_matches);
/* Same with every buffer in DEVICE memory (chains onto b2_match_pairs_device). */
int b2_verify_pairs_device(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs_dev,
                           const int64_t* match_offsets_dev, const uint32_t* matches_dev,
                           const b2_two_view_options* opt, const uint32_t* seeds_dev,
                           b2_two_view_result* results_dev, uint32_t* inlier_matches_dev);
/* TwoViewGeometry::EstimateWithRelativePose after EstimateCalibrated (two_view_geometry.cc:239-289): for every
 * pair whose two cameras have a prior focal length (the dispatch of TwoViewGeometry::Estimate, :113-126) and whose
 * config is CALIBRATED / UNCALIBRATED (pose from E, base/essential_matrix.cc:41-88) or PLANAR_OR_PANORAMIC / WATERMARK
 * (pose from H, base/homography_matrix.cc:65-197; also PLANAR / PANORAMIC, which TwoViewGeometry::EstimateRelativePose,
 * :169-230, accepts when it is re-run on a stored geometry): the (R, t) candidate with the most inliers in front of both
 * cameras (CheckCheirality, base/pose.cc:225-248; ties keep the later candidate), qvec = RotationMatrixToQuaternion(R),
 * tri_angle = median triangulation angle of those points (base/triangulation.cc:183-215), and PLANAR_OR_PANORAMIC
 * resolved to PANORAMIC (|t| == 0, tri_angle 0) or PLANAR.  Every other pair (a camera without prior focal length,
 * DEGENERATE, MULTIPLE) gets qvec 0, tvec 0, tri_angle 0 and its config unchanged -- the values TwoViewGeometry's
 * constructor sets (two_view_geometry.h:159-166).  HOST buffers; results / inlier_matches exactly as b2_verify_pairs
 * wrote them. */
typedef struct b2_relative_pose {
  double qvec[4];                    /* w, x, y, z */
  double tvec[3];
  double tri_angle;                  /* radians */
  int32_t config;                    /* as b2_two_view_result::config, with 6 resolved to 4 PLANAR / 5 PANORAMIC */
  int32_t n_points3D;                /* inliers that passed the cheirality test of the chosen candidate */
} b2_relative_pose;
int b2_verify_relative_pose(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs, const int64_t* match_offsets,
                            const b2_two_view_result* results, const uint32_t* inlier_matches, b2_relative_pose* poses);
/* Same with every buffer in DEVICE memory (chains onto b2_verify_pairs_device). */
int b2_verify_relative_pose_device(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs_dev,
                                   const int64_t* match_offsets_dev, const b2_two_view_result* results_dev,
                                   const uint32_t* inlier_matches_dev, b2_relative_pose* poses_dev);
/* Kernel-level seam == Estimator::Residuals + InlierSupportMeasurer::Evaluate
 * (src/optim/support_measurement.cc:36-48) for n_models models over n points (HOST buffers).
 * type: 0/1 Sampson (E/F), 2 homography transfer.  counts[n_models], sums[n_models]
 * (residual sum in index order), masks[n_models][n] bytes. */
int b2_score_models(b2_verifier* v, int32_t type, int32_t n, const double* xy1, const double* xy2,
                    int32_t n_models, const double* models, double max_residual,
                    int32_t* counts, double* sums, uint8_t* masks);
/* Test hook: the sampler's index stream (RandomSampler over std::mt19937(seed)), n_trials x k. */
int b2_verify_debug_sample_stream(b2_verifier* v, uint32_t seed, int32_t total, int32_t k,
                                  int32_t n_trials, int32_t* out);
/* Test hook: the minimal / local solvers on one point set. type 0 E5, 1 F7, 2 H4, 3 F8(LO).
 * models_out [10][9]; returns the model count in *n_models. */
int b2_verify_debug_solve(b2_verifier* v, int32_t type, int32_t n, const double* xy1,
                          const double* xy2, double* models_out, int32_t* n_models);
/* Test hook: the normalised coordinates (Camera::ImageToWorld) the verifier holds for the keypoints of one image. */
int b2_verify_debug_normalized(b2_verifier* v, int32_t image, double* out_xy);
/* Device seconds (CUDA events) of the last b2_verify_pairs* call. */
int b2_verify_last_timing(b2_verifier* v, double* kernel_s);

/* ===================================================================== BA ==
 * Replaces: BundleAdjuster::Solve (src/optim/bundle_adjustment.h:165-197,
 * .cc:258-310) as configured by DistributedMapperController::GlobalBundleAdjustment /
 * AdjustGlobalBundle (src/controllers/distributed_mapper_controller.cpp:522-542,836-933)
 * and BundleAdjustmentController::Run (src/controllers/bundle_adjustment.cc:69-103):
 * reprojection residuals of BundleAdjustmentCostFunction / ...ConstantPoseCostFunction
 * (src/base/cost_functions.h:44-158), QuaternionParameterization on qvec,
 * SubsetParameterization on tvec / intrinsics, constant cameras / poses / points
 * (bundle_adjustment.cc:338-526), and the Levenberg-Marquardt + Schur-complement solve
 * that ceres::Solve performs: DENSE_SCHUR / SPARSE_SCHUR (exact step, dense Cholesky of the
 * reduced camera system) up to 1000 images and ITERATIVE_SCHUR + SCHUR_JACOBI (conjugate
 * gradients on the implicit Schur complement, block-Jacobi preconditioner) above, the rule of
 * bundle_adjustment.cc:274-284.
 * The problem is passed the way ParallelBundleAdjuster::SetUp packs it for PBA
 * (bundle_adjustment.cc:654-772): flat arrays, observations sorted by point.
 * All parameters are updated IN PLACE (as Ceres updates Image::Qvec/Tvec,
 * Camera::Params, Point3D::XYZ).
 */
typedef struct b2_ba b2_ba;

typedef struct b2_ba_problem {
  int32_t n_images, n_cameras, n_points;
  int64_t n_obs;
  double* qvec;              /* [n_images][4] w,x,y,z   (normalised on entry, .cc:345) */
  double* tvec;              /* [n_images][3] */
  const int32_t* image_camera; /* [n_images] -> camera (intrinsics may be shared, .cc:349) */
  const uint8_t* const_pose;   /* [n_images] 1 = BundleAdjustmentConfig::SetConstantPose     */
  const uint8_t* const_tvec;   /* [n_images] bitmask of SetConstantTvec components           */
  const int32_t* camera_model; /* [n_cameras] model ids of camera_models.h:117-129 (0 SIMPLE_PINHOLE ... 10
                                  THIN_PRISM_FISHEYE, layouts as in b2_camera); 0 / 1 / 2 run the fast path */
  double* camera_params;       /* [n_cameras][camera_params_stride] */
  const uint8_t* const_camera; /* [n_cameras] 1 = BundleAdjustmentConfig::SetConstantCamera  */
  double* xyz;                 /* [n_points][3] */
  const uint8_t* const_point;  /* [n_points] 1 = AddConstantPoint */
  const int32_t* obs_image;    /* [n_obs] */
  const int32_t* obs_point;    /* [n_obs] non-decreasing (sorted by point => CSR) */
  const double* obs_xy;        /* [n_obs][2] */
  int32_t camera_params_stride;/* doubles per camera in camera_params; 0 = 4 (enough for models 0-2), up to 12 */
  int32_t reserved;
} b2_ba_problem;

typedef struct b2_ba_options {   /* BundleAdjustmentOptions (bundle_adjustment.h:48-103) */
  int32_t max_num_iterations;    /* solver_options.max_num_iterations (final BA: 50)      */
  int32_t refine_focal_length;   /* 1 */
  int32_t refine_principal_point;/* 0 */
  int32_t refine_extra_params;   /* 1 */
  double function_tolerance;     /* final BA: 0   */
  double gradient_tolerance;     /* final BA: 1.0 */
  double parameter_tolerance;    /* final BA: 0   */
  int32_t loss_function_type;    /* BundleAdjustmentOptions::LossFunctionType (bundle_adjustment.h:49-51):
                                    0 TRIVIAL (final / global BA), 1 SOFT_L1 (the mapper's local BA,
                                    incremental_mapper_controller.cc:252-253), 2 CAUCHY */
  int32_t linear_solver_type;    /* 0 = the reference's rule on n_images (bundle_adjustment.cc:274-284:
                                    <= 1000 exact Schur step, else ITERATIVE_SCHUR + SCHUR_JACOBI),
                                    1 = exact (DENSE_SCHUR / SPARSE_SCHUR), 2 = ITERATIVE_SCHUR */
  double loss_function_scale;    /* 1.0 */
  int32_t max_linear_solver_iterations; /* CG iterations per LM step (ITERATIVE_SCHUR); final BA: 100
                                    (distributed_mapper_controller.cpp:529), BundleAdjustmentOptions: 200 */
  int32_t reserved;
} b2_ba_options;

typedef struct b2_ba_summary {   /* the fields of ceres::Solver::Summary the reference reads */
  double initial_cost, final_cost;      /* 1/2 sum r^2 */
  int32_t num_successful_steps, num_unsuccessful_steps;
  int32_t termination_type;             /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
  int32_t num_residuals_reduced, num_effective_parameters_reduced;
  int32_t num_iterations;               /* successful + unsuccessful */
  double solve_seconds;                 /* device time of the LM loop (CUDA events) */
  double schur_kernel_seconds;          /* ... of which inside the Jacobian+Schur kernels */
  int64_t schur_kernel_launches;
  int64_t num_linear_solver_iterations; /* CG iterations over all LM steps (0 on the exact path) */
  int32_t linear_solver_type_used;      /* 1 exact, 2 ITERATIVE_SCHUR */
  int32_t exact_path_used;              /* 0 n/a (iterative), 1 staged blocks + dense S + library Cholesky (fallback),
                                           2 fused kernels + packed tiles + the library's own tiled Cholesky */
  double linear_solve_seconds;          /* device time of factorisation + triangular solves (fused exact path) */
  double reduced_system_bytes;          /* size of the buffer (S | rhs | g_c | diag) that a multi-GPU run all-reduces */
} b2_ba_summary;

/* Collective hook for multi-GPU runs (points sharded across ranks, cameras replicated):
 * called with a DEVICE buffer of n doubles that must be reduced in place across all ranks
 * (op 0 = SUM, 1 = MAX) before returning.  NULL (default) = single GPU. */
typedef void (*b2_allreduce_fn)(void* dev_buf, int64_t n_doubles, int32_t op, void* user);

void b2_ba_default_options(b2_ba_options* opt);   /* GlobalBundleAdjustment() values */
int b2_ba_create(int device, b2_ba** out);
int b2_ba_destroy(b2_ba* h);
int b2_ba_set_allreduce(b2_ba* h, b2_allreduce_fn fn, void* user);
/* The library's own NCCL communicator for multi-GPU bundle adjustment (one process per GPU, points sharded across the
 * ranks, cameras replicated -- the layout of DistributedMapperController's final BA over clusters,
 * src/controllers/distributed_mapper_controller.cpp:836-933): rank 0 obtains an id with b2_nccl_unique_id and hands its
 * 128 bytes to the other ranks by any means (MPI, a file, torch.distributed); every rank then calls b2_ba_init_nccl.
 * From then on b2_ba_solve reduces the camera normal equations with ncclAllReduce on its own stream -- no host
 * synchronisation, no callback.  libnccl.so.2 is bound at run time (already loaded, on the loader path, or
 * B2_NCCL_LIBRARY). */
int b2_nccl_unique_id(uint8_t* out128);
int b2_ba_init_nccl(b2_ba* h, int32_t n_ranks, int32_t rank, const uint8_t* id128);
/* Solves the problem (HOST arrays in, parameters updated in place).  On a multi-GPU run
 * every rank passes ALL cameras/images and ITS shard of points + observations. */
int b2_ba_solve(b2_ba* h, const b2_ba_problem* problem, const b2_ba_options* opt,
                b2_ba_summary* summary);
/* Test seam of the linear solver of the exact Schur step (the library's tiled Cholesky, dagsfm_b200/csrc/ba_chol.cu):
 * solves A x = b for a dense symmetric positive definite A [D*D] (row-major, upper triangle read; zero 64 x 64 blocks
 * are skipped as in the bundle adjuster) on the device.  info != 0: a pivot was not positive. */
int b2_ba_debug_cholesky_solve(b2_ba* h, int64_t D, const double* A, const double* b, double* x, int32_t* info,
                               int32_t* n_tiles);

/* Result metrics the reference reports after a bundle adjustment:
 * Reconstruction::ComputeMeanReprojectionError (src/base/reconstruction.cc:814-858) with
 * CalculateSquaredReprojectionError (src/base/projection.cc:119-136) over the problem's tracks (its CSR rows):
 * per observation |WorldToImage(R(q) X + t) - xy| unless the point is not in front of the camera (skipped);
 * point_errors[p] (may be NULL) = the track's error sum / track length, the value the reference stores with
 * Point3D::SetError; *mean_reprojection_error = total error sum / total track length.  HOST arrays, nothing is modified. */
int b2_ba_reprojection_errors(b2_ba* h, const b2_ba_problem* problem, double* point_errors,
                              double* mean_reprojection_error);

/* ======================================================================================================================
 * Image retrieval: candidate pairs from a vocabulary tree.
 * Replaces the retrieval loop of VocabSimilarityGraph::Run (src/graph/similarity_graph.cpp:101-200) over
 * retrieval::VisualIndex<uint8_t, 128, 64> (src/retrieval/visual_index.h): Add for every image, Prepare, Query of every
 * image with QueryOptions{max_num_images, num_neighbors}; spatial re-ranking (num_images_after_verification, off in
 * the reference's defaults) is not part of this seam.  The visual words are searched EXACTLY: the answer of the reference's
 * vendored FLANN in exact mode (flann::LinearIndex), which its autotuned approximate index approximates.  One handle per GPU. */
typedef struct b2_retrieval b2_retrieval;
int b2_retrieval_create(int device, b2_retrieval** out);
int b2_retrieval_destroy(b2_retrieval* r);
/* The vocabulary as VisualIndex::Read delivers it: words [n_words * 128] (uint8 centroids), the Hamming-embedding
 * projection [64 * 128] (row-major float, InvertedIndex::proj_matrix_), per-word thresholds [n_words * 64]
 * (InvertedFile::thresholds_) and has_embedding [n_words] (InvertedFile::status_ & HAS_EMBEDDING).  HOST buffers. */
int b2_retrieval_set_vocabulary(b2_retrieval* r, int32_t n_words, const uint8_t* words, const float* proj,
                                const float* thresholds, const uint8_t* has_embedding);
/* VisualIndex::Add (IndexOptions::num_neighbors = 1) for images 0 .. n_images-1 + Prepare().  descriptors: all images
 * concatenated [desc_offsets[n_images] * 128]; desc_offsets [n_images + 1] (HOST).  num_neighbors_query (1..8) nearest
 * words are kept per descriptor for b2_retrieval_query_all (QueryOptions::num_neighbors, 5 in the reference).  The
 * _device variant reads descriptors already resident in HBM (they must stay valid until the next index call). */
int b2_retrieval_index_images(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors, const int64_t* desc_offsets,
                              int32_t num_neighbors_query);
int b2_retrieval_index_images_device(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors_dev,
                                     const int64_t* desc_offsets_host, int32_t num_neighbors_query);
/* VisualIndex::Query of every indexed image: for query image q, out_ids / out_scores [q * max_num_images + k], k <
 * out_counts[q], sorted by descending score (equal scores: lower image id first); the image itself is among its results,
 * as in the reference (the caller keeps pairs with image_id < other, similarity_graph.cpp:186-191).  HOST buffers. */
int b2_retrieval_query_all(b2_retrieval* r, int32_t max_num_images, int32_t* out_ids, float* out_scores, int32_t* out_counts);
/* The same for the query images [q0, q1) only; outputs are indexed from q0 (out_ids [(q - q0) * max_num_images + k]). */
int b2_retrieval_query_range(b2_retrieval* r, int32_t q0, int32_t q1, int32_t max_num_images, int32_t* out_ids, float* out_scores,
                             int32_t* out_counts);
/* Multi-GPU (one handle per GPU, vocabulary set on each): a rank searches the words of ITS share of the descriptors
 * (b2_retrieval_word_search_device: n_desc rows at descriptors_dev -> out_word_ids_dev [n_desc * num_neighbors], device
 * buffers), the ranks all-gather the word ids (the stage's one collective; NCCL, by the caller), every rank builds the
 * same index from them (b2_retrieval_index_images_words_device: all descriptors resident, word_ids_dev [n_total *
 * num_neighbors_query]) and queries its own range of images (b2_retrieval_query_range). */
int b2_retrieval_word_search_device(b2_retrieval* r, const uint8_t* descriptors_dev, int64_t n_desc, int32_t num_neighbors,
                                    int32_t* out_word_ids_dev);
int b2_retrieval_index_images_words_device(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors_dev,
                                           const int64_t* desc_offsets_host, int32_t num_neighbors_query, const int32_t* word_ids_dev);
/* Test hooks: nearest words of the indexed descriptors [n_desc * num_neighbors_query]; the inverted index (any pointer may
 * be NULL): word_start [n_words + 1], per entry image / feature / 64-bit signature, idf [n_words], norm [n_images]. */
int b2_retrieval_debug_word_ids(b2_retrieval* r, int32_t* out);
/* the same list recomputed by the independent SIMT (dp4a) statement of the word search */
int b2_retrieval_debug_word_ids_simt(b2_retrieval* r, int32_t* out);
int b2_retrieval_debug_index(b2_retrieval* r, uint32_t* word_start, int32_t* entry_image, int32_t* entry_feature,
                             uint64_t* entry_bits, float* idf, float* norm);
int b2_retrieval_last_timing(b2_retrieval* r, double* word_search_s, double* index_build_s, double* query_s);

#ifdef __cplusplus
}
#endif
#endif /* DAGSFM_B200_H_ */
