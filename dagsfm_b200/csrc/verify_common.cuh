// Shared declarations of the two-view verifier (kernel arguments, estimator ids).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/dagsfm_b200.h"

namespace b2 {

enum { EST_E5 = 0, EST_F7 = 1, EST_H4 = 2, EST_T2 = 3 };
// Estimator::kMinNumSamples / LocalEstimator::kMinNumSamples of the four LORANSAC
// instantiations in two_view_geometry.cc:325-341,539-541.
__host__ __device__ inline int min_samples(int t) { return t == EST_E5 ? 5 : t == EST_F7 ? 7 : t == EST_H4 ? 4 : 1; }
__host__ __device__ inline int local_min_samples(int t) { return t == EST_E5 ? 5 : t == EST_F7 ? 8 : t == EST_H4 ? 4 : 1; }

// HBM layout of the verifier's image store:
//   cams[n_images], img_off[n_images+1] (prefix of keypoint counts),
//   xy / nxy [img_off[n_images]] double2: pixel and normalised (ImageToWorld) keypoints.
// What travels between the stage kernels of one pair (verify_kernel.cu): the pair's std::mt19937 (state vector, the FIFO
// of outputs drawn by a batch of trials and handed back when the loop ended earlier, its read / write positions) and the
// reports of the three LORANSACs.
struct VerifyRansacReport {
  int32_t success, num_inliers;
  long long num_trials;
  double model[9];
};
struct VerifyPairState {
  uint32_t mt[624];
  uint32_t ring[256];
  uint32_t mti, w, r, pad;
  VerifyRansacReport report[3];   // E, F, H
};

struct VerifyArgs {
  const b2_camera* cams;
  const int64_t* img_off;
  int32_t n_images;
  const double2* xy;
  const double2* nxy;
  int64_t n_pairs;
  const uint32_t* pairs;
  const int64_t* match_off;
  const uint32_t* matches;
  const uint32_t* seeds;
  b2_two_view_options opt;
  b2_two_view_result* results;
  uint32_t* inlier_out;
  uint8_t* scratch;        // per-warp scratch, scratch_stride bytes each
  size_t scratch_stride;
  int32_t m_cap;           // max matches of any pair in this call
  int32_t max_workers;     // warps the scratch was sized for
  unsigned long long* work_counter;   // [4], one per stage
  VerifyPairState* state;  // [n_pairs]
  uint8_t* masks;          // [3][mask_stride] inlier masks of the E, F, H reports, a pair's slice at match_off[p] - mask_base
  int64_t mask_stride, mask_base;
  int* err;
  unsigned long long* prof;  // optional [8] cycle counters: sample, solve, score, lo, gather, other (B2_VERIFY_PROFILE=1)
};

size_t verify_scratch_stride(int m_cap);
int verify_warps_per_block();
int verify_blocks_per_sm();
cudaError_t launch_normalize_points(const b2_camera* cams, const int64_t* img_off, int n_images, const double* xy,
                                    double* nxy, int64_t n_total, cudaStream_t s);
cudaError_t launch_verify_pairs(const VerifyArgs& a, int n_sm, cudaStream_t s);
cudaError_t launch_relative_pose(const b2_camera* cams, const int64_t* img_off, int n_images, const double* nxy,
                                 int64_t n_pairs, const uint32_t* pairs, const int64_t* match_off,
                                 const b2_two_view_result* results, const uint32_t* inliers, b2_relative_pose* poses,
                                 double* angles, int* err, int n_sm, cudaStream_t s);
cudaError_t launch_score_models(int type, int n, const double* p1, const double* p2, int n_models, const double* models,
                                double max_res, int* counts, double* sums, uint8_t* masks, cudaStream_t s);
cudaError_t launch_debug_sample_stream(uint32_t seed, int total, int k, int n_trials, uint32_t* idx, int* out,
                                       cudaStream_t s);
cudaError_t launch_debug_solve(int type, int n, const double* p1, const double* p2, double* G, uint32_t* inl,
                               double* models, int* n_models, cudaStream_t s);

}  // namespace b2
