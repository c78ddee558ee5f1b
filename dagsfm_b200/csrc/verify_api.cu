// C ABI of the two-view verifier (include/dagsfm_b200.h, VERIFY section): HBM image
// store (cameras, keypoints, normalised keypoints), per-call scratch sizing, launches.
// Replaces the TwoViewGeometryVerifier threads of SiftFeatureMatcher
// (reference src/feature/matching.cc:571-608,647-673).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dagsfm_b200.h"
#include "common_host.h"
#include "verify_common.cuh"
#include "verify_multiple.h"

using namespace b2;

struct b2_verifier {
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  // image store
  int32_t n_images = 0;
  b2_camera* d_cams = nullptr;
  int64_t* d_img_off = nullptr;
  double* d_xy = nullptr;
  double* d_nxy = nullptr;
  int64_t n_pts_total = 0;
  // per-call
  uint8_t* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  uint8_t *d_state = nullptr, *d_masks = nullptr;   // inter-stage pair state and inlier masks (verify_kernel.cu)
  size_t state_bytes = 0, mask_bytes = 0;
  unsigned long long* d_counter = nullptr;
  int* d_err = nullptr;
  int* d_maxm = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_kernel_s = 0;
  // host-call staging
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  // relative-pose scratch (triangulation angles, one double per match)
  double* d_angles = nullptr;
  size_t angles_count = 0;
};

namespace {

__global__ void max_matches_kernel(const int64_t* off, int64_t n_pairs, int* out) {
  int m = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n_pairs; p += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (int)(off[p + 1] - off[p]));
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

int check_options(const b2_two_view_options* o) {
  if (!o) return set_error(B2_ERR_INVALID, "options == NULL");
  // RANSACOptions::Check (ransac.h:63-70) + TwoViewGeometry::Options::Check
  if (!(o->max_error > 0) || o->min_inlier_ratio < 0 || o->min_inlier_ratio > 1 || o->confidence < 0 ||
      o->confidence > 1 || o->min_num_trials > o->max_num_trials || o->min_num_trials < 0 || o->min_num_inliers < 0 ||
      o->min_E_F_inlier_ratio < 0 || o->min_E_F_inlier_ratio > 1 || o->max_H_inlier_ratio < 0 ||
      o->max_H_inlier_ratio > 1 || o->watermark_min_inlier_ratio < 0 || o->watermark_min_inlier_ratio > 1 ||
      o->watermark_border_size < 0 || o->watermark_border_size > 1)
    return set_error(B2_ERR_INVALID, "TwoViewGeometry::Options::Check failed");
  return B2_OK;
}

// Pairs per launch group: bounds the inter-stage state (3.8 KB per pair) whatever the size of the call.
constexpr int64_t kStageBatch = 65536;

int run_device(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs, const int64_t* off, const uint32_t* matches,
               const b2_two_view_options* opt, const uint32_t* seeds, b2_two_view_result* results, uint32_t* inl) {
  B2_TRY(check_options(opt));
  if (n_pairs < 0) return set_error(B2_ERR_INVALID, "n_pairs < 0");
  if (n_pairs == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(v->device));
  cudaStream_t s = v->stream;
  // largest match list of the call sizes the per-warp scratch; the match offsets at the batch boundaries size the masks
  B2_CUDA(cudaMemsetAsync(v->d_maxm, 0, sizeof(int), s));
  max_matches_kernel<<<64, 256, 0, s>>>(off, n_pairs, v->d_maxm);
  B2_CUDA(cudaGetLastError());
  int m_cap = 0;
  B2_CUDA(cudaMemcpyAsync(&m_cap, v->d_maxm, sizeof(int), cudaMemcpyDeviceToHost, s));
  const int64_t n_batches = (n_pairs + kStageBatch - 1) / kStageBatch;
  std::vector<int64_t> bound((size_t)n_batches + 1);
  for (int64_t b = 0; b <= n_batches; ++b)
    B2_CUDA(cudaMemcpyAsync(&bound[(size_t)b], off + std::min(b * kStageBatch, n_pairs), sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  m_cap = std::max(m_cap, 32);
  int64_t max_span = 0;
  for (int64_t b = 0; b < n_batches; ++b) {
    if (bound[(size_t)b + 1] < bound[(size_t)b]) return set_error(B2_ERR_INVALID, "match offsets are not ascending");
    max_span = std::max(max_span, bound[(size_t)b + 1] - bound[(size_t)b]);
  }
  const int wpb = verify_warps_per_block();
  int blocks = v->n_sm * verify_blocks_per_sm();  // resident blocks only, dynamic work counter
  blocks = (int)std::min<int64_t>(blocks, (std::min(n_pairs, kStageBatch) + wpb - 1) / wpb);
  const size_t stride = verify_scratch_stride(m_cap);
  // bound the scratch (large match lists -> fewer concurrent warps)
  const size_t budget = (size_t)8 << 30;
  while (blocks > 1 && stride * (size_t)blocks * wpb > budget) blocks /= 2;
  const size_t need = stride * (size_t)blocks * wpb;
  if (need > v->scratch_bytes) {
    if (v->d_scratch) cudaFree(v->d_scratch);
    v->d_scratch = nullptr;
    v->scratch_bytes = 0;
    B2_CUDA(cudaMalloc(&v->d_scratch, need));
    v->scratch_bytes = need;
  }
  const size_t state_need = (size_t)std::min(n_pairs, kStageBatch) * sizeof(VerifyPairState);
  if (state_need > v->state_bytes) {
    if (v->d_state) cudaFree(v->d_state);
    v->d_state = nullptr;
    v->state_bytes = 0;
    B2_CUDA(cudaMalloc(&v->d_state, state_need));
    v->state_bytes = state_need;
  }
  const size_t mask_need = 3 * (size_t)std::max<int64_t>(max_span, 1);
  if (mask_need > v->mask_bytes) {
    if (v->d_masks) cudaFree(v->d_masks);
    v->d_masks = nullptr;
    v->mask_bytes = 0;
    B2_CUDA(cudaMalloc(&v->d_masks, mask_need));
    v->mask_bytes = mask_need;
  }
  B2_CUDA(cudaMemsetAsync(v->d_err, 0, sizeof(int), s));
  VerifyArgs a;
  a.cams = v->d_cams;
  a.img_off = v->d_img_off;
  a.n_images = v->n_images;
  a.xy = (const double2*)v->d_xy;
  a.nxy = (const double2*)v->d_nxy;
  a.matches = matches;
  a.opt = *opt;
  a.inlier_out = inl;
  a.scratch = v->d_scratch;
  a.scratch_stride = stride;
  a.m_cap = m_cap;
  a.max_workers = blocks * wpb;
  a.work_counter = v->d_counter;
  a.state = (VerifyPairState*)v->d_state;
  a.masks = v->d_masks;
  a.mask_stride = std::max<int64_t>(max_span, 1);
  a.err = v->d_err;
  a.prof = nullptr;
  unsigned long long* d_prof = nullptr;
  if (getenv("B2_VERIFY_PROFILE")) {
    B2_CUDA(cudaMalloc(&d_prof, 8 * sizeof(unsigned long long)));
    B2_CUDA(cudaMemsetAsync(d_prof, 0, 8 * sizeof(unsigned long long), s));
    a.prof = d_prof;
  }
  B2_CUDA(cudaEventRecord(v->ev0, s));
  for (int64_t b = 0; b < n_batches; ++b) {
    const int64_t p0 = b * kStageBatch, p1 = std::min(n_pairs, p0 + kStageBatch);
    a.n_pairs = p1 - p0;
    a.pairs = pairs + 2 * p0;
    a.match_off = off + p0;
    a.seeds = seeds + p0;
    a.results = results + p0;
    a.mask_base = bound[(size_t)b];
    B2_CUDA(cudaMemsetAsync(v->d_counter, 0, 4 * sizeof(unsigned long long), s));
    B2_CUDA(launch_verify_pairs(a, v->n_sm, s));
    count_launches(4);
  }
  B2_CUDA(cudaEventRecord(v->ev1, s));
  count_launches(1);
  int err = 0;
  B2_CUDA(cudaMemcpyAsync(&err, v->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, v->ev0, v->ev1));
  v->last_kernel_s = ms * 1e-3;
  if (d_prof) {
    unsigned long long hp[8];
    cudaMemcpy(hp, d_prof, sizeof hp, cudaMemcpyDeviceToHost);
    cudaFree(d_prof);
    fprintf(stderr, "[b2 verify profile] warp-cycles: sample %.3g solve %.3g score %.3g lo %.3g | E %.3g F %.3g H %.3g T %.3g | kernel %.3f s\n",
            (double)hp[0], (double)hp[1], (double)hp[2], (double)hp[3], (double)hp[4], (double)hp[5], (double)hp[6], (double)hp[7], v->last_kernel_s);
  }
  if (err == 2) return set_error(B2_ERR_CUDA, "sampler FIFO exhausted inside one batch of trials (more than 256 raw draws)");
  if (err) return set_error(B2_ERR_INVALID, "pair references an image outside the store");
  return B2_OK;
}

int stage(b2_verifier* v, size_t bytes) {
  if (bytes <= v->stage_bytes) return B2_OK;
  if (v->d_stage) cudaFree(v->d_stage);
  v->d_stage = nullptr;
  v->stage_bytes = 0;
  B2_CUDA(cudaMalloc(&v->d_stage, bytes));
  v->stage_bytes = bytes;
  return B2_OK;
}

}  // namespace

extern "C" {

void b2_two_view_default_options(b2_two_view_options* o) {
  if (!o) return;
  o->min_num_inliers = 15;
  o->detect_watermark = 1;
  o->min_E_F_inlier_ratio = 0.95;
  o->max_H_inlier_ratio = 0.8;
  o->watermark_min_inlier_ratio = 0.7;
  o->watermark_border_size = 0.1;
  o->max_error = 4.0;
  o->min_inlier_ratio = 0.25;
  o->confidence = 0.999;
  o->min_num_trials = 30;
  o->max_num_trials = 10000;
}

int b2_verify_create(int device, b2_verifier** out) {
  if (!out) return set_error(B2_ERR_INVALID, "out == NULL");
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return set_error(B2_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
  }
  if (device < 0 || device >= n_dev) return set_error(B2_ERR_INVALID, "bad device ordinal");
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_error(B2_ERR_NO_DEVICE, "device is not sm_100");
  B2_CUDA(cudaSetDevice(device));
  b2_verifier* v = new b2_verifier();
  v->device = device;
  v->n_sm = prop.multiProcessorCount;
  const int rc = [&]() -> int {
    B2_CUDA(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    B2_CUDA(cudaMalloc(&v->d_counter, 4 * sizeof(unsigned long long)));
    B2_CUDA(cudaMalloc(&v->d_err, sizeof(int)));
    B2_CUDA(cudaMalloc(&v->d_maxm, sizeof(int)));
    B2_CUDA(cudaEventCreate(&v->ev0));
    B2_CUDA(cudaEventCreate(&v->ev1));
    return B2_OK;
  }();
  if (rc != B2_OK) {  // a half-built handle is released here, never handed out
    b2_verify_destroy(v);
    return rc;
  }
  *out = v;
  return B2_OK;
}

int b2_verify_destroy(b2_verifier* v) {
  if (!v) return B2_OK;
  cudaSetDevice(v->device);
  if (v->stream) cudaStreamSynchronize(v->stream);
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(v->d_cams); fr(v->d_img_off); fr(v->d_xy); fr(v->d_nxy); fr(v->d_scratch); fr(v->d_state); fr(v->d_masks); fr(v->d_counter);
  fr(v->d_err); fr(v->d_maxm); fr(v->d_stage); fr(v->d_angles);
  if (v->ev0) cudaEventDestroy(v->ev0);
  if (v->ev1) cudaEventDestroy(v->ev1);
  if (v->stream) cudaStreamDestroy(v->stream);
  delete v;
  return B2_OK;
}

int b2_verify_set_images(b2_verifier* v, int32_t n_images, const b2_camera* cams, const double* const* xy,
                         const int32_t* n_pts) {
  if (!v || n_images < 0 || (n_images > 0 && (!cams || !xy || !n_pts))) return set_error(B2_ERR_INVALID, "NULL argument");
  B2_CUDA(cudaSetDevice(v->device));
  for (int32_t i = 0; i < n_images; ++i) {
    if (n_pts[i] < 0) return set_error(B2_ERR_INVALID, "negative keypoint count");
    if (cams[i].model < 0 || cams[i].model > 10)
      return set_error(B2_ERR_INVALID, "unknown camera model id (0 SIMPLE_PINHOLE ... 10 THIN_PRISM_FISHEYE, camera_models.h:117-129)");
  }
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(v->d_cams); fr(v->d_img_off); fr(v->d_xy); fr(v->d_nxy);
  v->d_cams = nullptr; v->d_img_off = nullptr; v->d_xy = nullptr; v->d_nxy = nullptr;
  v->n_images = 0;  // the store stays empty (every pair is rejected) unless this call completes
  v->n_pts_total = 0;
  std::vector<int64_t> off(n_images + 1, 0);
  for (int32_t i = 0; i < n_images; ++i) off[i + 1] = off[i] + n_pts[i];
  for (int32_t i = 0; i < n_images; ++i)
    if (n_pts[i] > 0 && !xy[i]) return set_error(B2_ERR_INVALID, "NULL keypoint pointer");
  const size_t np = (size_t)std::max<int64_t>(off[n_images], 1);
  B2_CUDA(cudaMalloc(&v->d_cams, std::max<size_t>(1, n_images) * sizeof(b2_camera)));
  B2_CUDA(cudaMalloc(&v->d_img_off, (n_images + 1) * sizeof(int64_t)));
  B2_CUDA(cudaMalloc(&v->d_xy, np * 16));
  B2_CUDA(cudaMalloc(&v->d_nxy, np * 16));
  cudaStream_t s = v->stream;
  // Keypoints are normalised with the camera's own model.  The array the verification kernels keep afterwards only
  // serves ImageToWorldThreshold / CalibrationMatrix, which depend on the parameter LAYOUT alone: every model with
  // two focal lengths (fx, fy, cx, cy first) is stored there as PINHOLE (id 1), so those kernels need one test.
  struct Scratch {  // the true-model copy lives for this call only, whichever way it ends
    b2_camera* p = nullptr;
    ~Scratch() { if (p) cudaFree(p); }
  } true_cams;
  b2_camera*& d_true = true_cams.p;
  std::vector<b2_camera> view(cams, cams + n_images);
  bool any_general = false;
  for (auto& c : view)
    if (c.model > 2) {
      any_general = true;
      const bool two = c.model == 4 || c.model == 5 || c.model == 6 || c.model == 7 || c.model == 10;
      c.model = two ? 1 : 0;
    }
  if (n_images > 0) {
    B2_CUDA(cudaMemcpyAsync(v->d_cams, view.data(), n_images * sizeof(b2_camera), cudaMemcpyHostToDevice, s));
    if (any_general) {
      B2_CUDA(cudaMalloc(&d_true, n_images * sizeof(b2_camera)));
      B2_CUDA(cudaMemcpyAsync(d_true, cams, n_images * sizeof(b2_camera), cudaMemcpyHostToDevice, s));
    }
  }
  B2_CUDA(cudaMemcpyAsync(v->d_img_off, off.data(), (n_images + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  {  // keypoint arrays that are adjacent in the caller's memory travel as one copy (the device array is contiguous anyway)
    const double* run_src = nullptr;
    int64_t run_first = 0, run_pts = 0;
    for (int32_t i = 0; i <= n_images; ++i) {
      const bool last = i == n_images;
      if (!last && n_pts[i] == 0) continue;
      if (!last && run_pts && xy[i] == run_src + 2 * run_pts) {
        run_pts += n_pts[i];
        continue;
      }
      if (run_pts) B2_CUDA(cudaMemcpyAsync(v->d_xy + 2 * run_first, run_src, (size_t)run_pts * 16, cudaMemcpyHostToDevice, s));
      if (last) break;
      run_src = xy[i];
      run_first = off[i];
      run_pts = n_pts[i];
    }
  }
  B2_CUDA(launch_normalize_points(d_true ? d_true : v->d_cams, v->d_img_off, n_images, v->d_xy, v->d_nxy, off[n_images], s));
  count_launches(1);
  B2_CUDA(cudaStreamSynchronize(s));
  v->n_images = n_images;
  v->n_pts_total = off[n_images];
  return B2_OK;
}

int b2_verify_debug_normalized(b2_verifier* v, int32_t image, double* out_xy) {
  if (!v || !out_xy || image < 0 || image >= v->n_images) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(v->device));
  int64_t off[2];
  B2_CUDA(cudaMemcpy(off, v->d_img_off + image, sizeof off, cudaMemcpyDeviceToHost));
  if (off[1] > off[0]) B2_CUDA(cudaMemcpy(out_xy, v->d_nxy + 2 * off[0], (size_t)(off[1] - off[0]) * 16, cudaMemcpyDeviceToHost));
  return B2_OK;
}

int b2_verify_pairs_device(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs_dev, const int64_t* match_offsets_dev,
                           const uint32_t* matches_dev, const b2_two_view_options* opt, const uint32_t* seeds_dev,
                           b2_two_view_result* results_dev, uint32_t* inlier_matches_dev) {
  if (!v || (n_pairs > 0 && (!pairs_dev || !match_offsets_dev || !seeds_dev || !results_dev)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  return run_device(v, n_pairs, pairs_dev, match_offsets_dev, matches_dev, opt, seeds_dev, results_dev,
                    inlier_matches_dev);
}

int b2_verify_pairs(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs, const int64_t* match_offsets,
                    const uint32_t* matches, const b2_two_view_options* opt, const uint32_t* seeds,
                    b2_two_view_result* results, uint32_t* inlier_matches) {
  if (!v || n_pairs < 0 || (n_pairs > 0 && (!pairs || !match_offsets || !seeds || !results)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs == 0) return check_options(opt);
  B2_CUDA(cudaSetDevice(v->device));
  const int64_t total = match_offsets[n_pairs];
  if (total < 0 || (total > 0 && (!matches || !inlier_matches))) return set_error(B2_ERR_INVALID, "NULL match buffers");
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t b_pairs = al(n_pairs * 8), b_off = al((n_pairs + 1) * 8), b_m = al((size_t)std::max<int64_t>(total, 1) * 8),
               b_seed = al(n_pairs * 4), b_res = al(n_pairs * sizeof(b2_two_view_result));
  B2_TRY(stage(v, b_pairs + b_off + 2 * b_m + b_seed + b_res));
  uint8_t* p = (uint8_t*)v->d_stage;
  uint32_t* d_pairs = (uint32_t*)p; p += b_pairs;
  int64_t* d_off = (int64_t*)p; p += b_off;
  uint32_t* d_m = (uint32_t*)p; p += b_m;
  uint32_t* d_inl = (uint32_t*)p; p += b_m;
  uint32_t* d_seed = (uint32_t*)p; p += b_seed;
  b2_two_view_result* d_res = (b2_two_view_result*)p;
  cudaStream_t s = v->stream;
  B2_CUDA(cudaMemcpyAsync(d_pairs, pairs, n_pairs * 8, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_off, match_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
  if (total > 0) B2_CUDA(cudaMemcpyAsync(d_m, matches, total * 8, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_seed, seeds, n_pairs * 4, cudaMemcpyHostToDevice, s));
  B2_TRY(run_device(v, n_pairs, d_pairs, d_off, d_m, opt, d_seed, d_res, d_inl));
  B2_CUDA(cudaMemcpyAsync(results, d_res, n_pairs * sizeof(b2_two_view_result), cudaMemcpyDeviceToHost, s));
  if (total > 0) B2_CUDA(cudaMemcpyAsync(inlier_matches, d_inl, total * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  return B2_OK;
}


}  // extern "C"
namespace {
int pose_device(b2_verifier* v, int64_t n_pairs, int64_t total, const uint32_t* pairs, const int64_t* off,
                const b2_two_view_result* results, const uint32_t* inl, b2_relative_pose* poses) {
  cudaStream_t s = v->stream;
  const size_t need = (size_t)std::max<int64_t>(total, 1);
  if (need > v->angles_count) {
    if (v->d_angles) cudaFree(v->d_angles);
    v->d_angles = nullptr;
    v->angles_count = 0;
    B2_CUDA(cudaMalloc(&v->d_angles, need * sizeof(double)));
    v->angles_count = need;
  }
  B2_CUDA(cudaMemsetAsync(v->d_err, 0, sizeof(int), s));
  B2_CUDA(launch_relative_pose(v->d_cams, v->d_img_off, v->n_images, v->d_nxy, n_pairs, pairs, off, results, inl, poses,
                               v->d_angles, v->d_err, v->n_sm, s));
  count_launches(1);
  int err = 0;
  B2_CUDA(cudaMemcpyAsync(&err, v->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  if (err == 2)
    return set_error(B2_ERR_INVALID, "inlier list inconsistent with its pair (n_inliers > match count, or a keypoint index out of range)");
  if (err) return set_error(B2_ERR_INVALID, "pair references an image outside the store");
  return B2_OK;
}
}  // namespace
extern "C" {

int b2_verify_relative_pose_device(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs_dev, const int64_t* match_offsets_dev,
                                   const b2_two_view_result* results_dev, const uint32_t* inlier_matches_dev,
                                   b2_relative_pose* poses_dev) {
  if (!v || n_pairs < 0 || (n_pairs > 0 && (!pairs_dev || !match_offsets_dev || !results_dev || !poses_dev)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(v->device));
  int64_t total = 0;
  B2_CUDA(cudaMemcpyAsync(&total, match_offsets_dev + n_pairs, sizeof(int64_t), cudaMemcpyDeviceToHost, v->stream));
  B2_CUDA(cudaStreamSynchronize(v->stream));
  if (total < 0 || (total > 0 && !inlier_matches_dev)) return set_error(B2_ERR_INVALID, "bad match offsets / NULL inlier buffer");
  return pose_device(v, n_pairs, total, pairs_dev, match_offsets_dev, results_dev, inlier_matches_dev, poses_dev);
}

int b2_verify_relative_pose(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs, const int64_t* match_offsets,
                            const b2_two_view_result* results, const uint32_t* inlier_matches, b2_relative_pose* poses) {
  if (!v || n_pairs < 0 || (n_pairs > 0 && (!pairs || !match_offsets || !results || !poses)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(v->device));
  const int64_t total = match_offsets[n_pairs];
  if (total < 0 || (total > 0 && !inlier_matches)) return set_error(B2_ERR_INVALID, "bad match offsets / NULL inlier buffer");
  for (int64_t p = 0; p < n_pairs; ++p)
    if (results[p].n_inliers < 0 || results[p].n_inliers > match_offsets[p + 1] - match_offsets[p])
      return set_error(B2_ERR_INVALID, "n_inliers exceeds the pair's match count");
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t b_pairs = al(n_pairs * 8), b_off = al((n_pairs + 1) * 8), b_m = al((size_t)std::max<int64_t>(total, 1) * 8),
               b_res = al(n_pairs * sizeof(b2_two_view_result)), b_pose = al(n_pairs * sizeof(b2_relative_pose));
  B2_TRY(stage(v, b_pairs + b_off + b_m + b_res + b_pose));
  uint8_t* q = (uint8_t*)v->d_stage;
  uint32_t* d_pairs = (uint32_t*)q; q += b_pairs;
  int64_t* d_off = (int64_t*)q; q += b_off;
  uint32_t* d_inl = (uint32_t*)q; q += b_m;
  b2_two_view_result* d_res = (b2_two_view_result*)q; q += b_res;
  b2_relative_pose* d_pose = (b2_relative_pose*)q;
  cudaStream_t s = v->stream;
  B2_CUDA(cudaMemcpyAsync(d_pairs, pairs, n_pairs * 8, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_off, match_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
  if (total > 0) B2_CUDA(cudaMemcpyAsync(d_inl, inlier_matches, total * 8, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_res, results, n_pairs * sizeof(b2_two_view_result), cudaMemcpyHostToDevice, s));
  B2_TRY(pose_device(v, n_pairs, total, d_pairs, d_off, d_res, d_inl, d_pose));
  B2_CUDA(cudaMemcpyAsync(poses, d_pose, n_pairs * sizeof(b2_relative_pose), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  return B2_OK;
}

int b2_verify_pairs_multiple(b2_verifier* v, int64_t n_pairs, const uint32_t* pairs, const int64_t* match_offsets,
                             const uint32_t* matches, const b2_two_view_options* opt, int32_t multiple_ignore_watermark,
                             const uint32_t* seeds, b2_two_view_result* results, uint32_t* inlier_matches) {
  if (!v || n_pairs < 0 || (n_pairs > 0 && (!pairs || !match_offsets || !seeds || !results)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs == 0) return check_options(opt);
  const int64_t total = match_offsets[n_pairs];
  if (total < 0 || (total > 0 && (!matches || !inlier_matches))) return set_error(B2_ERR_INVALID, "NULL match buffers");
  // one round = one ordinary batched Estimate of the still-active pairs
  auto round = [&](int64_t na, const int64_t* ids, const int64_t* off, const uint32_t* m, const uint32_t* sd,
                   b2_two_view_result* res, uint32_t* inl) -> int {
    std::vector<uint32_t> pr((size_t)na * 2);
    for (int64_t k = 0; k < na; ++k) {
      pr[2 * k] = pairs[2 * ids[k]];
      pr[2 * k + 1] = pairs[2 * ids[k] + 1];
    }
    return b2_verify_pairs(v, na, pr.data(), off, m, opt, sd, res, inl);
  };
  return estimate_multiple(n_pairs, match_offsets, matches, seeds, multiple_ignore_watermark != 0, round, results,
                           inlier_matches);
}

int b2_score_models(b2_verifier* v, int32_t type, int32_t n, const double* xy1, const double* xy2, int32_t n_models,
                    const double* models, double max_residual, int32_t* counts, double* sums, uint8_t* masks) {
  if (!v || n < 0 || n_models < 0 || type < 0 || type > 3) return set_error(B2_ERR_INVALID, "bad argument");
  if (n_models == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(v->device));
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t bp = al((size_t)std::max(n, 1) * 16), bm = al((size_t)n_models * 72), bc = al((size_t)n_models * 4),
               bs = al((size_t)n_models * 8), bk = al((size_t)n_models * std::max(n, 1));
  B2_TRY(stage(v, 2 * bp + bm + bc + bs + bk));
  uint8_t* p = (uint8_t*)v->d_stage;
  double* d1 = (double*)p; p += bp;
  double* d2 = (double*)p; p += bp;
  double* dm = (double*)p; p += bm;
  int* dc = (int*)p; p += bc;
  double* ds = (double*)p; p += bs;
  uint8_t* dk = p;
  cudaStream_t s = v->stream;
  if (n > 0) {
    B2_CUDA(cudaMemcpyAsync(d1, xy1, (size_t)n * 16, cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaMemcpyAsync(d2, xy2, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  }
  B2_CUDA(cudaMemcpyAsync(dm, models, (size_t)n_models * 72, cudaMemcpyHostToDevice, s));
  B2_CUDA(launch_score_models(type, n, d1, d2, n_models, dm, max_residual, dc, ds, dk, s));
  count_launches(1);
  B2_CUDA(cudaMemcpyAsync(counts, dc, (size_t)n_models * 4, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(sums, ds, (size_t)n_models * 8, cudaMemcpyDeviceToHost, s));
  if (n > 0 && masks) B2_CUDA(cudaMemcpyAsync(masks, dk, (size_t)n_models * n, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  return B2_OK;
}

int b2_verify_debug_sample_stream(b2_verifier* v, uint32_t seed, int32_t total, int32_t k, int32_t n_trials,
                                  int32_t* out) {
  if (!v || !out || total < k || k < 1 || k > 8 || n_trials < 0) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(v->device));
  B2_TRY(stage(v, (size_t)total * 4 + (size_t)std::max(n_trials, 1) * k * 4 + 512));
  uint32_t* d_idx = (uint32_t*)v->d_stage;
  int* d_out = (int*)((uint8_t*)v->d_stage + ((size_t)total * 4 + 255) / 256 * 256);
  B2_CUDA(launch_debug_sample_stream(seed, total, k, n_trials, d_idx, d_out, v->stream));
  B2_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_trials * k * 4, cudaMemcpyDeviceToHost, v->stream));
  B2_CUDA(cudaStreamSynchronize(v->stream));
  return B2_OK;
}

int b2_verify_debug_solve(b2_verifier* v, int32_t type, int32_t n, const double* xy1, const double* xy2,
                          double* models_out, int32_t* n_models) {
  if (!v || !xy1 || !xy2 || !models_out || !n_models || type < 0 || type > 3 || n < 1)
    return set_error(B2_ERR_INVALID, "bad argument");
  if ((type != 3 && n < min_samples(type)) || (type == 3 && n < 8) || (type == 1 && n != 7))
    return set_error(B2_ERR_INVALID, "bad sample size for the solver");
  B2_CUDA(cudaSetDevice(v->device));
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  const size_t bp = al((size_t)n * 16), bg = al((size_t)n * 2 * 9 * 8), bi = al((size_t)n * 4);
  B2_TRY(stage(v, 2 * bp + bg + bi + al(90 * 8) + 256));
  uint8_t* p = (uint8_t*)v->d_stage;
  double* d1 = (double*)p; p += bp;
  double* d2 = (double*)p; p += bp;
  double* dg = (double*)p; p += bg;
  uint32_t* di = (uint32_t*)p; p += bi;
  double* dm = (double*)p; p += al(90 * 8);
  int* dn = (int*)p;
  cudaStream_t s = v->stream;
  B2_CUDA(cudaMemcpyAsync(d1, xy1, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d2, xy2, (size_t)n * 16, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemsetAsync(dn, 0, 4, s));
  B2_CUDA(launch_debug_solve(type, n, d1, d2, dg, di, dm, dn, s));
  int nm = 0;
  B2_CUDA(cudaMemcpyAsync(&nm, dn, 4, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(models_out, dm, 90 * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  *n_models = nm;
  return B2_OK;
}

int b2_verify_last_timing(b2_verifier* v, double* kernel_s) {
  if (!v) return set_error(B2_ERR_INVALID, "NULL verifier");
  if (kernel_s) *kernel_s = v->last_kernel_s;
  return B2_OK;
}

}  // extern "C"
