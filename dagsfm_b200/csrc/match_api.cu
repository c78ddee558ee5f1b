// C ABI of the descriptor matcher (include/dagsfm_b200.h, MATCH section) and the
// host-side logic above the kernels: the HBM image store, the integer threshold
// tables, the TMA tensor map, chunked scheduling of pair batches.
//
// Reference call sites this replaces: SiftGPUFeatureMatcher::Run
// (src/feature/matching.cc:376-427) -> MatchSiftFeaturesGPU (src/feature/sift.cc:941-985)
// -> SiftMatchGPU::SetDescriptors / GetSiftMatch (lib/SiftGPU/SiftMatchCU.cpp:99-199).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dagsfm_b200.h"
#include "common_host.h"
#include "match_common.cuh"
#include "match_guided.cuh"
#include "match_thresholds.h"

namespace b2 {

// kernels / launchers defined in match_tc_ts.cu and match_post.cu
cudaError_t launch_pair_items(const uint32_t* pairs, int64_t n_pairs, const int32_t* img_n,
                              int32_t n_images, uint32_t* n_items_of_pair, int* err, cudaStream_t s);
cudaError_t launch_scan_u32(const uint32_t* in, int64_t n, uint32_t* out, uint32_t* total,
                            cudaStream_t s);
cudaError_t launch_scan_counts(const uint32_t* counts, int64_t n, int64_t* offsets,
                               int64_t* carry_inout, bool write_last, cudaStream_t s);
cudaError_t launch_fill_items(const uint32_t* pairs, int64_t n_pairs, const int32_t* img_n, int32_t n_images,
                              const uint32_t* img_row, const uint32_t* item_start, MatchItem* items,
                              PairMeta* meta, uint32_t y_block_rows, cudaStream_t s);
cudaError_t launch_match_top2_ts(const CUtensorMap& tmap, const uint8_t* pool, const MatchItem* items,
                                 const uint32_t* n_items_ptr, int thr_dist, const int* ratio_lim, int* midx,
                                 uint4* cands, unsigned int* cand_count, unsigned int cand_capacity, int grid,
                                 cudaStream_t stream);
cudaError_t launch_fixup(const uint8_t* pool, const MatchItem* items, const uint4* cands,
                         const unsigned int* cand_count, unsigned int cand_capacity,
                         const int* ratio_lim, int* midx, int* err, int n_sm, cudaStream_t s);
cudaError_t launch_cross_count(const PairMeta* meta, int64_t n_pairs, const int* midx,
                               int cross_check, uint32_t* counts, cudaStream_t s);
cudaError_t launch_cross_write(const PairMeta* meta, int64_t n_pairs, const int* midx,
                               int cross_check, const int64_t* offsets, uint32_t* out_matches,
                               int64_t capacity, cudaStream_t s);
cudaError_t launch_guided_item_pairs(const PairMeta* meta, int64_t n_pairs, uint32_t* item_pair, cudaStream_t s);
cudaError_t launch_geoms_from_results(int64_t n_pairs, const void* results, int min_num_inliers, GuidedGeom* geoms, cudaStream_t s);
cudaError_t launch_guided_match(const uint8_t* pool, const float* kp_pool, const MatchItem* items,
                                const uint32_t* item_pair, const uint32_t* n_items_ptr, const PairMeta* meta,
                                const GuidedGeom* geoms, float max_residual, int thr_dist, const int* ratio_lim,
                                int* midx, int n_sm, cudaStream_t s);

static inline uint32_t pad_up(uint32_t n, uint32_t m) { return (n + m - 1) / m * m; }
// Rows reserved for an image of n descriptors: whole 256-row X supertiles (the zero padding rows also keep every
// 128-row Y block of the tensor-core kernel inside the image's own rows).
static inline uint32_t image_rows(uint32_t n) { return pad_up(n, kSuperRows); }

// ---------------------------------------------------------------- tensor map
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static int make_pool_tmap(CUtensorMap* tm, void* pool, uint64_t rows, uint32_t box_rows) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) return set_error(B2_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)kDescBytes, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)kDescBytes};
  const cuuint32_t box[2] = {(cuuint32_t)kDescBytes, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, pool, gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return set_error(B2_ERR_CUDA, buf);
  }
  return B2_OK;
}

// --------------------------------------------------------------- image store
struct ImageStore {
  uint8_t* pool = nullptr;
  float* kp_pool = nullptr;  // (x, y) per pool row, guided matching only
  uint64_t pool_rows = 0;
  int32_t n_images = 0;
  int32_t* d_img_n = nullptr;
  uint32_t* d_img_row = nullptr;
  std::vector<int32_t> h_img_n;
  std::vector<uint32_t> h_img_row;
  uint32_t max_n = 0;
  CUtensorMap tmap;    // box 128 rows x 128 B, SWIZZLE_128B

  void release() {
    if (pool) cudaFree(pool);
    if (kp_pool) cudaFree(kp_pool);
    kp_pool = nullptr;
    if (d_img_n) cudaFree(d_img_n);
    if (d_img_row) cudaFree(d_img_row);
    pool = nullptr;
    d_img_n = nullptr;
    d_img_row = nullptr;
    pool_rows = 0;
    n_images = 0;
    max_n = 0;
    h_img_n.clear();
    h_img_row.clear();
  }
  // Lays the images out (each padded to 256 rows), allocates a zeroed pool.
  int layout(int32_t n, const int32_t* n_desc, cudaStream_t s) {
    if (n < 0) return set_error(B2_ERR_INVALID, "n_images < 0");
    // Same image sizes as the previous upload (the per-step re-upload of a fixed job): keep the
    // pool, its padding rows are still zero and every descriptor row is about to be overwritten.
    if (pool && n == n_images && n > 0 && std::equal(n_desc, n_desc + n, h_img_n.begin())) return B2_OK;
    release();
    h_img_n.assign(n_desc, n_desc + n);
    h_img_row.resize(n);
    uint64_t rows = 0;
    for (int32_t i = 0; i < n; ++i) {
      if (n_desc[i] < 0) return set_error(B2_ERR_INVALID, "negative descriptor count");
      h_img_row[i] = (uint32_t)rows;
      rows += image_rows((uint32_t)n_desc[i]);
      max_n = std::max<uint32_t>(max_n, (uint32_t)n_desc[i]);
      if (rows > 0xFFFF0000ull) return set_error(B2_ERR_INVALID, "image store exceeds 2^32 rows");
    }
    rows += kSuperRows;  // slack so that every TMA box stays inside the tensor
    pool_rows = rows;
    n_images = n;
    B2_CUDA(cudaMalloc(&pool, rows * kDescBytes));
    B2_CUDA(cudaMemsetAsync(pool, 0, rows * kDescBytes, s));
    B2_CUDA(cudaMalloc(&d_img_n, std::max<size_t>(1, n) * sizeof(int32_t)));
    B2_CUDA(cudaMalloc(&d_img_row, std::max<size_t>(1, n) * sizeof(uint32_t)));
    if (n > 0) {
      B2_CUDA(cudaMemcpyAsync(d_img_n, h_img_n.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      B2_CUDA(cudaMemcpyAsync(d_img_row, h_img_row.data(), n * sizeof(uint32_t),
                              cudaMemcpyHostToDevice, s));
    }
    return make_pool_tmap(&tmap, pool, rows, kTileRows);
  }
  // SiftMatchCU::SetDescriptors clamps every upload to max_sift = max_num_matches (SiftMatchCU.cpp:108, sift.cc:200-209):
  // features beyond the clamp take no part in any match.  Applied to the store in place on the first call whose options
  // need it (rows beyond the clamp are zeroed: a zero row can never be a best or second-best match); a later call with a
  // larger max_num_matches needs a new upload, as it does in the reference.
  int clamp(int32_t max_n_per_image, cudaStream_t s) {
    if ((int64_t)max_n <= (int64_t)max_n_per_image) return B2_OK;
    uint32_t new_max = 0;
    for (int32_t i = 0; i < n_images; ++i) {
      if (h_img_n[i] > max_n_per_image) {
        B2_CUDA(cudaMemsetAsync(pool + ((size_t)h_img_row[i] + (size_t)max_n_per_image) * kDescBytes, 0,
                                (size_t)(h_img_n[i] - max_n_per_image) * kDescBytes, s));
        h_img_n[i] = max_n_per_image;
      }
      new_max = std::max<uint32_t>(new_max, (uint32_t)h_img_n[i]);
    }
    max_n = new_max;
    B2_CUDA(cudaMemcpyAsync(d_img_n, h_img_n.data(), (size_t)n_images * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s));  // h_img_n may be re-laid-out by the caller right after
    return B2_OK;
  }
};

// ------------------------------------------------------------ threshold tables
// Device copy of the integer threshold tables of match_thresholds.h.
struct ThresholdTables {
  float max_ratio = -1.f, max_distance = -1.f;
  int thr_dist = 0;
  int* d_ratio_lim = nullptr;  // [kDotClamp + 1]
  HostThresholds host;

  int build(float ratio, float dist, cudaStream_t s) {
    if (ratio == max_ratio && dist == max_distance && d_ratio_lim) return B2_OK;
    if (!host.build(ratio, dist, kDotClamp)) return set_error(B2_ERR_INTERNAL, "host acosf is not monotone");
    if (!d_ratio_lim) B2_CUDA(cudaMalloc(&d_ratio_lim, (kDotClamp + 1) * sizeof(int)));
    B2_CUDA(cudaMemcpyAsync(d_ratio_lim, host.ratio_lim.data(), host.ratio_lim.size() * sizeof(int),
                            cudaMemcpyHostToDevice, s));
    B2_CUDA(cudaStreamSynchronize(s));
    thr_dist = host.thr_dist;
    max_ratio = ratio;
    max_distance = dist;
    return B2_OK;
  }
};

}  // namespace b2

using namespace b2;

struct b2_matcher {
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  ImageStore store;       // batched seam
  ImageStore slots;       // two-slot seam (image 0 / image 1)
  uint32_t slot_cap = 0;  // rows reserved per slot
  ThresholdTables tables;
  // chunk scratch
  uint64_t row_budget = 0;
  int64_t cap_pairs = 0;
  uint32_t* d_nitems = nullptr;      // [cap_pairs]
  uint32_t* d_item_start = nullptr;  // [cap_pairs]
  uint32_t* d_total_items = nullptr; // [1]
  PairMeta* d_meta = nullptr;        // [cap_pairs]
  uint32_t* d_counts = nullptr;      // [cap_pairs]
  MatchItem* d_items = nullptr;      // [row_budget/256]
  uint32_t* d_item_pair = nullptr;   // [row_budget/256] (pair << 1 | direction), guided matching
  GuidedGeom* d_geoms = nullptr; int64_t d_geoms_cap = 0;
  int* d_midx = nullptr;             // [row_budget]
  uint4* d_cands = nullptr;          // [row_budget]
  unsigned int* d_cand_count = nullptr;
  int64_t* d_carry = nullptr;
  int* d_err = nullptr;
  // host-call staging
  uint32_t* d_pairs = nullptr; int64_t d_pairs_cap = 0;
  int64_t* d_offsets = nullptr; int64_t d_offsets_cap = 0;
  uint32_t* d_matches = nullptr; int64_t d_matches_cap = 0;
  // timing of the last call
  std::vector<cudaEvent_t> ev;
  double last_tc_s = 0, last_all_s = 0;
  int64_t last_tc_launches = 0, last_cands = 0;
};

static int ensure_scratch(b2_matcher* m, uint32_t max_n) {
  uint64_t budget = 64ull << 20;  // rows of per-chunk scratch (midx 256 MiB, cands 1 GiB)
  if (const char* e = getenv("B2_MATCH_ROW_BUDGET")) budget = std::max<uint64_t>(1u << 16, strtoull(e, nullptr, 10));
  const uint64_t rows_per_pair = 2ull * pad_up(std::max<uint32_t>(max_n, 1), kSuperRows);  // midx rows (X supertiles)
  budget = std::max(budget, rows_per_pair);
  const int64_t cap_pairs = (int64_t)std::min<uint64_t>(budget / rows_per_pair, 1u << 20);
  if (m->row_budget >= budget && m->cap_pairs >= cap_pairs) {
    m->cap_pairs = cap_pairs;
    return B2_OK;
  }
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(m->d_nitems); fr(m->d_item_start); fr(m->d_meta); fr(m->d_counts); fr(m->d_items); fr(m->d_item_pair);
  m->d_item_pair = nullptr;
  fr(m->d_midx); fr(m->d_cands);
  m->d_nitems = m->d_item_start = m->d_counts = nullptr;
  m->d_meta = nullptr; m->d_items = nullptr; m->d_midx = nullptr; m->d_cands = nullptr;
  B2_CUDA(cudaMalloc(&m->d_nitems, cap_pairs * sizeof(uint32_t)));
  B2_CUDA(cudaMalloc(&m->d_item_start, cap_pairs * sizeof(uint32_t)));
  B2_CUDA(cudaMalloc(&m->d_meta, cap_pairs * sizeof(PairMeta)));
  B2_CUDA(cudaMalloc(&m->d_counts, cap_pairs * sizeof(uint32_t)));
  B2_CUDA(cudaMalloc(&m->d_items, (budget / kSuperRows + 1) * sizeof(MatchItem)));
  B2_CUDA(cudaMalloc(&m->d_item_pair, (budget / kSuperRows + 1) * sizeof(uint32_t)));
  B2_CUDA(cudaMalloc(&m->d_midx, budget * sizeof(int)));
  B2_CUDA(cudaMalloc(&m->d_cands, budget * sizeof(uint4)));
  m->row_budget = budget;
  m->cap_pairs = cap_pairs;
  return B2_OK;
}

// The whole device-side pipeline for n_pairs pairs of `st`; everything async on
// m->stream except the final read-back of {total, err}.
static int run_pairs_device(b2_matcher* m, ImageStore& st, int64_t n_pairs,
                            const uint32_t* pairs_dev, const b2_match_options* opt,
                            int64_t* out_offsets_dev, uint32_t* out_matches_dev, int64_t capacity,
                            int64_t* n_total) {
  if (!opt) return set_error(B2_ERR_INVALID, "options == NULL");
  // SiftMatchingOptions::Check (sift.cc:236-250)
  if (!(opt->max_ratio > 0) || !(opt->max_distance > 0) || opt->max_num_matches <= 0)
    return set_error(B2_ERR_INVALID, "SiftMatchingOptions::Check failed");
  if (n_pairs < 0 || capacity < 0) return set_error(B2_ERR_INVALID, "negative size");
  B2_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  B2_TRY(m->tables.build(opt->max_ratio, opt->max_distance, s));
  B2_TRY(st.clamp(opt->max_num_matches, s));
  B2_TRY(ensure_scratch(m, st.max_n));
  B2_CUDA(cudaMemsetAsync(m->d_carry, 0, sizeof(int64_t), s));
  B2_CUDA(cudaMemsetAsync(m->d_err, 0, sizeof(int), s));
  B2_CUDA(cudaMemsetAsync(out_offsets_dev, 0, sizeof(int64_t), s));

  const int64_t n_chunks = (n_pairs + m->cap_pairs - 1) / std::max<int64_t>(m->cap_pairs, 1);
  while ((int64_t)m->ev.size() < 2 * n_chunks + 2) {
    cudaEvent_t e;
    B2_CUDA(cudaEventCreate(&e));
    m->ev.push_back(e);
  }
  std::vector<unsigned int> h_cands(std::max<int64_t>(n_chunks, 1), 0);
  unsigned int* d_cand_hist = nullptr;
  B2_CUDA(cudaMalloc(&d_cand_hist, std::max<int64_t>(n_chunks, 1) * sizeof(unsigned int)));
  B2_CUDA(cudaEventRecord(m->ev[0], s));
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t p0 = c * m->cap_pairs, np = std::min(m->cap_pairs, n_pairs - p0);
    const uint32_t* pr = pairs_dev + 2 * p0;
    B2_CUDA(launch_pair_items(pr, np, st.d_img_n, st.n_images, m->d_nitems, m->d_err, s));
    B2_CUDA(launch_scan_u32(m->d_nitems, np, m->d_item_start, m->d_total_items, s));
    B2_CUDA(launch_fill_items(pr, np, st.d_img_n, st.n_images, st.d_img_row, m->d_item_start, m->d_items, m->d_meta,
                              (uint32_t)kTileRows, s));
    B2_CUDA(cudaMemsetAsync(m->d_cand_count, 0, sizeof(unsigned int), s));
    B2_CUDA(cudaEventRecord(m->ev[2 + 2 * c], s));
    B2_CUDA(launch_match_top2_ts(st.tmap, st.pool, m->d_items, m->d_total_items, m->tables.thr_dist,
                                 m->tables.d_ratio_lim, m->d_midx, m->d_cands, m->d_cand_count,
                                 (unsigned int)std::min<uint64_t>(m->row_budget, 0xFFFFFFFFu), m->n_sm, s));
    B2_CUDA(cudaEventRecord(m->ev[3 + 2 * c], s));
    B2_CUDA(launch_fixup(st.pool, m->d_items, m->d_cands, m->d_cand_count,
                         (unsigned int)std::min<uint64_t>(m->row_budget, 0xFFFFFFFFu),
                         m->tables.d_ratio_lim, m->d_midx, m->d_err, m->n_sm, s));
    B2_CUDA(cudaMemcpyAsync(d_cand_hist + c, m->d_cand_count, sizeof(unsigned int),
                            cudaMemcpyDeviceToDevice, s));
    B2_CUDA(launch_cross_count(m->d_meta, np, m->d_midx, opt->cross_check, m->d_counts, s));
    B2_CUDA(launch_scan_counts(m->d_counts, np, out_offsets_dev + p0, m->d_carry, true, s));
    B2_CUDA(launch_cross_write(m->d_meta, np, m->d_midx, opt->cross_check, out_offsets_dev + p0,
                               out_matches_dev, capacity, s));
    count_launches(np > 0 ? 8 : 3);
  }
  B2_CUDA(cudaEventRecord(m->ev[1], s));
  int64_t total = 0;
  int err = 0;
  B2_CUDA(cudaMemcpyAsync(&total, m->d_carry, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(&err, m->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(h_cands.data(), d_cand_hist, std::max<int64_t>(n_chunks, 1) * sizeof(unsigned int),
                          cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_cand_hist);
  float ms = 0;
  m->last_tc_s = 0;
  for (int64_t c = 0; c < n_chunks; ++c) {
    B2_CUDA(cudaEventElapsedTime(&ms, m->ev[2 + 2 * c], m->ev[3 + 2 * c]));
    m->last_tc_s += ms * 1e-3;
  }
  B2_CUDA(cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
  m->last_all_s = ms * 1e-3;
  m->last_tc_launches = n_chunks;
  m->last_cands = 0;
  for (int64_t c = 0; c < n_chunks; ++c) m->last_cands += h_cands[c];
  if (n_total) *n_total = total;
  if (err == 1) return set_error(B2_ERR_INVALID, "pair references an image outside the store");
  if (err != 0)
    return set_error(B2_ERR_INTERNAL, "fix-up did not find the tensor-core maximum in its chunk");
  if (total > capacity) return set_error(B2_ERR_CAPACITY, "out_matches capacity too small");
  return B2_OK;
}

// Guided variant of the pipeline: the tensor-core kernel + fix-up are replaced by the thread-per-row
// guided kernel (match_guided.cu); items, cross-check and compaction are shared.
static int run_guided_device(b2_matcher* m, ImageStore& st, int64_t n_pairs, const uint32_t* pairs_dev,
                             const GuidedGeom* geoms_dev, double max_error, const b2_match_options* opt,
                             int64_t* out_offsets_dev, uint32_t* out_matches_dev, int64_t capacity,
                             int64_t* n_total) {
  if (!opt) return set_error(B2_ERR_INVALID, "options == NULL");
  if (!(opt->max_ratio > 0) || !(opt->max_distance > 0) || opt->max_num_matches <= 0 || !(max_error > 0))
    return set_error(B2_ERR_INVALID, "SiftMatchingOptions::Check failed");
  if (n_pairs < 0 || capacity < 0) return set_error(B2_ERR_INVALID, "negative size");
  if (!st.kp_pool) return set_error(B2_ERR_INVALID, "guided matching needs b2_match_set_keypoints first");
  B2_CUDA(cudaSetDevice(m->device));
  cudaStream_t s = m->stream;
  B2_TRY(m->tables.build(opt->max_ratio, opt->max_distance, s));
  B2_TRY(st.clamp(opt->max_num_matches, s));
  B2_TRY(ensure_scratch(m, st.max_n));
  B2_CUDA(cudaMemsetAsync(m->d_carry, 0, sizeof(int64_t), s));
  B2_CUDA(cudaMemsetAsync(m->d_err, 0, sizeof(int), s));
  B2_CUDA(cudaMemsetAsync(out_offsets_dev, 0, sizeof(int64_t), s));
  const float max_residual = (float)(max_error * max_error);  // sift.cc:833
  const int64_t n_chunks = (n_pairs + m->cap_pairs - 1) / std::max<int64_t>(m->cap_pairs, 1);
  while ((int64_t)m->ev.size() < 2) {
    cudaEvent_t e;
    B2_CUDA(cudaEventCreate(&e));
    m->ev.push_back(e);
  }
  B2_CUDA(cudaEventRecord(m->ev[0], s));
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t p0 = c * m->cap_pairs, np = std::min(m->cap_pairs, n_pairs - p0);
    const uint32_t* pr = pairs_dev + 2 * p0;
    B2_CUDA(launch_pair_items(pr, np, st.d_img_n, st.n_images, m->d_nitems, m->d_err, s));
    B2_CUDA(launch_scan_u32(m->d_nitems, np, m->d_item_start, m->d_total_items, s));
    B2_CUDA(launch_fill_items(pr, np, st.d_img_n, st.n_images, st.d_img_row, m->d_item_start, m->d_items, m->d_meta,
                              (uint32_t)kTileRows, s));
    B2_CUDA(launch_guided_item_pairs(m->d_meta, np, m->d_item_pair, s));
    B2_CUDA(launch_guided_match(st.pool, st.kp_pool, m->d_items, m->d_item_pair, m->d_total_items, m->d_meta,
                                geoms_dev + p0, max_residual, m->tables.thr_dist, m->tables.d_ratio_lim, m->d_midx,
                                m->n_sm, s));
    B2_CUDA(launch_cross_count(m->d_meta, np, m->d_midx, opt->cross_check, m->d_counts, s));
    B2_CUDA(launch_scan_counts(m->d_counts, np, out_offsets_dev + p0, m->d_carry, true, s));
    B2_CUDA(launch_cross_write(m->d_meta, np, m->d_midx, opt->cross_check, out_offsets_dev + p0,
                               out_matches_dev, capacity, s));
    count_launches(np > 0 ? 8 : 3);
  }
  B2_CUDA(cudaEventRecord(m->ev[1], s));
  int64_t total = 0;
  int err = 0;
  B2_CUDA(cudaMemcpyAsync(&total, m->d_carry, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(&err, m->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
  m->last_all_s = ms * 1e-3;
  m->last_tc_s = 0;
  m->last_tc_launches = 0;
  m->last_cands = 0;
  if (n_total) *n_total = total;
  if (err == 1) return set_error(B2_ERR_INVALID, "pair references an image outside the store");
  if (total > capacity) return set_error(B2_ERR_CAPACITY, "out_matches capacity too small");
  return B2_OK;
}

extern "C" {

void b2_match_default_options(b2_match_options* opt) {
  if (!opt) return;
  opt->max_ratio = 0.8f;
  opt->max_distance = 0.7f;
  opt->cross_check = 1;
  opt->max_num_matches = 32768;
}

int b2_match_create(int device, b2_matcher** out) {
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
  if (prop.major != 10)
    return set_error(B2_ERR_NO_DEVICE, "device is not sm_100 (this library ships sm_100a code only)");
  B2_CUDA(cudaSetDevice(device));
  b2_matcher* m = new b2_matcher();
  m->device = device;
  m->n_sm = prop.multiProcessorCount;
  const int rc = [&]() -> int {
    B2_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    B2_CUDA(cudaMalloc(&m->d_total_items, sizeof(uint32_t)));
    B2_CUDA(cudaMalloc(&m->d_cand_count, sizeof(unsigned int)));
    B2_CUDA(cudaMalloc(&m->d_carry, sizeof(int64_t)));
    B2_CUDA(cudaMalloc(&m->d_err, sizeof(int)));
    return B2_OK;
  }();
  if (rc != B2_OK) {  // a half-built handle is released here, never handed out
    b2_match_destroy(m);
    return rc;
  }
  *out = m;
  return B2_OK;
}

int b2_match_destroy(b2_matcher* m) {
  if (!m) return B2_OK;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  m->store.release();
  m->slots.release();
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(m->tables.d_ratio_lim);
  fr(m->d_nitems); fr(m->d_item_start); fr(m->d_total_items); fr(m->d_meta); fr(m->d_counts);
  fr(m->d_items); fr(m->d_midx); fr(m->d_cands); fr(m->d_cand_count); fr(m->d_carry); fr(m->d_err);
  fr(m->d_pairs); fr(m->d_offsets); fr(m->d_matches); fr(m->d_item_pair); fr(m->d_geoms);
  for (auto e : m->ev) if (e) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return B2_OK;
}

int b2_match_set_images(b2_matcher* m, int32_t n_images, const uint8_t* const* desc,
                        const int32_t* n_desc) {
  if (!m || (n_images > 0 && (!desc || !n_desc))) return set_error(B2_ERR_INVALID, "NULL argument");
  B2_CUDA(cudaSetDevice(m->device));
  B2_TRY(m->store.layout(n_images, n_desc, m->stream));
  // images that follow each other in the caller's memory AND in the pool (a contiguous descriptor array of images whose
  // counts are multiples of the pool's row padding) travel as one copy: one DMA instead of thousands of 256 KB ones
  const uint8_t* run_src = nullptr;
  uint8_t* run_dst = nullptr;
  size_t run_bytes = 0;
  auto flush = [&]() -> cudaError_t {
    if (!run_bytes) return cudaSuccess;
    const cudaError_t e = cudaMemcpyAsync(run_dst, run_src, run_bytes, cudaMemcpyHostToDevice, m->stream);
    run_bytes = 0;
    return e;
  };
  for (int32_t i = 0; i < n_images; ++i) {
    if (n_desc[i] == 0) continue;
    if (!desc[i]) return set_error(B2_ERR_INVALID, "NULL descriptor pointer");
    uint8_t* dst = m->store.pool + (size_t)m->store.h_img_row[i] * kDescBytes;
    const size_t bytes = (size_t)n_desc[i] * kDescBytes;
    if (run_bytes && desc[i] == run_src + run_bytes && dst == run_dst + run_bytes) {
      run_bytes += bytes;
      continue;
    }
    B2_CUDA(flush());
    run_src = desc[i];
    run_dst = dst;
    run_bytes = bytes;
  }
  B2_CUDA(flush());
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_set_images_device(b2_matcher* m, int32_t n_images, const uint8_t* desc_dev,
                               const int64_t* row_offset, const int32_t* n_desc) {
  if (!m || (n_images > 0 && (!desc_dev || !row_offset || !n_desc)))
    return set_error(B2_ERR_INVALID, "NULL argument");
  B2_CUDA(cudaSetDevice(m->device));
  B2_TRY(m->store.layout(n_images, n_desc, m->stream));
  for (int32_t i = 0; i < n_images; ++i) {
    if (n_desc[i] == 0) continue;
    B2_CUDA(cudaMemcpyAsync(m->store.pool + (size_t)m->store.h_img_row[i] * kDescBytes,
                            desc_dev + (size_t)row_offset[i] * kDescBytes,
                            (size_t)n_desc[i] * kDescBytes, cudaMemcpyDeviceToDevice, m->stream));
  }
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_pairs_device(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs_dev,
                          const b2_match_options* opt, int64_t* out_offsets_dev,
                          uint32_t* out_matches_dev, int64_t capacity, int64_t* n_total) {
  if (!m || !out_offsets_dev || (n_pairs > 0 && !pairs_dev) || (capacity > 0 && !out_matches_dev))
    return set_error(B2_ERR_INVALID, "NULL argument");
  return run_pairs_device(m, m->store, n_pairs, pairs_dev, opt, out_offsets_dev, out_matches_dev,
                          capacity, n_total);
}

static int grow(void** p, int64_t* cap, int64_t need, size_t elem) {
  if (*cap >= need && *p) return B2_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  const int64_t n = std::max<int64_t>(need, 1);
  B2_CUDA(cudaMalloc(p, n * elem));
  *cap = n;
  return B2_OK;
}

static int run_pairs_host(b2_matcher* m, ImageStore& st, int64_t n_pairs, const uint32_t* pairs,
                          const b2_match_options* opt, int64_t* out_offsets, uint32_t* out_matches,
                          int64_t capacity, int64_t* n_total) {
  B2_CUDA(cudaSetDevice(m->device));
  B2_TRY(grow((void**)&m->d_pairs, &m->d_pairs_cap, n_pairs, 2 * sizeof(uint32_t)));
  B2_TRY(grow((void**)&m->d_offsets, &m->d_offsets_cap, n_pairs + 1, sizeof(int64_t)));
  B2_TRY(grow((void**)&m->d_matches, &m->d_matches_cap, capacity, 2 * sizeof(uint32_t)));
  if (n_pairs > 0)
    B2_CUDA(cudaMemcpyAsync(m->d_pairs, pairs, n_pairs * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            m->stream));
  int64_t total = 0;
  const int rc = run_pairs_device(m, st, n_pairs, m->d_pairs, opt, m->d_offsets, m->d_matches,
                                  capacity, &total);
  if (n_total) *n_total = total;
  if (rc != B2_OK) return rc;
  B2_CUDA(cudaMemcpyAsync(out_offsets, m->d_offsets, (n_pairs + 1) * sizeof(int64_t),
                          cudaMemcpyDeviceToHost, m->stream));
  if (total > 0)
    B2_CUDA(cudaMemcpyAsync(out_matches, m->d_matches, total * 2 * sizeof(uint32_t),
                            cudaMemcpyDeviceToHost, m->stream));
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_pairs(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs, const b2_match_options* opt,
                   int64_t* out_offsets, uint32_t* out_matches, int64_t capacity, int64_t* n_total) {
  if (!m || !out_offsets || (n_pairs > 0 && !pairs) || (capacity > 0 && !out_matches))
    return set_error(B2_ERR_INVALID, "NULL argument");
  return run_pairs_host(m, m->store, n_pairs, pairs, opt, out_offsets, out_matches, capacity, n_total);
}

int b2_match_set_keypoints(b2_matcher* m, int32_t n_images, const float* const* xy, const int32_t* n_keypoints) {
  if (!m || (n_images > 0 && (!xy || !n_keypoints))) return set_error(B2_ERR_INVALID, "NULL argument");
  ImageStore& st = m->store;
  if (n_images != st.n_images || !st.pool) return set_error(B2_ERR_INVALID, "keypoints must follow b2_match_set_images of the same images");
  for (int32_t i = 0; i < n_images; ++i)
    if (n_keypoints[i] != st.h_img_n[i])  // sift.cc:82-87 CHECK_EQ(keypoints->size(), descriptors.rows())
      return set_error(B2_ERR_INVALID, "keypoint count differs from descriptor count");
  B2_CUDA(cudaSetDevice(m->device));
  if (!st.kp_pool) B2_CUDA(cudaMalloc(&st.kp_pool, st.pool_rows * 2 * sizeof(float)));
  B2_CUDA(cudaMemsetAsync(st.kp_pool, 0, st.pool_rows * 2 * sizeof(float), m->stream));
  for (int32_t i = 0; i < n_images; ++i) {
    if (n_keypoints[i] == 0) continue;
    if (!xy[i]) return set_error(B2_ERR_INVALID, "NULL keypoint pointer");
    B2_CUDA(cudaMemcpyAsync(st.kp_pool + 2 * (size_t)st.h_img_row[i], xy[i], (size_t)n_keypoints[i] * 2 * sizeof(float),
                            cudaMemcpyHostToDevice, m->stream));
  }
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_guided_pairs(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs, const b2_guided_geometry* geoms,
                          double max_error, const b2_match_options* opt, int64_t* out_offsets,
                          uint32_t* out_matches, int64_t capacity, int64_t* n_total) {
  if (!m || !out_offsets || (n_pairs > 0 && (!pairs || !geoms)) || (capacity > 0 && !out_matches))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs < 0 || capacity < 0) return set_error(B2_ERR_INVALID, "negative size");
  B2_CUDA(cudaSetDevice(m->device));
  B2_TRY(grow((void**)&m->d_pairs, &m->d_pairs_cap, n_pairs, 2 * sizeof(uint32_t)));
  B2_TRY(grow((void**)&m->d_offsets, &m->d_offsets_cap, n_pairs + 1, sizeof(int64_t)));
  B2_TRY(grow((void**)&m->d_matches, &m->d_matches_cap, capacity, 2 * sizeof(uint32_t)));
  B2_TRY(grow((void**)&m->d_geoms, &m->d_geoms_cap, n_pairs, sizeof(GuidedGeom)));
  std::vector<GuidedGeom> hg((size_t)n_pairs);
  for (int64_t p = 0; p < n_pairs; ++p) hg[p] = make_guided_geom(geoms[p].config, geoms[p].F, geoms[p].H);
  if (n_pairs > 0) {
    B2_CUDA(cudaMemcpyAsync(m->d_pairs, pairs, n_pairs * 2 * sizeof(uint32_t), cudaMemcpyHostToDevice, m->stream));
    B2_CUDA(cudaMemcpyAsync(m->d_geoms, hg.data(), n_pairs * sizeof(GuidedGeom), cudaMemcpyHostToDevice, m->stream));
    B2_CUDA(cudaStreamSynchronize(m->stream));  // hg is a stack-lifetime host buffer
  }
  int64_t total = 0;
  const int rc = run_guided_device(m, m->store, n_pairs, m->d_pairs, m->d_geoms, max_error, opt, m->d_offsets,
                                   m->d_matches, capacity, &total);
  if (n_total) *n_total = total;
  if (rc != B2_OK) return rc;
  B2_CUDA(cudaMemcpyAsync(out_offsets, m->d_offsets, (n_pairs + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream));
  if (total > 0)
    B2_CUDA(cudaMemcpyAsync(out_matches, m->d_matches, total * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, m->stream));
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_guided_pairs_device(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs_dev, const b2_two_view_result* results_dev,
                                 int32_t min_num_inliers, double max_error, const b2_match_options* opt,
                                 int64_t* out_offsets_dev, uint32_t* out_matches_dev, int64_t capacity, int64_t* n_total) {
  if (!m || !out_offsets_dev || (n_pairs > 0 && (!pairs_dev || !results_dev)) || (capacity > 0 && !out_matches_dev))
    return set_error(B2_ERR_INVALID, "NULL argument");
  if (n_pairs < 0 || capacity < 0) return set_error(B2_ERR_INVALID, "negative size");
  B2_CUDA(cudaSetDevice(m->device));
  B2_TRY(grow((void**)&m->d_geoms, &m->d_geoms_cap, n_pairs, sizeof(GuidedGeom)));
  B2_CUDA(launch_geoms_from_results(n_pairs, results_dev, min_num_inliers, m->d_geoms, m->stream));
  count_launches(n_pairs > 0 ? 1 : 0);
  int64_t total = 0;
  const int rc = run_guided_device(m, m->store, n_pairs, pairs_dev, m->d_geoms, max_error, opt, out_offsets_dev, out_matches_dev,
                                   capacity, &total);
  if (n_total) *n_total = total;
  return rc;
}

int b2_match_set_descriptors(b2_matcher* m, int slot, int32_t n, const uint8_t* desc) {
  if (!m || slot < 0 || slot > 1 || n < 0) return set_error(B2_ERR_INVALID, "bad slot / count");
  B2_CUDA(cudaSetDevice(m->device));
  ImageStore& st = m->slots;
  const uint32_t need = image_rows(std::max<uint32_t>((uint32_t)n, 1));
  if (st.n_images != 2 || need > m->slot_cap) {
    // (re)build the two-slot store with room for `need` rows per slot, keeping the other slot
    const uint32_t new_cap = std::max(need, m->slot_cap);
    ImageStore old = st;  // shallow copy of pointers
    std::vector<int32_t> keep_n = old.h_img_n;
    st = ImageStore();
    const int32_t caps[2] = {(int32_t)new_cap, (int32_t)new_cap};
    B2_TRY(st.layout(2, caps, m->stream));
    st.h_img_n = {0, 0};
    if (old.n_images == 2) {
      for (int k = 0; k < 2; ++k) {
        if (k == slot || keep_n[k] == 0) continue;
        B2_CUDA(cudaMemcpyAsync(st.pool + (size_t)st.h_img_row[k] * kDescBytes,
                                old.pool + (size_t)old.h_img_row[k] * kDescBytes,
                                (size_t)keep_n[k] * kDescBytes, cudaMemcpyDeviceToDevice, m->stream));
        st.h_img_n[k] = keep_n[k];
      }
      B2_CUDA(cudaStreamSynchronize(m->stream));
      old.release();
    }
    m->slot_cap = new_cap;
  }
  if (desc) {  // NULL keeps the previous upload (sift.h:232-234)
    const uint32_t old_n = (uint32_t)st.h_img_n[slot];
    uint8_t* base = st.pool + (size_t)st.h_img_row[slot] * kDescBytes;
    if (n > 0) B2_CUDA(cudaMemcpyAsync(base, desc, (size_t)n * kDescBytes, cudaMemcpyHostToDevice, m->stream));
    const uint32_t zero_to = image_rows(std::max<uint32_t>(old_n, (uint32_t)n));
    if (zero_to > (uint32_t)n)
      B2_CUDA(cudaMemsetAsync(base + (size_t)n * kDescBytes, 0, (size_t)(zero_to - n) * kDescBytes, m->stream));
    st.h_img_n[slot] = n;
  }
  st.max_n = (uint32_t)std::max(st.h_img_n[0], st.h_img_n[1]);
  B2_CUDA(cudaMemcpyAsync(st.d_img_n, st.h_img_n.data(), 2 * sizeof(int32_t), cudaMemcpyHostToDevice, m->stream));
  B2_CUDA(cudaStreamSynchronize(m->stream));
  return B2_OK;
}

int b2_match_run(b2_matcher* m, const b2_match_options* opt, uint32_t* out_matches, int32_t* n_out) {
  if (!m || !opt || !n_out) return set_error(B2_ERR_INVALID, "NULL argument");
  *n_out = 0;
  if (m->slots.n_images != 2) return set_error(B2_ERR_INVALID, "descriptors not set");
  // SiftMatchCU clamps both sets to max_sift = max_num_matches at upload time (SiftMatchCU.cpp:108); run_pairs_device
  // applies that clamp to the slots.  At most one match per row of image 0 comes back (without cross-check several rows
  // may share a column, so the count can exceed n1); GetSiftMatch returns the first max_match of them (SiftMatchCU.cpp:
  // 178-199 stops filling the caller's buffer at max_match).
  ImageStore& st = m->slots;
  const uint32_t pair[2] = {0, 1};
  int64_t offsets[2] = {0, 0};
  int64_t total = 0;
  const int64_t cap = std::min<int32_t>(st.h_img_n[0], opt->max_num_matches > 0 ? opt->max_num_matches : 0);
  std::vector<uint32_t> tmp(2 * std::max<int64_t>(cap, 1));
  const int rc = run_pairs_host(m, st, 1, pair, opt, offsets, tmp.data(), std::max<int64_t>(cap, 1), &total);
  if (rc != B2_OK) return rc;
  const int64_t n_ret = std::min<int64_t>(total, opt->max_num_matches);
  if (n_ret > 0 && !out_matches) return set_error(B2_ERR_INVALID, "out_matches == NULL");
  if (n_ret > 0) memcpy(out_matches, tmp.data(), (size_t)n_ret * 2 * sizeof(uint32_t));
  *n_out = (int32_t)n_ret;
  return B2_OK;
}

int b2_match_last_timing(b2_matcher* m, double* tc_kernel_s, double* all_kernels_s,
                         int64_t* tc_launches, int64_t* fixup_candidates) {
  if (!m) return set_error(B2_ERR_INVALID, "NULL matcher");
  if (tc_kernel_s) *tc_kernel_s = m->last_tc_s;
  if (all_kernels_s) *all_kernels_s = m->last_all_s;
  if (tc_launches) *tc_launches = m->last_tc_launches;
  if (fixup_candidates) *fixup_candidates = m->last_cands;
  return B2_OK;
}

}  // extern "C"
