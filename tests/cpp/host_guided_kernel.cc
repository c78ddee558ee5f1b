// Runs the guided matcher's CUDA kernels (dagsfm_b200/csrc/match_guided_kernels.cuh) on the HOST: the kernel
// bodies use neither warp intrinsics nor shared memory, so with blockIdx / threadIdx as globals they are
// plain C++ and can be executed block by block, thread by thread.  The surrounding pipeline stages
// (items, cross-check) are re-stated from match_post.cu as the launch configuration the library uses.
// Test infrastructure only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static Dim3 blockIdx, threadIdx, gridDim, blockDim;
struct uint4 { unsigned x, y, z, w; };
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
template <class T> static inline T __ldg(const T* p) { return *p; }
#define B2_HD inline

#include "../../dagsfm_b200/csrc/match_guided_kernels.cuh"
#include "../../dagsfm_b200/csrc/match_thresholds.h"

using namespace b2;

static uint32_t pad_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }
static uint32_t image_rows(uint32_t n) { return pad_up(pad_up(n, 96), 256); }   // match_api.cu

// n_images images (descriptors + keypoints), n_pairs pairs with one geometry each.  Lays the images out in a
// pool like ImageStore::layout, builds items/meta like pair_items + fill_items, runs the two guided kernels
// with the library's launch shapes, then the cross-check of match_cross_kernel.  out_offsets[n_pairs + 1].
extern "C" int host_guided_pipeline(int n_images, const uint8_t* const* desc, const float* const* kp, const int* n_desc,
                                    int n_pairs, const uint32_t* pairs, const int* config, const double* F /*[n_pairs][9]*/,
                                    const double* H, double max_error, float max_ratio, float max_distance,
                                    int cross_check, int grid_blocks, int64_t* out_offsets, uint32_t* out, int64_t cap) {
  // ---- image store
  std::vector<uint32_t> img_row(n_images);
  uint64_t rows = 0;
  for (int i = 0; i < n_images; ++i) { img_row[i] = (uint32_t)rows; rows += image_rows((uint32_t)n_desc[i]); }
  rows += kSuperRows;
  std::vector<uint8_t> pool(rows * kDescBytes, 0);
  std::vector<float> kp_pool(rows * 2, 0.0f);
  for (int i = 0; i < n_images; ++i) {
    if (n_desc[i] == 0) continue;
    memcpy(&pool[(size_t)img_row[i] * kDescBytes], desc[i], (size_t)n_desc[i] * kDescBytes);
    memcpy(&kp_pool[2 * (size_t)img_row[i]], kp[i], (size_t)n_desc[i] * 2 * sizeof(float));
  }
  // ---- items + meta (pair_items_kernel / fill_items_kernel, match_post.cu)
  std::vector<PairMeta> meta(n_pairs);
  std::vector<MatchItem> items;
  for (int p = 0; p < n_pairs; ++p) {
    const uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
    const uint32_t n1 = n_desc[i1], n2 = n_desc[i2];
    PairMeta pm;
    pm.item_start = (uint32_t)items.size(); pm.n1 = n1; pm.n2 = n2; pm.nt1 = 0;
    if (n1 > 0 && n2 > 0) {
      const uint32_t nt1 = (n1 + kSuperRows - 1) / kSuperRows, nt2 = (n2 + kSuperRows - 1) / kSuperRows;
      pm.nt1 = nt1;
      for (uint32_t t = 0; t < nt1; ++t) items.push_back(MatchItem{img_row[i1] + t * kSuperRows, img_row[i2], (n2 + 127) / 128, 0u});
      for (uint32_t t = 0; t < nt2; ++t) items.push_back(MatchItem{img_row[i2] + t * kSuperRows, img_row[i1], (n1 + 127) / 128, 0u});
    }
    meta[p] = pm;
  }
  const uint32_t n_items = (uint32_t)items.size();
  std::vector<uint32_t> item_pair(std::max<size_t>(items.size(), 1), 0xdeadbeefu);
  std::vector<int> midx(std::max<size_t>((size_t)n_items * kSuperRows, 1), -7);
  std::vector<GuidedGeom> geoms(n_pairs);
  for (int p = 0; p < n_pairs; ++p) geoms[p] = make_guided_geom(config[p], F + 9 * p, H + 9 * p);
  static HostThresholds th;
  if (!th.build(max_ratio, max_distance, kDotClamp)) return -3;
  // ---- guided_item_pairs_kernel<<<ceil(n_pairs / 256), 256>>>
  gridDim.x = (unsigned)((n_pairs + 255) / 256); blockDim.x = 256;
  for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
    for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x)
      guided_item_pairs_kernel(meta.data(), n_pairs, item_pair.data());
  // ---- guided_match_kernel<<<grid_blocks, 256>>> (grid-stride over items)
  gridDim.x = (unsigned)grid_blocks; blockDim.x = kSuperRows;
  const float max_residual = (float)(max_error * max_error);
  for (blockIdx.x = 0; blockIdx.x < gridDim.x; ++blockIdx.x)
    for (threadIdx.x = 0; threadIdx.x < blockDim.x; ++threadIdx.x)
      guided_match_kernel(pool.data(), kp_pool.data(), items.data(), item_pair.data(), &n_items, meta.data(), geoms.data(),
                          max_residual, th.thr_dist, th.ratio_lim.data(), midx.data());
  // ---- match_cross_kernel semantics
  int64_t total = 0;
  out_offsets[0] = 0;
  for (int p = 0; p < n_pairs; ++p) {
    const PairMeta pm = meta[p];
    if (pm.n1 > 0 && pm.n2 > 0) {
      const int* m12 = midx.data() + (size_t)pm.item_start * kSuperRows;
      const int* m21 = midx.data() + (size_t)(pm.item_start + pm.nt1) * kSuperRows;
      for (uint32_t i = 0; i < pm.n1; ++i) {
        const int j = m12[i];
        bool ok = j >= 0;
        if (ok && cross_check) ok = (m21[j] == (int)i);
        if (!ok) continue;
        if (total < cap) { out[2 * total] = i; out[2 * total + 1] = (uint32_t)j; }
        ++total;
      }
    }
    out_offsets[p + 1] = total;
  }
  return (int)total;
}
