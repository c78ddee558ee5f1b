// Guided descriptor matching kernels -- see match_guided.cuh for the algorithm.
// Compiled with --fmad=false: the float32 filter must round like the oracle's.
#include <cuda_runtime.h>

#include <cstdint>

#include "match_common.cuh"
#include "match_guided.cuh"
#include "match_guided_kernels.cuh"
#include "../../include/dagsfm_b200.h"

namespace b2 {

cudaError_t launch_guided_item_pairs(const PairMeta* meta, int64_t n_pairs, uint32_t* item_pair, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  guided_item_pairs_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(meta, n_pairs, item_pair);
  return cudaGetLastError();
}
cudaError_t launch_guided_match(const uint8_t* pool, const float* kp_pool, const MatchItem* items,
                                const uint32_t* item_pair, const uint32_t* n_items_ptr, const PairMeta* meta,
                                const GuidedGeom* geoms, float max_residual, int thr_dist, const int* ratio_lim,
                                int* midx, int n_sm, cudaStream_t s) {
  guided_match_kernel<<<n_sm * 8, kSuperRows, 0, s>>>(pool, kp_pool, items, item_pair, n_items_ptr, meta, geoms,
                                                      max_residual, thr_dist, ratio_lim, midx);
  return cudaGetLastError();
}

// Geometry of every pair straight from the verifier's device results (the guided stage of the reference's pipeline,
// GuidedSiftGPUFeatureMatcher::Run, matching.cc:493-530): pairs with fewer than min_num_inliers inliers are passed
// through unmatched (:508-512), configurations without a guided filter too (sift.cc:1049-1051).
__global__ void geoms_from_results_kernel(int64_t n_pairs, const b2_two_view_result* __restrict__ results, int min_num_inliers,
                                          GuidedGeom* __restrict__ geoms) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const b2_two_view_result& r = results[p];
  GuidedGeom g = make_guided_geom(r.n_inliers < min_num_inliers ? 0 : r.config, r.F, r.H);
  geoms[p] = g;
}
cudaError_t launch_geoms_from_results(int64_t n_pairs, const void* results, int min_num_inliers, GuidedGeom* geoms, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  geoms_from_results_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(n_pairs, (const b2_two_view_result*)results,
                                                                            min_num_inliers, geoms);
  return cudaGetLastError();
}

}  // namespace b2
