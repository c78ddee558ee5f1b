// Kernels of the guided matcher, kept apart from their launchers so that tests/cpp/host_guided_kernel.cc
// can compile these very functions for the host (they use no warp intrinsics and no shared memory)
// and run them block by block, thread by thread, against the oracle.
#pragma once
#include <cstdint>

#include "match_common.cuh"
#include "match_guided.cuh"

namespace b2 {

// One thread per pair: (pair, direction) of each of its items, so that a block can find the
// geometry and the row/column counts of the item it was handed.
__global__ void guided_item_pairs_kernel(const PairMeta* __restrict__ meta, int64_t n_pairs,
                                         uint32_t* __restrict__ item_pair) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const PairMeta pm = meta[p];
  if (pm.n1 == 0 || pm.n2 == 0) return;
  const uint32_t nt2 = (pm.n2 + kSuperRows - 1) / kSuperRows;
  for (uint32_t t = 0; t < pm.nt1; ++t) item_pair[pm.item_start + t] = (uint32_t)(p << 1);
  for (uint32_t t = 0; t < nt2; ++t) item_pair[pm.item_start + pm.nt1 + t] = (uint32_t)(p << 1) | 1u;
}

// One block per item (256 rows of the X image against the whole Y image), one thread per row.
__global__ void __launch_bounds__(kSuperRows)
guided_match_kernel(const uint8_t* __restrict__ pool, const float* __restrict__ kp_pool,
                    const MatchItem* __restrict__ items, const uint32_t* __restrict__ item_pair,
                    const uint32_t* __restrict__ n_items_ptr, const PairMeta* __restrict__ meta,
                    const GuidedGeom* __restrict__ geoms, float max_residual, int thr_dist,
                    const int* __restrict__ ratio_lim, int* __restrict__ midx) {
  const uint32_t n_items = *n_items_ptr;
  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const MatchItem w = items[item];
    const uint32_t pd = item_pair[item];
    const uint32_t p = pd >> 1, dir = pd & 1u;
    const PairMeta pm = meta[p];
    const GuidedGeom g = geoms[p];
    const uint32_t tile = item - (pm.item_start + (dir ? pm.nt1 : 0u));
    const uint32_t n_x = dir ? pm.n2 : pm.n1, n_y = dir ? pm.n1 : pm.n2;
    const uint32_t row_in_image = tile * kSuperRows + threadIdx.x;
    int result = -1;
    if (g.kind != 0 && row_in_image < n_x) {
      const uint32_t xr = w.x_row + threadIdx.x;
      uint32_t xd[32];
      const uint4* src = reinterpret_cast<const uint4*>(pool + (size_t)xr * kDescBytes);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(src + q);
        xd[4 * q] = v.x; xd[4 * q + 1] = v.y; xd[4 * q + 2] = v.z; xd[4 * q + 3] = v.w;
      }
      result = guided_row_best(g, max_residual, dir == 0, xd, kp_pool[2 * (size_t)xr], kp_pool[2 * (size_t)xr + 1],
                               pool + (size_t)w.y_row * kDescBytes, kp_pool + 2 * (size_t)w.y_row, (int)n_y,
                               thr_dist, ratio_lim, kDotClamp);
    }
    midx[(size_t)item * kSuperRows + threadIdx.x] = result;
  }
}

}  // namespace b2
