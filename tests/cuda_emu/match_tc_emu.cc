// Host stand-in for the tcgen05 kernel (match_tc_ts.cu), which has no CPU meaning: it
// produces what that kernel is CONTRACTED to produce (match_common.cuh) -- per query
// row the exact best dot, the first 32-column chunk attaining it and the best maximum over all other
// chunks; rows that can still pass the integer threshold tests go to the candidate list, midx = -1 --
// so that the rest of the matcher (items, fix-up, cross-check, compaction, chunking in match_api.cu: all
// real code) can be run against the oracle on the CPU.  TEST INFRASTRUCTURE ONLY.
#include <cuda.h>
#include <cuda_runtime.h>

#include "match_common.cuh"

namespace b2 {

static cudaError_t tc_stage(const uint8_t* pool, const MatchItem* items, const uint32_t* n_items_ptr, int thr_dist,
                            const int* ratio_lim, int* midx, uint4* cands, unsigned int* cand_count,
                            unsigned int cand_capacity) {
  const uint32_t n_items = *n_items_ptr;
  for (uint32_t item = 0; item < n_items; ++item) {
    const MatchItem w = items[item];
    for (uint32_t r = 0; r < (uint32_t)kSuperRows; ++r) {
      const uint8_t* x = pool + (size_t)(w.x_row + r) * kDescBytes;
      int best = 0, second = 0, bchunk = 0;
      for (uint32_t c = 0; c < w.y_nblk * 4; ++c) {
        int m = 0;
        for (uint32_t j = 0; j < (uint32_t)kChunk; ++j) {
          const uint8_t* y = pool + (size_t)(w.y_row + c * kChunk + j) * kDescBytes;
          int d = 0;
          for (int k = 0; k < kDescBytes; ++k) d += (int)x[k] * (int)y[k];
          m = d > m ? d : m;
        }
        second = second > (best < m ? best : m) ? second : (best < m ? best : m);
        bchunk = (m > best) ? (int)c : bchunk;
        best = best > m ? best : m;
      }
      const uint32_t out = item * kSuperRows + r;
      midx[out] = -1;
      if (best >= thr_dist && second <= ratio_lim[best < kDotClamp ? best : kDotClamp]) {
        const unsigned pos = (*cand_count)++;
        if (pos < cand_capacity) cands[pos] = make_uint4(out, (uint32_t)bchunk, (uint32_t)best, (uint32_t)second);
      }
    }
  }
  return cudaSuccess;
}

cudaError_t launch_match_top2_ts(const CUtensorMap&, const uint8_t* pool, const MatchItem* items, const uint32_t* n_items_ptr,
                                 int thr_dist, const int* ratio_lim, int* midx, uint4* cands, unsigned int* cand_count,
                                 unsigned int cand_capacity, int, cudaStream_t) {
  cuda_emu::DeviceWindow window;
  return tc_stage(pool, items, n_items_ptr, thr_dist, ratio_lim, midx, cands, cand_count, cand_capacity);
}

}  // namespace b2
