// Host harness of the EstimateMultiple round loop (dagsfm_b200/csrc/verify_multiple.h) with the
// batched estimator supplied by the caller as a C callback -- the CPU suite plugs the oracle in and
// checks the loop the product runs around the GPU kernel.  Test infrastructure only.
#include "../../dagsfm_b200/csrc/verify_multiple.h"

typedef int (*estimate_cb)(int64_t n_active, const int64_t* pair_ids, const int64_t* offsets, const uint32_t* matches,
                           const uint32_t* seeds, b2_two_view_result* results, uint32_t* inliers);

extern "C" int host_estimate_multiple(int64_t n_pairs, const int64_t* match_offsets, const uint32_t* matches,
                                      const uint32_t* seeds, int ignore_watermark, estimate_cb cb,
                                      b2_two_view_result* results, uint32_t* inlier_matches) {
  return b2::estimate_multiple(n_pairs, match_offsets, matches, seeds, ignore_watermark != 0, cb, results,
                               inlier_matches);
}
