// TwoViewGeometry::EstimateMultiple (src/estimators/two_view_geometry.cc:128-167) for a batch of
// pairs, expressed as ROUNDS of the batched Estimate: every still-active pair is estimated on its
// remaining matches, the inliers of a non-degenerate geometry are removed (ExtractOutlierMatches,
// :67-88, removal BY VALUE of the (idx1, idx2) pair), and the pair stays active until a round comes
// back DEGENERATE.  One geometry -> that geometry; several -> config MULTIPLE with the inlier lists
// concatenated in round order (E, F, H stay zero like the default-constructed TwoViewGeometry).
//
// Host-only, templated on the batched estimator so that the product (GPU kernel behind
// b2_verify_pairs) and the CPU test harness (oracle behind the same loop) run the same code.
//
// Randomness: the reference continues its thread-local PRNG from one Estimate to the next; here the
// pair's seed is re-derived per round, seed_r = seed + r * 0x9E3779B9 (mod 2^32), so that a round is
// an ordinary seeded Estimate and the whole procedure stays reproducible.
#pragma once
#include <cstdint>
#include <cstring>
#include <set>
#include <utility>
#include <vector>

#include "../../include/dagsfm_b200.h"

namespace b2 {

constexpr int32_t kConfigDegenerate = 1, kConfigWatermark = 7, kConfigMultiple = 8;
constexpr uint32_t kRoundSeedStride = 0x9E3779B9u;

// estimate(n_active, pair_ids, offsets[n_active + 1], matches, seeds, results, inliers) -> 0 on success;
// inliers of active pair k are written at offsets[k] (results[k].n_inliers of them).
template <class EstimateBatch>
int estimate_multiple(int64_t n_pairs, const int64_t* match_offsets, const uint32_t* matches, const uint32_t* seeds,
                      bool multiple_ignore_watermark, EstimateBatch&& estimate, b2_two_view_result* results,
                      uint32_t* inlier_matches) {
  std::vector<std::vector<uint32_t>> remaining((size_t)n_pairs), kept_inl((size_t)n_pairs);
  std::vector<std::vector<b2_two_view_result>> kept((size_t)n_pairs);
  std::vector<int64_t> active;
  for (int64_t p = 0; p < n_pairs; ++p) {
    remaining[p].assign(matches + 2 * match_offsets[p], matches + 2 * match_offsets[p + 1]);
    active.push_back(p);
  }
  for (uint32_t round = 0; !active.empty(); ++round) {
    const int64_t na = (int64_t)active.size();
    std::vector<int64_t> off((size_t)na + 1, 0);
    for (int64_t k = 0; k < na; ++k) off[k + 1] = off[k] + (int64_t)remaining[active[k]].size() / 2;
    std::vector<uint32_t> m((size_t)std::max<int64_t>(off[na], 1) * 2), inl(m.size()), sd((size_t)na);
    for (int64_t k = 0; k < na; ++k) {
      const std::vector<uint32_t>& r = remaining[active[k]];
      if (!r.empty()) memcpy(m.data() + 2 * off[k], r.data(), r.size() * sizeof(uint32_t));
      sd[k] = seeds[active[k]] + round * kRoundSeedStride;
    }
    std::vector<b2_two_view_result> res((size_t)na);
    const int rc = estimate(na, active.data(), off.data(), m.data(), sd.data(), res.data(), inl.data());
    if (rc != 0) return rc;
    std::vector<int64_t> next;
    for (int64_t k = 0; k < na; ++k) {
      const int64_t p = active[k];
      const b2_two_view_result& r = res[k];
      if (r.config == kConfigDegenerate) continue;  // :138-140 break
      const uint32_t* in = inl.data() + 2 * off[k];
      if (!(multiple_ignore_watermark && r.config == kConfigWatermark)) {  // :142-148
        kept[p].push_back(r);
        kept_inl[p].insert(kept_inl[p].end(), in, in + 2 * (size_t)r.n_inliers);
      }
      // ExtractOutlierMatches (:67-88)
      std::set<std::pair<uint32_t, uint32_t>> gone;
      for (int32_t i = 0; i < r.n_inliers; ++i) gone.emplace(in[2 * i], in[2 * i + 1]);
      std::vector<uint32_t> rest;
      rest.reserve(remaining[p].size());
      for (size_t i = 0; i + 1 < remaining[p].size(); i += 2)
        if (!gone.count({remaining[p][i], remaining[p][i + 1]})) {
          rest.push_back(remaining[p][i]);
          rest.push_back(remaining[p][i + 1]);
        }
      // a non-degenerate geometry always has inliers; guard against a non-shrinking list all the same
      if (rest.size() < remaining[p].size()) next.push_back(p);
      remaining[p].swap(rest);
    }
    active.swap(next);
  }
  for (int64_t p = 0; p < n_pairs; ++p) {
    b2_two_view_result out;
    memset(&out, 0, sizeof out);
    if (kept[p].empty()) {
      out.config = kConfigDegenerate;  // :154-155
    } else if (kept[p].size() == 1) {
      out = kept[p][0];                // :156-157
    } else {
      out.config = kConfigMultiple;    // :158-166
      out.n_inliers = (int32_t)(kept_inl[p].size() / 2);
    }
    results[p] = out;
    if (!kept_inl[p].empty() && !kept[p].empty())
      memcpy(inlier_matches + 2 * match_offsets[p], kept_inl[p].data(), kept_inl[p].size() * sizeof(uint32_t));
  }
  return 0;
}

}  // namespace b2
