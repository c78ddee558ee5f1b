// Integer forms of the reference's float32 threshold tests (host-only, no CUDA).
//
// FindBestMatchesOneWay (sift.cc:111-162) evaluates, in float32,
//     a(v)   = acos(min(kDistNorm * v, 1))
//     reject   if a(best) >  max_distance
//     reject   if a(best) >= max_ratio * a(second)
// a() is monotone non-increasing in the integer dot v, so both tests are threshold tests on
// integers: accept iff best >= thr_dist and second <= ratio_lim[best].  The tables are built with
// the host's own acosf -- the very function the reference CPU path calls -- so a device decision
// taken from them is identical by construction.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace b2 {

struct HostThresholds {
  std::vector<float> a;        // a(v), v = 0 .. dot_clamp
  int thr_dist = 0;
  std::vector<int> ratio_lim;  // [dot_clamp + 1], -1 = no second-best passes

  // Returns false if the host's acosf is not monotone on the grid (never observed).
  bool build(float ratio, float dist, int dot_clamp) {
    const float kDistNorm = 1.0f / (512.0f * 512.0f);
    if ((int)a.size() != dot_clamp + 1) {
      a.resize(dot_clamp + 1);
      for (int v = 0; v <= dot_clamp; ++v) a[v] = std::acos(std::min(kDistNorm * (float)v, 1.0f));
      for (int v = 1; v <= dot_clamp; ++v)
        if (a[v] > a[v - 1]) return false;
    }
    int td = dot_clamp + 1;
    for (int v = 1; v <= dot_clamp; ++v)
      if (!(a[v] > dist)) { td = v; break; }
    ratio_lim.assign(dot_clamp + 1, -1);
    for (int b = std::max(td, 1); b <= dot_clamp; ++b) {
      // largest s in [0, b] with NOT (a[b] >= ratio * a[s]); pass(s) is monotone (true first)
      int lo = -1, hi = b;  // invariant: pass(lo) (or lo == -1), search in (lo, hi]
      while (lo < hi) {
        const int mid = lo + (hi - lo + 1) / 2;
        const bool pass = !(a[b] >= ratio * a[mid]);
        if (pass) lo = mid; else hi = mid - 1;
      }
      ratio_lim[b] = lo;
    }
    thr_dist = td;
    return true;
  }
};

}  // namespace b2
