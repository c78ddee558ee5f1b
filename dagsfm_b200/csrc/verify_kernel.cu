// Two-view geometric verification on the GPU: one warp owns one image pair and runs
// the reference's whole decision procedure for it,
//
//   TwoViewGeometry::Estimate                    src/estimators/two_view_geometry.cc:113-126
//   EstimateCalibrated / EstimateUncalibrated    :292-425 / :427-489
//   DetectWatermark                              :491-555
//   LORANSAC::Estimate                           src/optim/loransac.h:92-233
//   RandomSampler + Shuffle over std::mt19937    src/optim/random_sampler.cc:43-62, util/random.h:89-129
//   InlierSupportMeasurer                        src/optim/support_measurement.cc:36-60
//   Sampson / transfer residuals                 src/estimators/utils.cc:87-131, homography_matrix.cc:94-131
//
// Batching inside the warp: 32 RANSAC trials at a time.  Lane j solves the minimal
// problem of trial j (FP64, per-lane, see verify_solvers.cuh); then the warp replays the
// trials IN ORDER and scores each hypothesis with all 32 lanes: coalesced 16-byte match
// loads, one residual per lane, __ballot_sync + popc inlier counting.  The sequential
// semantics (strict support comparison, local optimisation on every new best, dynamic
// trial bound, the loop-counter quirk) are replayed exactly; the PRNG draws of trials that
// were sampled but fall after the abort point are pushed back into the stream.
//
// Compiled with --fmad=false so that residuals are the same IEEE operations as the
// reference's scalar C++ code.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/dagsfm_b200.h"
#include "verify_common.cuh"
#include "verify_solvers.cuh"
#include "camera_models.cuh"

namespace b2 {
namespace vf {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = 32 * kWarpsPerBlock;
using LaneView = View<1>;  // per-lane solver workspace: a thread-local array (L1-cached local memory)

// -------------------------------------------------------------------- PRNG
// std::mt19937 (result_type = uint_fast32_t, 32 significant bits).
constexpr int kIdxSmem = 2048;
struct WarpShared {
  double V[81];         // right singular vectors of the warp-level Jacobi SVD
  double R[81];         // triangular factor of the local estimator's constraint matrix
  double sig[9];
  uint32_t mt[624];
  uint32_t ring[256];   // FIFO of raw mt outputs (a batch draws <= 32 x 7 + rejections); rollback = moving r back
  uint16_t idx16[kIdxSmem];  // the sampler's persistent index vector, when the pair has at most kIdxSmem matches
  uint32_t samp[32][8]; // sample indices of the 32 trials of a batch
  uint32_t pos_after[32];
  int cnt[320];         // support count of hypothesis (trial j, model mi) at [10 j + mi]
  uint16_t flat[320];   // hypotheses of the batch in replay order (10 j + mi)
  uint16_t off[34];     // first entry of trial j in flat[]
  uint16_t lo_ids[12];  // 0..9: hypotheses of a local estimate
  int lo_cnt[12];
  int nm[32];
  int lo_nm;
  uint32_t mti, w, r;
  uint32_t r0, overflow;  // FIFO position at the start of the batch; raised if a batch ever needs more than the ring holds
};

static_assert(sizeof(WarpShared) % 8 == 0, "WarpShared must keep doubles aligned");
constexpr size_t kDynSmemBytes = sizeof(WarpShared) * kWarpsPerBlock;

__device__ inline void mt_seed(WarpShared& s, uint32_t seed) {
  s.mt[0] = seed;
  for (int i = 1; i < 624; ++i) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
  s.mti = 624;
  s.w = 0;
  s.r = 0;
  s.r0 = 0;
}
__device__ inline uint32_t mt_next(WarpShared& s) {
  if (s.mti >= 624) {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[(k + 1) % 624] & 0x7fffffffu);
      s.mt[k] = s.mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s.mti = 0;
  }
  uint32_t y = s.mt[s.mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// next raw output through the FIFO (lane 0 only)
__device__ inline uint32_t raw_next(WarpShared& s) {
  if (s.r == s.w) {
    if (s.w - s.r0 >= 256u) s.overflow = 1;  // a rollback could no longer reach the start of the batch
    s.ring[s.w & 255u] = mt_next(s);
    ++s.w;
  }
  return s.ring[(s.r++) & 255u];
}
// libstdc++ (GCC >= 11) std::uniform_int_distribution<uint32_t>(a, b) on a 32-bit URBG:
// Lemire's nearly-divisionless method (bits/uniform_int_dist.h, _S_nd<uint64_t>).
__device__ inline uint32_t uniform_u32(WarpShared& s, uint32_t a, uint32_t b) {
  const uint32_t urange = b - a;
  if (urange == 0xffffffffu) return raw_next(s) + a;
  const uint32_t range = urange + 1;
  uint64_t product = (uint64_t)raw_next(s) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) {
      product = (uint64_t)raw_next(s) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return (uint32_t)(product >> 32) + a;
}

// ---- the same generator, advanced by the whole warp ------------------------------------------------------------------
// std::mt19937's regeneration: new[k] = mt[(k + 397) % 624] ^ twist(mt[k], mt[(k + 1) % 624]) for k = 0 .. 623 IN ORDER.  For
// k < 227 every operand is an old word; from k = 227 on the first operand is the already regenerated word k - 227, which a
// 32-wide step never shares with its own range; the second operand k + 1 is regenerated only by a later step (or by a
// higher lane of this one, hence read-then-barrier-then-write).  k = 623 needs the regenerated words 0 and 396.
__device__ __forceinline__ uint32_t mt_twist_word(uint32_t a, uint32_t b, uint32_t m) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ inline void mt_twist_warp(WarpShared& s, int lane) {
  for (int k0 = 0; k0 < 623; k0 += 32) {
    const int k = k0 + lane;
    uint32_t v = 0;
    if (k < 623) v = mt_twist_word(s.mt[k], s.mt[k + 1], s.mt[k < 227 ? k + 397 : k - 227]);
    __syncwarp();
    if (k < 623) s.mt[k] = v;
    __syncwarp();
  }
  if (lane == 0) {
    s.mt[623] = mt_twist_word(s.mt[623], s.mt[0], s.mt[396]);
    s.mti = 0;
  }
  __syncwarp();
}
// Appends raw outputs to the FIFO until it holds `need` unread ones (need <= 224: a batch of 32 trials of <= 7 draws), 32 per
// step: lane l tempers state word mti + l.  The FIFO never holds more than need + 31 <= 255 words behind r0 = r.
__device__ inline void ring_fill_warp(WarpShared& s, uint32_t need, int lane) {
  for (;;) {
    const uint32_t have = s.w - s.r;   // uniform: read by all lanes after a barrier
    if (have >= need) break;
    if (s.mti >= 624) mt_twist_warp(s, lane);
    const uint32_t mti = s.mti, w = s.w;
    const uint32_t c = min(32u, 624u - mti);
    if ((uint32_t)lane < c) {
      uint32_t y = s.mt[mti + lane];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680u;
      y ^= (y << 15) & 0xefc60000u;
      y ^= (y >> 18);
      s.ring[(w + lane) & 255u] = y;
    }
    __syncwarp();
    if (lane == 0) { s.mti = mti + c; s.w = w + c; }
    __syncwarp();
  }
}

// ----------------------------------------------------------------- residuals
__device__ __forceinline__ double sampson(const double* E, double x1_0, double x1_1, double x2_0, double x2_1) {
  const double Ex1_0 = E[0] * x1_0 + E[1] * x1_1 + E[2];
  const double Ex1_1 = E[3] * x1_0 + E[4] * x1_1 + E[5];
  const double Ex1_2 = E[6] * x1_0 + E[7] * x1_1 + E[8];
  const double Etx2_0 = E[0] * x2_0 + E[3] * x2_1 + E[6];
  const double Etx2_1 = E[1] * x2_0 + E[4] * x2_1 + E[7];
  const double x2tEx1 = x2_0 * Ex1_0 + x2_1 * Ex1_1 + Ex1_2;
  return x2tEx1 * x2tEx1 / (Ex1_0 * Ex1_0 + Ex1_1 * Ex1_1 + Etx2_0 * Etx2_0 + Etx2_1 * Etx2_1);
}
__device__ __forceinline__ double transfer(const double* H, double s_0, double s_1, double d_0, double d_1) {
  const double pd_0 = H[0] * s_0 + H[1] * s_1 + H[2];
  const double pd_1 = H[3] * s_0 + H[4] * s_1 + H[5];
  const double pd_2 = H[6] * s_0 + H[7] * s_1 + H[8];
  const double inv_pd_2 = 1.0 / pd_2;
  const double dd_0 = d_0 - pd_0 * inv_pd_2;
  const double dd_1 = d_1 - pd_1 * inv_pd_2;
  return dd_0 * dd_0 + dd_1 * dd_1;
}
__device__ __forceinline__ double translation_res(const double* T, double a0, double a1, double b0, double b1) {
  const double e0 = (b0 - a0) - T[0], e1 = (b1 - a1) - T[1];
  return e0 * e0 + e1 * e1;
}
template <int TYPE>
__device__ __forceinline__ double residual_t(const double* M, double2 a, double2 b) {
  if (TYPE == EST_H4) return transfer(M, a.x, a.y, b.x, b.y);
  if (TYPE == EST_T2) return translation_res(M, a.x, a.y, b.x, b.y);
  return sampson(M, a.x, a.y, b.x, b.y);
}

// Support counts of up to four hypotheses in one pass over the matches: the counts are pure
// functions of (model, points), so they can be taken ahead of the ordered replay; four independent
// residual chains per lane hide the FP64 divide latency and the points are loaded once.
constexpr int kGroup = 4;

// Counts are kept per lane and reduced once per group: no warp-wide exchange inside the loop over the matches.
// (Division-free variants of the two decisions -- an exact band test for Sampson, a bounded reciprocal for the transfer
// error, both with the reference expression as fallback -- were measured in round 2 and were SLOWER: the bound arithmetic
// costs more FP64 issue slots than the division it removes.  profiles/README.md, session 4.)
template <int TYPE, int G>
__device__ __noinline__ void score_group(const double2* __restrict__ P1, const double2* __restrict__ P2, int M,
                                         const double* __restrict__ models, const uint16_t* ids, int n, double max_res,
                                         int lane, int* cnt_out) {
  double m[G][9];
#pragma unroll
  for (int u = 0; u < G; ++u) {
    const int id = ids[u < n ? u : n - 1];
    const double* src = models + (id / 10) * 90 + (id % 10) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) m[u][k] = src[k];
  }
  int c[G];
#pragma unroll
  for (int u = 0; u < G; ++u) c[u] = 0;
  // the next match's points are in flight while the current one is scored (the scratch copy of a pair's points does
  // not stay in L1: the shared-memory carve-out leaves little of it)
  int i = lane;
  double2 a = P1[i < M ? i : 0], b = P2[i < M ? i : 0];
  for (; i < M; i += 32) {
    const int nx = i + 32 < M ? i + 32 : i;
    const double2 an = P1[nx], bn = P2[nx];
#pragma unroll
    for (int u = 0; u < G; ++u) c[u] += (residual_t<TYPE>(m[u], a, b) <= max_res) ? 1 : 0;
    a = an;
    b = bn;
  }
#pragma unroll
  for (int u = 0; u < G; ++u) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c[u] += __shfl_xor_sync(kFull, c[u], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int u = 0; u < G; ++u)
      if (u < n) cnt_out[ids[u]] = c[u];
  }
  __syncwarp();
}
// E and F share the Sampson residual: one instantiation serves both
template <int TYPE> struct ResidualOf { static constexpr int value = (TYPE == EST_H4 || TYPE == EST_T2) ? TYPE : EST_F7; };

// InlierSupportMeasurer::Evaluate, count only (all lanes return the same value).
template <int TYPE>
__device__ __noinline__ int score_count(const double2* P1, const double2* P2, int M, const double* model, double max_res,
                           int lane) {
  int cnt = 0;
  for (int i0 = 0; i0 < M; i0 += 32) {
    const int i = i0 + lane;
    bool in = false;
    if (i < M) in = residual_t<TYPE>(model, P1[i], P2[i]) <= max_res;
    cnt += __popc(__ballot_sync(kFull, in));
  }
  return cnt;
}
// ... and its residual_sum: a sequential FP64 sum in index order (support_measurement.cc:43-46);
// optionally writes the inlier mask.  Needed only on ties and for a new best.
template <int TYPE>
__device__ __noinline__ double score_sum(const double2* P1, const double2* P2, int M, const double* model, double max_res,
                            int lane, uint8_t* mask_out) {
  double sum = 0;
  for (int i0 = 0; i0 < M; i0 += 32) {
    const int i = i0 + lane;
    bool in = false;
    double r = 0;
    if (i < M) {
      r = residual_t<TYPE>(model, P1[i], P2[i]);
      in = r <= max_res;
      if (mask_out) mask_out[i] = in ? 1 : 0;
    }
    unsigned m = __ballot_sync(kFull, in);
    while (m) {
      const int b = __ffs(m) - 1;
      const double v = __shfl_sync(kFull, r, b);
      sum += v;
      m &= m - 1;
    }
  }
  return sum;
}

// x86-64 GCC semantics of static_cast<size_t>(double) (cvttsd2si based), which the
// reference relies on in ComputeNumTrials when log(denom) == 0 (ransac.h:166).
__device__ inline unsigned long long x86_f64_to_u64(double x) {
  const double two63 = 9223372036854775808.0;
  if (x != x) return 0x8000000000000000ull;
  if (x >= two63) {
    const double y = x - two63;
    if (!(y < two63)) return 0ull;  // cvttsd2si indefinite (0x8000...) xor sign bit
    return (unsigned long long)(long long)y ^ 0x8000000000000000ull;
  }
  if (x <= -two63) return 0x8000000000000000ull;
  return (unsigned long long)(long long)x;
}
// RANSAC::ComputeNumTrials (ransac.h:149-167)
__device__ inline unsigned long long compute_num_trials(unsigned long long num_inliers, unsigned long long num_samples,
                                                        double confidence, int kmin) {
  const double inlier_ratio = (double)num_inliers / (double)num_samples;
  const double nom = 1 - confidence;
  if (nom <= 0) return 0xffffffffffffffffull;
  const double denom = 1 - pow(inlier_ratio, (double)kmin);
  if (denom <= 0) return 1;
  return x86_f64_to_u64(ceil(log(nom) / log(denom)));
}

// ------------------------------------------------------- warp-level Jacobi SVD
// G: rows x 9, column-major (G[c * ld + r]) in global scratch; V (9x9 row-major) in shared.
// One-sided Jacobi; lanes split the rows.  On return V's columns are sorted by descending
// singular value (sig in shared too).
// Sum of nine terms held one per lane on lanes 0..8 (zeros elsewhere) as the 32-lane xor butterfly (offsets 16, 8, 4, 2, 1)
// associates it: (((x0 + x8) + x4) + (x2 + x6)) + ((x1 + x5) + (x3 + x7)); the additions of +0.0 it also performs are exact.
__device__ __forceinline__ double butterfly9(const double* x) {
  return (((x[0] + x[8]) + x[4]) + (x[2] + x[6])) + ((x[1] + x[5]) + (x[3] + x[7]));
}
__device__ __noinline__ void warp_jacobi9(double* G, int rows, int ld, double* V, double* sig, int lane) {
  for (int i = lane; i < 81; i += 32) V[i] = (i / 9 == i % 9) ? 1.0 : 0.0;
  __syncwarp();
  double frob2 = 0;
  for (int c = 0; c < 9; ++c)
    for (int i = lane; i < rows; i += 32) frob2 += G[(size_t)c * ld + i] * G[(size_t)c * ld + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) frob2 += __shfl_xor_sync(kFull, frob2, o);
  const double tiny = frob2 * 1e-40;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 8; ++p) {
      for (int q = p + 1; q < 9; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        double* gp = G + (size_t)p * ld;
        double* gq = G + (size_t)q * ld;
        if (rows <= 9) {
          // The 9 x 9 triangular factor (or a minimal problem): every lane forms the three sums itself from broadcast loads,
          // in the association the xor butterfly below gives nine terms on lanes 0..8 -- bit-identical, without 15 shuffles.
          double xa[9], xb[9], xg[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            double a = 0.0, b = 0.0;
            if (i < rows) { a = gp[i]; b = gq[i]; }
            xa[i] = a * a; xb[i] = b * b; xg[i] = a * b;
          }
          alpha = butterfly9(xa);
          beta = butterfly9(xb);
          gamma = butterfly9(xg);
          __syncwarp();  // every lane has read both columns before the lanes that own rows rotate them
        } else {
          for (int i = lane; i < rows; i += 32) {
            const double a = gp[i], b = gq[i];
            alpha += a * a;
            beta += b * b;
            gamma += a * b;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            alpha += __shfl_xor_sync(kFull, alpha, o);
            beta += __shfl_xor_sync(kFull, beta, o);
            gamma += __shfl_xor_sync(kFull, gamma, o);
          }
        }
        if (gamma == 0.0 || (alpha <= tiny || beta <= tiny) || fabs(gamma) <= kEps * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double s = c * t;
        for (int i = lane; i < rows; i += 32) {
          const double a = gp[i], b = gq[i];
          gp[i] = c * a - s * b;
          gq[i] = s * a + c * b;
        }
        if (lane < 9) {
          const double a = V[lane * 9 + p], b = V[lane * 9 + q];
          V[lane * 9 + p] = c * a - s * b;
          V[lane * 9 + q] = s * a + c * b;
        }
        __syncwarp();
      }
    }
    if (!rotated) break;
  }
  // column norms -> order (every lane computes the same permutation)
  double nrm[9];
  for (int j = 0; j < 9; ++j) {
    double s = 0;
    const double* g = G + (size_t)j * ld;
    if (rows <= 9) {
      double x[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const double a = i < rows ? g[i] : 0.0;
        x[i] = a * a;
      }
      s = butterfly9(x);
    } else {
      for (int i = lane; i < rows; i += 32) s += g[i] * g[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    }
    nrm[j] = sqrt(s);
  }
  int order[9];
  for (int j = 0; j < 9; ++j) {
    int rank = 0;
    for (int i = 0; i < 9; ++i) rank += (nrm[i] > nrm[j] || (nrm[i] == nrm[j] && i < j)) ? 1 : 0;
    order[rank] = j;
  }
  double row[9];
  if (lane < 9)
    for (int j = 0; j < 9; ++j) row[j] = V[lane * 9 + order[j]];
  __syncwarp();
  if (lane < 9) {
    for (int j = 0; j < 9; ++j) V[lane * 9 + j] = row[j];
    sig[lane] = nrm[order[lane]];
  }
  __syncwarp();
}

// Householder QR of G (rows x 9, column-major in global scratch, destroyed) by the warp: R (9 x 9
// upper triangular, column-major like G) goes to shared memory.  A and R share singular values and right
// singular vectors, so the Jacobi sweeps then run on 81 shared doubles instead of streaming the
// rows x 9 matrix through L2 36 times per sweep (oracle: householder_r; Eigen::JacobiSVD
// preconditions tall input with a QR too).  rows > 9.
__device__ __noinline__ void warp_qr9(double* G, int rows, int ld, double* R, int lane) {
  for (int i = lane; i < 81; i += 32) R[i] = 0.0;
  __syncwarp();
  for (int k = 0; k < 9; ++k) {
    double* gk = G + (size_t)k * ld;
    const int nj = 8 - k;  // columns right of k
    // one pass over the rows below k: |v|^2 and v . g_j for every remaining column (independent loads)
    double s = 0, w[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) w[jj] = 0;
    for (int r = lane; r < rows; r += 32) {
      if (r <= k) continue;
      const double a = gk[r];
      s += a * a;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (jj < nj) w[jj] += a * gk[(size_t)(jj + 1) * ld + r];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(kFull, s, o);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) w[jj] += __shfl_xor_sync(kFull, w[jj], o);
    }
    const double x0 = gk[k];
    const double nrm = sqrt(x0 * x0 + s);
    if (nrm == 0.0) {
      if (lane == 0)
        for (int j = k + 1; j < 9; ++j) R[j * 9 + k] = G[(size_t)j * ld + k];
      __syncwarp();
      continue;
    }
    const double v0 = x0 + (x0 >= 0 ? nrm : -nrm);
    const double beta = 2.0 / (v0 * v0 + s);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      if (jj < nj) {
        const double gkj = gk[(size_t)(jj + 1) * ld + k];
        w[jj] = (w[jj] + v0 * gkj) * beta;
        if (lane == 0) R[(k + 1 + jj) * 9 + k] = gkj - w[jj] * v0;
      }
    }
    for (int r = lane; r < rows; r += 32) {
      if (r <= k) continue;
      const double a = gk[r];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (jj < nj) gk[(size_t)(jj + 1) * ld + r] -= w[jj] * a;
    }
    if (lane == 0) R[k * 9 + k] = (x0 >= 0 ? -nrm : nrm);
    __syncwarp();
  }
}
// Right singular vectors (descending) of the rows x 9 constraint matrix of a local estimator.
__device__ void warp_svd9(double* G, int rows, int ld, WarpShared& sh, int lane) {
  if (rows > 9) {
    warp_qr9(G, rows, ld, sh.R, lane);
    warp_jacobi9(sh.R, 9, 9, sh.V, sh.sig, lane);
  } else {
    warp_jacobi9(G, rows, ld, sh.V, sh.sig, lane);
  }
}

// ---------------------------------------------------------------- local estimators
// Hartley statistics of the inlier points (index list inl[0..N)) -- warp-parallel sums.
__device__ void warp_hartley(const double2* P, const uint32_t* inl, int N, int lane, double* T) {
  double cx = 0, cy = 0;
  for (int k = lane; k < N; k += 32) {
    const double2 p = P[inl[k]];
    cx += p.x;
    cy += p.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cx += __shfl_xor_sync(kFull, cx, o);
    cy += __shfl_xor_sync(kFull, cy, o);
  }
  cx /= N;
  cy /= N;
  double rms = 0;
  for (int k = lane; k < N; k += 32) {
    const double2 p = P[inl[k]];
    const double dx = p.x - cx, dy = p.y - cy;
    rms += dx * dx + dy * dy;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rms += __shfl_xor_sync(kFull, rms, o);
  rms = sqrt(rms / N);
  const double nf = sqrt(2.0) / rms;
  T[0] = nf; T[1] = 0; T[2] = -nf * cx; T[3] = 0; T[4] = nf; T[5] = -nf * cy; T[6] = 0; T[7] = 0; T[8] = 1;
}
__device__ __forceinline__ double2 apply_T(const double* T, double2 p) {
  const double n0 = T[0] * p.x + T[1] * p.y + T[2];
  const double n1 = T[3] * p.x + T[4] * p.y + T[5];
  const double n2 = T[6] * p.x + T[7] * p.y + T[8];
  const double inv = 1.0 / n2;
  return make_double2(n0 * inv, n1 * inv);
}

// Local estimator on the N inliers listed in inl[]; writes models (<= 10 x 9) to `models`
// (global scratch, visible to all lanes) and returns the count (uniform).
template <int type>
__device__ __noinline__ int local_estimate(const double2* P1, const double2* P2, const uint32_t* inl, int N, double* G,
                              int ld, WarpShared& sh, double* sig_sh, double* models, int lane) {
  if (type == EST_T2) {  // translation_transform.h:84-102: mean_dst - mean_src, sums in index order
    if (lane == 0) {
      double sx = 0, sy = 0, dx = 0, dy = 0;
      for (int k = 0; k < N; ++k) {
        const double2 a = P1[inl[k]], b = P2[inl[k]];
        sx += a.x; sy += a.y; dx += b.x; dy += b.y;
      }
      sx /= N; sy /= N; dx /= N; dy /= N;
      for (int k = 0; k < 9; ++k) models[k] = 0.0;
      models[0] = dx - sx;
      models[1] = dy - sy;
    }
    __syncwarp();
    return 1;
  }
  int nm = 0;
  if (type == EST_E5) {
    for (int k = lane; k < N; k += 32) {
      const double2 a = P1[inl[k]], b = P2[inl[k]];
      G[0 * (size_t)ld + k] = a.x * b.x; G[1 * (size_t)ld + k] = a.y * b.x; G[2 * (size_t)ld + k] = b.x;
      G[3 * (size_t)ld + k] = a.x * b.y; G[4 * (size_t)ld + k] = a.y * b.y; G[5 * (size_t)ld + k] = b.y;
      G[6 * (size_t)ld + k] = a.x; G[7 * (size_t)ld + k] = a.y; G[8 * (size_t)ld + k] = 1;
    }
    __syncwarp();
    warp_svd9(G, N, ld, sh, lane);
    if (lane == 0) {
      double Eb[36];
      for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 9; ++i) Eb[9 * k + i] = sh.V[i * 9 + 5 + k];
      sh.lo_nm = solve_e5_from_basis(Eb, models);
    }
    __syncwarp();
    nm = sh.lo_nm;
    __syncwarp();
    return nm;
  }
  double T1[9], T2[9];
  warp_hartley(P1, inl, N, lane, T1);
  warp_hartley(P2, inl, N, lane, T2);
  int rows;
  if (type == EST_F7) {  // 8-point, fundamental_matrix.cc:150-192
    rows = N;
    for (int k = lane; k < N; k += 32) {
      const double2 a = apply_T(T1, P1[inl[k]]), b = apply_T(T2, P2[inl[k]]);
      G[0 * (size_t)ld + k] = a.x * b.x; G[1 * (size_t)ld + k] = a.y * b.x; G[2 * (size_t)ld + k] = 1.0 * b.x;
      G[3 * (size_t)ld + k] = a.x * b.y; G[4 * (size_t)ld + k] = a.y * b.y; G[5 * (size_t)ld + k] = 1.0 * b.y;
      G[6 * (size_t)ld + k] = a.x; G[7 * (size_t)ld + k] = a.y; G[8 * (size_t)ld + k] = 1.0;
    }
  } else {  // homography DLT, homography_matrix.cc:44-92: rows [0,N) and [N,2N)
    rows = 2 * N;
    for (int k = lane; k < N; k += 32) {
      const double2 s = apply_T(T1, P1[inl[k]]), d = apply_T(T2, P2[inl[k]]);
      const int j = N + k;
      G[0 * (size_t)ld + k] = -s.x; G[1 * (size_t)ld + k] = -s.y; G[2 * (size_t)ld + k] = -1;
      G[3 * (size_t)ld + k] = 0; G[4 * (size_t)ld + k] = 0; G[5 * (size_t)ld + k] = 0;
      G[6 * (size_t)ld + k] = s.x * d.x; G[7 * (size_t)ld + k] = s.y * d.x; G[8 * (size_t)ld + k] = d.x;
      G[0 * (size_t)ld + j] = 0; G[1 * (size_t)ld + j] = 0; G[2 * (size_t)ld + j] = 0;
      G[3 * (size_t)ld + j] = -s.x; G[4 * (size_t)ld + j] = -s.y; G[5 * (size_t)ld + j] = -1;
      G[6 * (size_t)ld + j] = s.x * d.y; G[7 * (size_t)ld + j] = s.y * d.y; G[8 * (size_t)ld + j] = d.y;
    }
  }
  __syncwarp();
  warp_svd9(G, rows, ld, sh, lane);
  if (lane == 0) {
    double nv[9];
    for (int k = 0; k < 9; ++k) nv[k] = sh.V[k * 9 + 8];
    if (type == EST_F7) finish_f8(nv, T1, T2, models);
    else finish_h(nv, T1, T2, models);
  }
  __syncwarp();
  return 1;
}

struct RansacResult {
  bool success;
  int num_inliers;
  double residual_sum;
  long long num_trials;
  double model[9];
};

struct Scratch {
  double2 *px1, *px2;              // matched points of the stage (normalised for E, pixel for F / H)   [Mcap]
  double2 *ip1, *ip2;              // inlier pixel points (watermark)       [Mcap]
  uint32_t* idx;                   // sampler's persistent index vector     [Mcap]
  uint32_t* inl;                   // inlier index list                     [Mcap]
  double* G;                       // 9 columns x ld (ld = 2*Mcap)
  int ld;
  double* models;                  // [32][10][9] per-lane sample models
  double* lomodels;                // [10][9] local models
  uint8_t* tmask;                  // scratch mask
  unsigned long long* prof;        // optional cycle counters
};

// LORANSAC::Estimate for one estimator over the matched points (P1,P2)[0..M).
template <int type, int GS>
__device__ __noinline__ void ransac_warp(const double2* P1, const double2* P2, int M, double max_error,
                            double min_inlier_ratio, double confidence, long long min_num_trials,
                            long long max_num_trials_opt, WarpShared& sh, double* sig_sh, const Scratch& sc,
                            uint8_t* mask_out, RansacResult* out, int lane, LaneView ws) {
  const int kmin = min_samples(type), klo = local_min_samples(type);
  out->success = false;
  out->num_trials = 0;
  out->num_inliers = 0;
  out->residual_sum = DBL_MAX;
  for (int k = 0; k < 9; ++k) out->model[k] = 0.0;
  // RANSAC ctor (ransac.h:135-147)
  unsigned long long max_trials = (unsigned long long)max_num_trials_opt;
  {
    const unsigned long long dyn = compute_num_trials((unsigned long long)(min_inlier_ratio * 100000.0), 100000ull, confidence, kmin);
    if (dyn < max_trials) max_trials = dyn;
  }
  if (M < kmin) return;
  const double max_residual = max_error * max_error;
  int best_count = 0;
  double best_sum = DBL_MAX;
  double best_model[9];
  for (int k = 0; k < 9; ++k) best_model[k] = 0.0;
  bool abort = false;
  if (lane < 12) sh.lo_ids[lane] = (uint16_t)lane;
  // sampler.Initialize.  The shuffle below is a chain of dependent swaps by one lane: in shared memory when the pair's
  // matches fit (a swap through the global scratch costs two L2 round trips), in the scratch otherwise
  const bool idx_sm = M <= kIdxSmem;
  if (idx_sm) for (int i = lane; i < M; i += 32) sh.idx16[i] = (uint16_t)i;
  else for (int i = lane; i < M; i += 32) sc.idx[i] = (uint32_t)i;
  __syncwarp();
  unsigned long long dyn_max = max_trials;
  unsigned long long t0 = 0;
  unsigned long long num_trials = 0;
  bool ended = false;
  while (t0 < max_trials) {
    if (abort) {  // abort was raised by the last trial of the previous batch (loransac.h:131-134)
      num_trials = t0 + 1;
      ended = true;
      break;
    }
    const int nb = (int)((max_trials - t0) < 32ull ? (max_trials - t0) : 32ull);
    const long long c0 = clock64();
    // --- sample indices for trials t0 .. t0+nb-1 (Shuffle of the persistent vector).  The generator runs ahead by the whole
    // warp (regeneration and tempering 32 words at a time), every lane turns raw outputs into swap targets with Lemire's
    // multiply -- draw d of the batch reads output r + d, which holds as long as no draw is rejected (probability
    // range / 2^32 per draw) -- and lane 0 is left with the chain of swaps.  A batch with a possible rejection (low < range
    // for any draw) is sampled by the sequential code instead; both consume the stream exactly as the reference does.
    {
      const uint32_t need = (uint32_t)(nb * kmin);
      ring_fill_warp(sh, need, lane);
      const uint32_t r = sh.r;
      bool suspect = false;
      for (uint32_t d = lane; d < need; d += 32) {
        const uint32_t i = d % (uint32_t)kmin, j = d / (uint32_t)kmin;
        const uint32_t range = (uint32_t)M - i;   // uniform_int_distribution(i, M - 1)
        const uint64_t product = (uint64_t)sh.ring[(r + d) & 255u] * (uint64_t)range;
        suspect |= (uint32_t)product < range;
        sh.samp[j][i] = (uint32_t)(product >> 32) + i;
      }
      const bool fast = !__any_sync(kFull, suspect);
      __syncwarp();
      if (lane == 0) {
        const uint32_t last = (uint32_t)(M - 1);
        sh.r0 = sh.r;
        for (int j = 0; j < nb; ++j) {
          for (uint32_t i = 0; i < (uint32_t)kmin; ++i) {
            const uint32_t jj = fast ? sh.samp[j][i] : uniform_u32(sh, i, last);
            if (idx_sm) {
              const uint16_t a = sh.idx16[i], b = sh.idx16[jj];
              sh.idx16[i] = b;
              sh.idx16[jj] = a;
            } else {
              const uint32_t a = sc.idx[i], b = sc.idx[jj];
              sc.idx[i] = b;
              sc.idx[jj] = a;
            }
          }
          for (int i = 0; i < kmin; ++i) sh.samp[j][i] = idx_sm ? (uint32_t)sh.idx16[i] : sc.idx[i];
          if (fast) sh.r = r + (uint32_t)((j + 1) * kmin);
          sh.pos_after[j] = sh.r;
        }
      }
    }
    __syncwarp();
    const long long c1 = clock64();
    // --- lane j: minimal solver of trial t0 + j
    {
      int nm = 0;
      double* mymodels = sc.models + (size_t)lane * 90;
      if (lane < nb) {
        double a[14], b[14];
        for (int i = 0; i < kmin; ++i) {
          const double2 pa = P1[sh.samp[lane][i]], pb = P2[sh.samp[lane][i]];
          a[2 * i] = pa.x; a[2 * i + 1] = pa.y; b[2 * i] = pb.x; b[2 * i + 1] = pb.y;
        }
        double mm[90];
        if (type == EST_E5) nm = solve_e5(ws, a, b, mm);
        else if (type == EST_F7) nm = solve_f7(ws, a, b, mm);
        else if (type == EST_H4) nm = solve_h4(ws, a, b, mm);
        else {  // translation from one sample: mean_dst - mean_src with n = 1
          for (int k = 0; k < 9; ++k) mm[k] = 0.0;
          mm[0] = b[0] / 1.0 - a[0] / 1.0;
          mm[1] = b[1] / 1.0 - a[1] / 1.0;
          nm = 1;
        }
        for (int k = 0; k < 9 * nm; ++k) mymodels[k] = mm[k];
      }
      sh.nm[lane] = nm;
      // hypotheses of the batch in replay order
      int incl = nm;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, d);
        if (lane >= d) incl += t;
      }
      const int first = incl - nm;
      sh.off[lane] = (uint16_t)first;
      if (lane == 31) sh.off[32] = (uint16_t)incl;
      for (int mi = 0; mi < nm; ++mi) sh.flat[first + mi] = (uint16_t)(lane * 10 + mi);
    }
    __syncwarp();
    const long long c2 = clock64();
    long long lo_cycles = 0;
    // --- ordered replay
    for (int j = 0; j < nb; ++j) {
      const unsigned long long trial = t0 + j;
      if (abort) {  // loransac.h:131-134: the for-increment already happened, then += 1
        num_trials = trial + 1;
        ended = true;
        if (lane == 0) sh.r = sh.pos_after[j - 1];  // un-draw trials >= j (j >= 1 here)
        break;
      }
      if ((j & 7) == 0) {  // support counts of the next eight trials' hypotheses
        const int g1 = sh.off[min(j + 8, 32)];
        for (int g = sh.off[j]; g < g1; g += GS)
          score_group<ResidualOf<type>::value, GS>(P1, P2, M, sc.models, sh.flat + g, min(GS, g1 - g), max_residual, lane, sh.cnt);
      }
      const int nm = sh.nm[j];
      for (int mi = 0; mi < nm; ++mi) {
        const int cnt = sh.cnt[j * 10 + mi];
        bool better = cnt > best_count;
        double sum = 0;
        double model[9];
        if (cnt >= best_count) {  // the hypothesis itself is only needed from here on (most trials stop at the count)
          for (int k = 0; k < 9; ++k) model[k] = sc.models[(size_t)j * 90 + 9 * mi + k];
          sum = score_sum<ResidualOf<type>::value>(P1, P2, M, model, max_residual, lane, sc.tmask);
          sum = __shfl_sync(kFull, sum, 0);
          if (cnt == best_count) better = sum < best_sum;
        }
        if (better) {
          best_count = cnt;
          best_sum = sum;
          for (int k = 0; k < 9; ++k) best_model[k] = model[k];
          if (cnt > kmin && cnt >= klo) {
            // inlier list of the sample model (mask just written by score_sum)
            __syncwarp();
            int N = 0;
            for (int i0 = 0; i0 < M; i0 += 32) {
              const int i = i0 + lane;
              const bool in = (i < M) && sc.tmask[i];
              const unsigned bm = __ballot_sync(kFull, in);
              if (in) sc.inl[N + __popc(bm & ((1u << lane) - 1))] = (uint32_t)i;
              N += __popc(bm);
            }
            __syncwarp();
            const long long cl0 = clock64();
            const int nlm = local_estimate<type>(P1, P2, sc.inl, N, sc.G, sc.ld, sh, sig_sh, sc.lomodels, lane);
            lo_cycles += clock64() - cl0;
            for (int g = 0; g < nlm; g += GS)
              score_group<ResidualOf<type>::value, GS>(P1, P2, M, sc.lomodels, sh.lo_ids + g, min(GS, nlm - g), max_residual, lane,
                                                       sh.lo_cnt);
            for (int li = 0; li < nlm; ++li) {
              double lm[9];
              for (int k = 0; k < 9; ++k) lm[k] = sc.lomodels[9 * li + k];
              const int lc = sh.lo_cnt[li];
              bool lbetter = lc > best_count;
              double lsum = 0;
              if (lc >= best_count) {
                lsum = score_sum<ResidualOf<type>::value>(P1, P2, M, lm, max_residual, lane, nullptr);
                lsum = __shfl_sync(kFull, lsum, 0);
                if (lc == best_count) lbetter = lsum < best_sum;
              }
              if (lbetter) {
                best_count = lc;
                best_sum = lsum;
                for (int k = 0; k < 9; ++k) best_model[k] = lm[k];
              }
            }
          }
          dyn_max = compute_num_trials((unsigned long long)best_count, (unsigned long long)M, confidence, kmin);
        }
        if (trial >= dyn_max && trial >= (unsigned long long)min_num_trials) {
          abort = true;
          break;
        }
      }
    }
    if (sc.prof && lane == 0) {
      const long long c3 = clock64();
      atomicAdd(sc.prof + 0, (unsigned long long)(c1 - c0));
      atomicAdd(sc.prof + 1, (unsigned long long)(c2 - c1));
      atomicAdd(sc.prof + 2, (unsigned long long)(c3 - c2 - lo_cycles));
      atomicAdd(sc.prof + 3, (unsigned long long)lo_cycles);
      atomicAdd(sc.prof + 4 + (type == EST_E5 ? 0 : type == EST_F7 ? 1 : type == EST_H4 ? 2 : 3), (unsigned long long)(c3 - c0));
    }
    if (ended) break;
    t0 += nb;
    __syncwarp();
  }
  if (!ended) num_trials = max_trials;
  out->num_trials = (long long)num_trials;
  out->num_inliers = best_count;
  out->residual_sum = best_sum;
  for (int k = 0; k < 9; ++k) out->model[k] = best_model[k];
  if (best_count < kmin) return;
  out->success = true;
  score_sum<ResidualOf<type>::value>(P1, P2, M, best_model, max_residual, lane, mask_out);
  __syncwarp();
}

// ------------------------------------------------------------------ cameras
__device__ inline void simple_radial_distortion(double k, double u, double v, double* du, double* dv) {
  const double u2 = u * u, v2 = v * v, r2 = u2 + v2, radial = k * r2;
  *du = u * radial;
  *dv = v * radial;
}
// Camera::ImageToWorld (camera_models.h:734-746) incl. IterativeUndistortion (:547-590)
__device__ inline double2 image_to_world(const b2_camera& c, double2 p) {
  double u, v;
  if (c.model == 0) {
    u = (p.x - c.params[1]) / c.params[0];
    v = (p.y - c.params[2]) / c.params[0];
  } else if (c.model == 1) {
    u = (p.x - c.params[2]) / c.params[0];
    v = (p.y - c.params[3]) / c.params[1];
  } else if (c.model != 2) {  // RADIAL ... THIN_PRISM_FISHEYE: the general table (camera_models.cuh)
    cam::image_to_world(c.model, c.params, p.x, p.y, &u, &v);
  } else {
    u = (p.x - c.params[1]) / c.params[0];
    v = (p.y - c.params[2]) / c.params[0];
    const double k = c.params[3];
    const double x0_0 = u, x0_1 = v;
    double x_0 = u, x_1 = v;
    for (int i = 0; i < 100; ++i) {
      const double step0 = fmax(DBL_EPSILON, fabs(1e-6 * x_0));
      const double step1 = fmax(DBL_EPSILON, fabs(1e-6 * x_1));
      double dx0, dx1, b00, b01, f00, f01, b10, b11, f10, f11;
      simple_radial_distortion(k, x_0, x_1, &dx0, &dx1);
      simple_radial_distortion(k, x_0 - step0, x_1, &b00, &b01);
      simple_radial_distortion(k, x_0 + step0, x_1, &f00, &f01);
      simple_radial_distortion(k, x_0, x_1 - step1, &b10, &b11);
      simple_radial_distortion(k, x_0, x_1 + step1, &f10, &f11);
      const double J00 = 1 + (f00 - b00) / (2 * step0);
      const double J01 = (f10 - b10) / (2 * step1);
      const double J10 = (f01 - b01) / (2 * step0);
      const double J11 = 1 + (f11 - b11) / (2 * step1);
      const double invdet = 1.0 / (J00 * J11 - J01 * J10);
      const double r0 = x_0 + dx0 - x0_0, r1 = x_1 + dx1 - x0_1;
      const double s0 = (J11 * invdet) * r0 + (-J01 * invdet) * r1;
      const double s1 = (-J10 * invdet) * r0 + (J00 * invdet) * r1;
      x_0 -= s0;
      x_1 -= s1;
      if (s0 * s0 + s1 * s1 < 1e-10) break;
    }
    u = x_0;
    v = x_1;
  }
  return make_double2(u, v);
}
__device__ inline double image_to_world_threshold(const b2_camera& c, double thr) {
  // b2_verify_set_images stores every two-focal-length model (PINHOLE, OPENCV, OPENCV_FISHEYE, FULL_OPENCV, FOV,
  // THIN_PRISM_FISHEYE: fx, fy at params[0], params[1]) with model id 1 in this array, see verify_api.cu
  const double mf = (c.model == 1) ? (c.params[0] + c.params[1]) / 2 : c.params[0];
  return thr / mf;
}

__global__ void normalize_points_kernel(const b2_camera* __restrict__ cams, const int64_t* __restrict__ img_off,
                                        int n_images, const double2* __restrict__ xy, double2* __restrict__ nxy,
                                        int64_t n_total) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  int lo = 0, hi = n_images - 1;  // image of keypoint i: last img with img_off <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (img_off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  nxy[i] = image_to_world(cams[lo], xy[i]);
}

__device__ __forceinline__ bool in_box(double2 p, double minx, double maxx, double miny, double maxy) {
  return p.x >= minx && p.x <= maxx && p.y >= miny && p.y <= maxy;
}

// ------------------------------------------------------------------ staged kernels
// TwoViewGeometry::Estimate of a pair is four sequential steps on one PRNG stream -- the E, F and H LORANSACs, then the
// decision with the inlier extraction and the watermark test -- and each step is its own kernel over all pairs of the
// call: STAGE 0 seeds the pair's mt19937 and runs the E estimator (calibrated pairs), 1 = F, 2 = H, 3 = decision.  The
// pair's PRNG state (mt vector, FIFO of drawn-but-returned outputs), the three RANSAC reports and the three inlier masks
// travel between the kernels through HBM (VerifyPairState, ~3.8 KB per pair).  Per pair the operations and their order
// are those of the single-kernel formulation (round 1), so the results are bit-identical; what changes is that every
// warp on the chip runs the SAME estimator at any time: the instruction footprint of a launch is one solver + one
// residual instead of all of them (the single kernel spent more issue slots waiting for instructions than for data,
// profiles/r2_verify_*), and each stage gets its own register budget / CTAs per SM (BPS) and hypotheses per scoring pass.
__device__ __forceinline__ void prng_save(const WarpShared& sh, VerifyPairState& st, int lane) {
  for (int i = lane; i < 624; i += 32) st.mt[i] = sh.mt[i];
  for (int i = lane; i < 256; i += 32) st.ring[i] = sh.ring[i];
  if (lane == 0) { st.mti = sh.mti; st.w = sh.w; st.r = sh.r; }
}
__device__ __forceinline__ void prng_load(WarpShared& sh, const VerifyPairState& st, int lane) {
  for (int i = lane; i < 624; i += 32) sh.mt[i] = st.mt[i];
  for (int i = lane; i < 256; i += 32) sh.ring[i] = st.ring[i];
  if (lane == 0) { sh.mti = st.mti; sh.w = st.w; sh.r = st.r; sh.r0 = st.r; }
}
__device__ __forceinline__ void report_store(VerifyRansacReport& d, const RansacResult& r, int lane) {
  if (lane == 0) {
    d.success = r.success ? 1 : 0;
    d.num_inliers = r.num_inliers;
    d.num_trials = r.num_trials;
    for (int k = 0; k < 9; ++k) d.model[k] = r.model[k];
  }
}

__host__ __device__ constexpr int stage_group(int bps) { return bps >= 2 ? 2 : kGroup; }

template <int STAGE, int BPS>
__global__ void __launch_bounds__(kThreads, BPS) verify_stage_kernel(VerifyArgs A) {
  extern __shared__ double warp_sh[];  // WarpShared[kWarpsPerBlock]
  double lane_work[kLaneWorkDoubles];
  const LaneView ws{lane_work};
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  WarpShared& sh = reinterpret_cast<WarpShared*>(warp_sh)[wib];
  double* sig_sh = sh.sig;
  const int worker = blockIdx.x * (int)(blockDim.x >> 5) + wib;
  Scratch sc;
  {
    uint8_t* base = A.scratch + (size_t)worker * A.scratch_stride;
    const size_t mc = (size_t)A.m_cap;
    sc.px1 = (double2*)base; base += mc * 16;
    sc.px2 = (double2*)base; base += mc * 16;
    sc.ip1 = (double2*)base; base += mc * 16;
    sc.ip2 = (double2*)base; base += mc * 16;
    sc.G = (double*)base; base += mc * 2 * 9 * 8;
    sc.ld = (int)(2 * mc);
    sc.models = (double*)base; base += 32 * 90 * 8;
    sc.lomodels = (double*)base; base += 90 * 8;
    sc.idx = (uint32_t*)base; base += mc * 4;
    sc.inl = (uint32_t*)base; base += mc * 4;
    sc.tmask = base;
    sc.prof = A.prof;
  }
  const b2_two_view_options& o = A.opt;
  constexpr int GS = stage_group(BPS);
  if (lane == 0) sh.overflow = 0;
  __syncwarp();
  for (;;) {
    long long p = 0;
    if (lane == 0) p = (long long)atomicAdd(A.work_counter + STAGE, 1ull);
    p = __shfl_sync(kFull, p, 0);
    if (p >= A.n_pairs) break;
    const uint32_t i1 = A.pairs[2 * p], i2 = A.pairs[2 * p + 1];
    const int64_t moff = A.match_off[p];
    const int M = (int)(A.match_off[p + 1] - moff);
    const bool invalid = i1 >= (uint32_t)A.n_images || i2 >= (uint32_t)A.n_images || M > A.m_cap;
    // DEGENERATE before any estimator runs (two_view_geometry.cc:298-301)
    const bool skip = invalid || (unsigned long long)M < (unsigned long long)o.min_num_inliers;
    VerifyPairState& st = A.state[p];
    uint8_t* const masks = A.masks + (moff - A.mask_base);   // [3][mask_stride], this pair's slice at +moff
    if (STAGE < 3) {
      if (skip) continue;
      const b2_camera c1 = A.cams[i1], c2 = A.cams[i2];
      const bool calibrated = c1.has_prior_focal_length && c2.has_prior_focal_length;
      const uint32_t* mt = A.matches + 2 * moff;
      const int64_t o1 = A.img_off[i1], o2 = A.img_off[i2];
      const double2* src = (STAGE == 0) ? A.nxy : A.xy;   // E works on normalised points, F and H on pixels
      if (STAGE == 0) {
        if (lane == 0) mt_seed(sh, A.seeds[p]);
      } else {
        prng_load(sh, st, lane);
      }
      const bool run = STAGE != 0 || calibrated;
      if (run)
        for (int i = lane; i < M; i += 32) {
          sc.px1[i] = src[o1 + mt[2 * i]];
          sc.px2[i] = src[o2 + mt[2 * i + 1]];
        }
      __syncwarp();
      RansacResult R;
      R.success = false; R.num_inliers = 0; R.num_trials = 0; R.residual_sum = 0;
      for (int k = 0; k < 9; ++k) R.model[k] = 0.0;
      if (run) {
        double err = o.max_error;
        if (STAGE == 0) err = (image_to_world_threshold(c1, o.max_error) + image_to_world_threshold(c2, o.max_error)) / 2;
        constexpr int TYPE = STAGE == 0 ? EST_E5 : STAGE == 1 ? EST_F7 : EST_H4;
        ransac_warp<TYPE, GS>(sc.px1, sc.px2, M, err, o.min_inlier_ratio, o.confidence, o.min_num_trials, o.max_num_trials,
                              sh, sig_sh, sc, masks + (size_t)STAGE * A.mask_stride, &R, lane, ws);
      }
      report_store(st.report[STAGE], R, lane);
      __syncwarp();
      prng_save(sh, st, lane);
      if (lane == 0 && sh.overflow) atomicExch(A.err, 2);
      __syncwarp();
      continue;
    }
    // ------------------------------------------------------------- STAGE 3: the decision
    b2_two_view_result res;
    res.config = 0;
    res.n_inliers = 0;
    res.E_num_inliers = res.F_num_inliers = res.H_num_inliers = 0;
    res.E_num_trials = res.F_num_trials = res.H_num_trials = 0;
    for (int k = 0; k < 9; ++k) res.E[k] = res.F[k] = res.H[k] = 0.0;
    if (invalid && lane == 0) atomicExch(A.err, 1);
    if (skip) {
      res.config = 1;
    } else {
      const b2_camera c1 = A.cams[i1], c2 = A.cams[i2];
      const bool calibrated = c1.has_prior_focal_length && c2.has_prior_focal_length;
      const uint32_t* mt = A.matches + 2 * moff;
      const int64_t o1 = A.img_off[i1], o2 = A.img_off[i2];
      const VerifyRansacReport E = st.report[0], F = st.report[1], H = st.report[2];
      for (int k = 0; k < 9; ++k) { res.E[k] = E.model[k]; res.F[k] = F.model[k]; res.H[k] = H.model[k]; }
      res.E_num_inliers = E.num_inliers; res.F_num_inliers = F.num_inliers; res.H_num_inliers = H.num_inliers;
      res.E_num_trials = (int)E.num_trials; res.F_num_trials = (int)F.num_trials; res.H_num_trials = (int)H.num_trials;
      const unsigned long long mni = (unsigned long long)o.min_num_inliers;
      const unsigned long long En = E.num_inliers, Fn = F.num_inliers, Hn = H.num_inliers;
      int best = -1;  // which mask
      unsigned long long num_inliers = 0;
      bool best_valid = true;  // an unsuccessful report has an EMPTY mask
      if (!calibrated) {
        if ((!F.success && !H.success) || (Fn < mni && Hn < mni)) {
          res.config = 1;
        } else {
          const double H_F = (double)Hn / (double)Fn;
          res.config = (H_F > o.max_H_inlier_ratio) ? 6 : 3;
          best = 1;
          num_inliers = Fn;
          best_valid = F.success;
        }
      } else {
        if ((!E.success && !F.success && !H.success) || (En < mni && Fn < mni && Hn < mni)) {
          res.config = 1;
        } else {
          const double E_F = (double)En / (double)Fn;
          const double H_F = (double)Hn / (double)Fn;
          const double H_E = (double)Hn / (double)En;
          if (E.success && E_F > o.min_E_F_inlier_ratio && En >= mni) {
            if (En >= Fn) { num_inliers = En; best = 0; best_valid = E.success; }
            else { num_inliers = Fn; best = 1; best_valid = F.success; }
            if (H_E > o.max_H_inlier_ratio) {
              res.config = 6;
              if (Hn > num_inliers) { num_inliers = Hn; best = 2; best_valid = H.success; }
            } else {
              res.config = 2;
            }
          } else if (F.success && Fn >= mni) {
            num_inliers = Fn; best = 1; best_valid = true;
            if (H_F > o.max_H_inlier_ratio) {
              res.config = 6;
              if (Hn > num_inliers) { num_inliers = Hn; best = 2; best_valid = H.success; }
            } else {
              res.config = 3;
            }
          } else if (H.success && Hn >= mni) {
            num_inliers = Hn; best = 2; best_valid = true;
            res.config = 6;
          } else {
            res.config = 1;
          }
        }
      }
      if (best >= 0 && res.config != 1) {
        // ExtractInlierMatches (two_view_geometry.cc:54-66) in match order
        const uint8_t* mk = masks + (size_t)best * A.mask_stride;
        uint32_t* outm = A.inlier_out + 2 * moff;
        int N = 0, nb_border = 0;
        const double d1 = sqrt((double)(c1.width * c1.width + c1.height * c1.height));
        const double d2 = sqrt((double)(c2.width * c2.width + c2.height * c2.height));
        const double minx1 = o.watermark_border_size * d1, maxx1 = c1.width - minx1, maxy1 = c1.height - minx1;
        const double minx2 = o.watermark_border_size * d2, maxx2 = c2.width - minx2, maxy2 = c2.height - minx2;
        for (int i0 = 0; i0 < M; i0 += 32) {
          const int i = i0 + lane;
          const bool in = best_valid && (i < M) && mk[i];
          const unsigned bm = __ballot_sync(kFull, in);
          bool border = false;
          if (in) {
            const int pos = N + __popc(bm & ((1u << lane) - 1));
            const uint32_t a = mt[2 * i], b = mt[2 * i + 1];
            const double2 q1 = A.xy[o1 + a], q2 = A.xy[o2 + b];
            outm[2 * pos] = a;
            outm[2 * pos + 1] = b;
            sc.ip1[pos] = q1;
            sc.ip2[pos] = q2;
            border = !in_box(q1, minx1, maxx1, minx1, maxy1) && !in_box(q2, minx2, maxx2, minx2, maxy2);
          }
          nb_border += __popc(__ballot_sync(kFull, border));
          N += __popc(bm);
        }
        res.n_inliers = N;
        __syncwarp();
        if (o.detect_watermark) {
          // DetectWatermark (:491-555); num_inliers is the report's count (== N when the mask is valid)
          const double ratio = (double)nb_border / (double)num_inliers;
          if (!(ratio < o.watermark_min_inlier_ratio)) {
            prng_load(sh, st, lane);
            __syncwarp();
            RansacResult T;
            ransac_warp<EST_T2, kGroup>(sc.ip1, sc.ip2, (int)num_inliers, o.max_error, o.watermark_min_inlier_ratio,
                                        o.confidence, o.min_num_trials, o.max_num_trials, sh, sig_sh, sc, sc.tmask, &T, lane, ws);
            const double inlier_ratio = (double)T.num_inliers / (double)num_inliers;
            if (inlier_ratio >= o.watermark_min_inlier_ratio) res.config = 7;
          }
        }
      }
    }
    if (lane == 0) {
      A.results[p] = res;
      if (sh.overflow) atomicExch(A.err, 2);
    }
    __syncwarp();
  }
}


// ----------------------------------------------------------------- test seams
__global__ void score_models_kernel(int type, int n, const double2* P1, const double2* P2, int n_models,
                                    const double* models, double max_res, int* counts, double* sums, uint8_t* masks) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_models) return;
  double model[9];
  for (int k = 0; k < 9; ++k) model[k] = models[9 * w + k];
  int c;
  double s;
  uint8_t* mk = masks + (size_t)w * n;
  if (type == EST_H4) { c = score_count<EST_H4>(P1, P2, n, model, max_res, lane); s = score_sum<EST_H4>(P1, P2, n, model, max_res, lane, mk); }
  else if (type == EST_T2) { c = score_count<EST_T2>(P1, P2, n, model, max_res, lane); s = score_sum<EST_T2>(P1, P2, n, model, max_res, lane, mk); }
  else { c = score_count<EST_F7>(P1, P2, n, model, max_res, lane); s = score_sum<EST_F7>(P1, P2, n, model, max_res, lane, mk); }
  s = __shfl_sync(kFull, s, 0);
  if (lane == 0) {
    counts[w] = c;
    sums[w] = s;
  }
}

__global__ void debug_sample_stream_kernel(uint32_t seed, int total, int k, int n_trials, uint32_t* idx, int* out) {
  __shared__ WarpShared sh;
  if (threadIdx.x == 0) {
    mt_seed(sh, seed);
    for (int i = 0; i < total; ++i) idx[i] = i;
    for (int t = 0; t < n_trials; ++t) {
      for (uint32_t i = 0; i < (uint32_t)k; ++i) {
        const uint32_t j = uniform_u32(sh, i, (uint32_t)(total - 1));
        const uint32_t a = idx[i], b = idx[j];
        idx[i] = b;
        idx[j] = a;
      }
      for (int i = 0; i < k; ++i) out[t * k + i] = (int)idx[i];
    }
  }
}

__global__ void debug_solve_kernel(int type, int n, const double2* P1, const double2* P2, double* G, uint32_t* inl,
                                   double* models, int* n_models) {
  extern __shared__ double warp_sh[];  // same per-lane workspace layout as the production kernel
  double lane_work[kLaneWorkDoubles];
  const LaneView ws{lane_work};
  WarpShared& sh = *reinterpret_cast<WarpShared*>(warp_sh);
  double* sig = sh.sig;
  if (threadIdx.x >= 32) return;        // one warp works; the block size only fixes the stride
  const int lane = threadIdx.x;
  if (type == 3) {  // F 8-point local estimator on all n points
    for (int i = lane; i < n; i += 32) inl[i] = i;
    __syncwarp();
    const int nm = local_estimate<EST_F7>(P1, P2, inl, n, G, 2 * n, sh, sig, models, lane);
    if (lane == 0) *n_models = nm;
    return;
  }
  if (type <= 2 && n > min_samples(type)) {  // local estimators E5 / H on n points
    for (int i = lane; i < n; i += 32) inl[i] = i;
    __syncwarp();
    const int nm = type == EST_E5 ? local_estimate<EST_E5>(P1, P2, inl, n, G, 2 * n, sh, sig, models, lane)
                 : type == EST_F7 ? local_estimate<EST_F7>(P1, P2, inl, n, G, 2 * n, sh, sig, models, lane)
                                  : local_estimate<EST_H4>(P1, P2, inl, n, G, 2 * n, sh, sig, models, lane);
    if (lane == 0) *n_models = nm;
    return;
  }
  {  // minimal solvers: every lane solves the same sample in its own workspace (the
     // production instantiation pattern); lane 0 reports
    double a[14], b[14], mm[90];
    for (int i = 0; i < n; ++i) { a[2 * i] = P1[i].x; a[2 * i + 1] = P1[i].y; b[2 * i] = P2[i].x; b[2 * i + 1] = P2[i].y; }
    int nm = 0;
    if (type == EST_E5) nm = solve_e5(ws, a, b, mm);
    else if (type == EST_F7) nm = solve_f7(ws, a, b, mm);
    else nm = solve_h4(ws, a, b, mm);
    if (lane == 0) {
      for (int k = 0; k < 9 * nm; ++k) models[k] = mm[k];
      *n_models = nm;
    }
  }
}

}  // namespace vf

// ---------------------------------------------------------------- launchers
size_t verify_scratch_stride(int m_cap) {
  const size_t mc = (size_t)m_cap;
  size_t b = mc * 16 * 4 + mc * 2 * 9 * 8 + 32 * 90 * 8 + 90 * 8 + mc * 4 * 2 + mc;
  return (b + 255) / 256 * 256;
}
// CTAs per SM of the E, F and H stages (1 -> 255 registers, 4 hypotheses per scoring pass; 2 -> 128 registers, 2 per pass).
// B2_VERIFY_BPS = three digits selects them for an A/B; the default is the measured-fastest setting.
void verify_stage_shapes(int bps[3]) {
  bps[0] = 1; bps[1] = 1; bps[2] = 1;
  const char* e = getenv("B2_VERIFY_BPS");
  if (e && strlen(e) == 3)
    for (int k = 0; k < 3; ++k) bps[k] = e[k] == '2' ? 2 : 1;
}
int verify_warps_per_block() { return vf::kWarpsPerBlock; }
int verify_blocks_per_sm() {   // the largest stage sizes the per-warp scratch
  int b[3];
  verify_stage_shapes(b);
  return std::max(b[0], std::max(b[1], b[2]));
}

cudaError_t launch_normalize_points(const b2_camera* cams, const int64_t* img_off, int n_images, const double* xy,
                                    double* nxy, int64_t n_total, cudaStream_t s) {
  if (n_total == 0) return cudaSuccess;
  vf::normalize_points_kernel<<<(unsigned)((n_total + 127) / 128), 128, 0, s>>>(
      cams, img_off, n_images, (const double2*)xy, (double2*)nxy, n_total);
  return cudaGetLastError();
}
template <int STAGE, int BPS>
static cudaError_t launch_stage(const VerifyArgs& a, int n_sm, cudaStream_t s) {
  const size_t dyn = vf::kDynSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(vf::verify_stage_kernel<STAGE, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  int blocks = n_sm * BPS;   // resident CTAs only: the pairs are handed out by the stage's work counter
  blocks = (int)std::min<int64_t>(blocks, (a.n_pairs + vf::kWarpsPerBlock - 1) / vf::kWarpsPerBlock);
  blocks = std::min(blocks, a.max_workers / vf::kWarpsPerBlock);
  if (blocks < 1) blocks = 1;
  vf::verify_stage_kernel<STAGE, BPS><<<blocks, vf::kThreads, dyn, s>>>(a);
  return cudaGetLastError();
}
// The four stages of one call, in stream order (work_counter[0..3] and err cleared by the caller).
cudaError_t launch_verify_pairs(const VerifyArgs& a, int n_sm, cudaStream_t s) {
  int bps[3];
  verify_stage_shapes(bps);
  cudaError_t e;
  e = bps[0] == 2 ? launch_stage<0, 2>(a, n_sm, s) : launch_stage<0, 1>(a, n_sm, s);
  if (e != cudaSuccess) return e;
  e = bps[1] == 2 ? launch_stage<1, 2>(a, n_sm, s) : launch_stage<1, 1>(a, n_sm, s);
  if (e != cudaSuccess) return e;
  e = bps[2] == 2 ? launch_stage<2, 2>(a, n_sm, s) : launch_stage<2, 1>(a, n_sm, s);
  if (e != cudaSuccess) return e;
  return launch_stage<3, 1>(a, n_sm, s);
}
cudaError_t launch_score_models(int type, int n, const double* p1, const double* p2, int n_models, const double* models,
                                double max_res, int* counts, double* sums, uint8_t* masks, cudaStream_t s) {
  if (n_models == 0) return cudaSuccess;
  vf::score_models_kernel<<<(n_models + 3) / 4, 128, 0, s>>>(type, n, (const double2*)p1, (const double2*)p2, n_models,
                                                              models, max_res, counts, sums, masks);
  return cudaGetLastError();
}
cudaError_t launch_debug_sample_stream(uint32_t seed, int total, int k, int n_trials, uint32_t* idx, int* out,
                                       cudaStream_t s) {
  vf::debug_sample_stream_kernel<<<1, 32, 0, s>>>(seed, total, k, n_trials, idx, out);
  return cudaGetLastError();
}
cudaError_t launch_debug_solve(int type, int n, const double* p1, const double* p2, double* G, uint32_t* inl,
                               double* models, int* n_models, cudaStream_t s) {
  const size_t dyn = vf::kDynSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(vf::debug_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  vf::debug_solve_kernel<<<1, vf::kThreads, dyn, s>>>(type, n, (const double2*)p1, (const double2*)p2, G, inl, models,
                                                     n_models);
  return cudaGetLastError();
}

}  // namespace b2
