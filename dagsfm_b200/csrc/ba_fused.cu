// Fused exact-Schur path of the bundle adjuster: the normal equations of one Levenberg-Marquardt iteration built
// WITHOUT staging per-observation Jacobian blocks.  Every kernel recomputes an observation's residual and analytic
// Jacobian (ba_eval.cuh, ~350 FP64 operations) from its 24-byte record + the L2-resident camera / point parameters,
// so an iteration streams 24 B per observation three times instead of writing and re-reading 224 B four times; and the
// Schur complement S -= sum_p W_p V_p^-1 W_p' is accumulated in REGISTERS per window of images instead of one FP64
// atomic per (observation pair, entry).
//
// Reference: BundleAdjuster::Solve -> ceres::Solve with SPARSE_SCHUR / DENSE_SCHUR (src/optim/bundle_adjustment.cc:
// 258-310, residual src/base/cost_functions.h:57-84); Ceres' SchurEliminator is restated, not ported.
//
//   camera_terms_kernel   one block per image over the image-major permutation: U_i = sum Jc'Jc (55 entries), g_c, diag
//                         in registers, written once per image (plain stores for entries the image owns).
//   schur_points_kernel   one thread per variable point: V = sum Jp'Jp + D_p, V^-1 = M M' (3x3 Cholesky), g_p, and per
//                         observation Z_a = (Jc_a' Jp_a) M (10 x 3), u_p = M' g_p -- so that
//                         W_a V^-1 W_b' = Z_a Z_b'  and  W_a V^-1 g_p = Z_a u_p.
//   schur_window_kernel   points are ordered by their lowest image and cut into chunks whose images fit a window of
//                         NLOC slots (host, once per solve: the structure is fixed across LM iterations).  Per chunk the
//                         block Z (NLOC*10 rows x 3*points columns, zero rows for slots a point does not see) lives in
//                         shared memory in panels of 12 points, and S_window -= Z Z' is a register-tiled product: one
//                         thread per 8 x 8 tile of the upper triangle, 64 accumulators, explicit DFMA.  The window is
//                         added to the packed system with one atomic per entry per CHUNK (not per point pair).
//   finish_kernel         LM diagonal, rhs += g_c, gradient max-norm.
//   backsub_kernel        one thread per point: dp = -V^-1 (g_p + sum Jp'(Jc dc)), the model cost change, the candidate.
// Algorithmic HBM bytes per LM iteration: 24 N_obs + 28 N_pts + 88 N_cam reads, 72 N_pts + 64 N_cam +
// 512 nnzb(S) writes; this path adds 2 x 240 B per observation for Z (written by schur_points, read by schur_window).
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "ba_common.cuh"
#include "ba_eval.cuh"
#include "ba_loss.cuh"

namespace b2 {
namespace baf {

constexpr unsigned kFull = 0xffffffffu;
constexpr int NC = 10, KI = 4;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// Residual and Jacobi-scaled (and loss-corrected) Jacobian blocks of observation o at the CURRENT parameters.
// col[k] = reduced-system column of camera-side column k or -1; pc = variable-point index or -1.
template <int LOSS>
__device__ __forceinline__ void eval_obs(const BaDev& P, int64_t o, int p, double loss_scale, double* r, double* Jc,
                                         double* Jp, int* col, int pc) {
  const int i = P.obs_img[o], cm = P.img_cam[i];
  const double2 xy = P.obs_xy[o];
  bak::evaluate<KI>(P.cam_model[cm], P.qvec + 4 * i, P.tvec + 3 * i, P.xyz + 3 * (int64_t)p, P.cam_params + KI * cm, xy.x,
                    xy.y, r, Jc, Jp);
  if (LOSS != 0) {
    double rho0, w;
    bak::loss_eval<LOSS>(loss_scale, r[0] * r[0] + r[1] * r[1], &rho0, &w);
    r[0] *= w;
    r[1] *= w;
#pragma unroll
    for (int k = 0; k < 2 * NC; ++k) Jc[k] *= w;
#pragma unroll
    for (int k = 0; k < 6; ++k) Jp[k] *= w;
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) col[k] = P.pose_col[6 * i + k];
#pragma unroll
  for (int k = 0; k < KI; ++k) col[6 + k] = P.intr_col[KI * cm + k];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const double s = (col[k] >= 0) ? P.scale_c[col[k]] : 0.0;
    Jc[k] *= s;
    Jc[NC + k] *= s;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double s = (pc >= 0) ? P.scale_p[3 * (int64_t)pc + k] : 0.0;
    Jp[k] *= s;
    Jp[3 + k] *= s;
  }
}

// ------------------------------------------------------------------ camera terms (image-major)
constexpr int kImageThreads = 128;
template <int LOSS>
__global__ void __launch_bounds__(kImageThreads) camera_terms_kernel(BaDev P, BaIter I, BaTiles T, double loss_scale) {
  constexpr int NU = NC * (NC + 1) / 2, NV = NU + 2 * NC;
  __shared__ double sh[kImageThreads / 32][NV];
  __shared__ int scol[NC];
  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cm = P.img_cam[i];
  if (tid < NC) scol[tid] = (tid < 6) ? P.pose_col[6 * i + tid] : P.intr_col[KI * cm + (tid - 6)];
  __syncthreads();
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  const int64_t s0 = I.img_start[i], s1 = I.img_start[i + 1];
  for (int64_t s = s0 + tid; s < s1; s += kImageThreads) {
    const int64_t o = I.img_obs[s];
    const int p = P.obs_pt[o];
    double r[2], Jc[2 * NC], Jp[6];
    int col[NC];
    eval_obs<LOSS>(P, o, p, loss_scale, r, Jc, Jp, col, P.pt_col[p]);
    int u = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double a0 = Jc[k], a1 = Jc[NC + k];
#pragma unroll
      for (int l = k; l < NC; ++l) acc[u++] += a0 * Jc[l] + a1 * Jc[NC + l];
      acc[NU + k] += a0 * r[0] + a1 * r[1];
      acc[NU + NC + k] += a0 * a0 + a1 * a1;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) sh[warp][k] = acc[k];
  __syncthreads();
  for (int v = tid; v < NV; v += kImageThreads) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kImageThreads / 32; ++w) t += sh[w][v];
    if (v < NU) {  // entry (k, l), k <= l, of the image's block in row-major upper-triangular order
      int k = 0, rem = v;
      while (rem >= NC - k) { rem -= NC - k; ++k; }
      const int l = k + rem;
      const int ck = scol[k], cl = scol[l];
      if (ck < 0 || cl < 0) continue;
      double* dst = tile_entry(T, min(ck, cl), max(ck, cl));
      if (l < 6) *dst += t;  // pose x pose: only this block touches the entry before the window kernel runs
      else atomicAdd(dst, t);
    } else {
      const int k = (v - NU) % NC;
      const int c = scol[k];
      if (c < 0) continue;
      double* dst = (v - NU < NC) ? P.g_c : P.diag_c;
      if (k < 6) dst[c] += t;
      else atomicAdd(dst + c, t);
    }
  }
}

// ------------------------------------------------------------------ per point: V^-1 = M M', Z_a, u_p
template <int LOSS>
__global__ void __launch_bounds__(128) schur_points_kernel(BaDev P, BaWin W, double radius, double min_diag,
                                                           double max_diag, double loss_scale) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n_pts) return;
  const int pc = P.pt_col[p];
  if (pc < 0) return;  // constant point: camera terms only (camera_terms_kernel)
  const int64_t o0 = P.pt_start[p];
  const int L = (int)(P.pt_start[p + 1] - o0);
  double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  for (int a = 0; a < L; ++a) {
    double r[2], Jc[2 * NC], Jp[6];
    int col[NC];
    eval_obs<LOSS>(P, o0 + a, p, loss_scale, r, Jc, Jp, col, pc);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double j0 = Jp[3 * i], j1 = Jp[3 * i + 1], j2 = Jp[3 * i + 2];
      v[0] += j0 * j0; v[1] += j0 * j1; v[2] += j0 * j2; v[3] += j1 * j1; v[4] += j1 * j2; v[5] += j2 * j2;
      g[0] += j0 * r[i]; g[1] += j1 * r[i]; g[2] += j2 * r[i];
    }
  }
  P.diag_p[3 * (int64_t)pc] = v[0];
  P.diag_p[3 * (int64_t)pc + 1] = v[3];
  P.diag_p[3 * (int64_t)pc + 2] = v[5];
  double gm = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    P.g_p[3 * (int64_t)pc + k] = g[k];
    gm = fmax(gm, fabs(g[k] / P.scale_p[3 * (int64_t)pc + k]));
  }
  atomic_max_nonneg(P.gmax, gm);
  const double V00 = v[0] + fmin(fmax(v[0], min_diag), max_diag) / radius;
  const double V11 = v[3] + fmin(fmax(v[3], min_diag), max_diag) / radius;
  const double V22 = v[5] + fmin(fmax(v[5], min_diag), max_diag) / radius;
  const double V01 = v[1], V02 = v[2], V12 = v[4];
  const double c00 = V11 * V22 - V12 * V12, c01 = V12 * V02 - V01 * V22, c02 = V01 * V12 - V11 * V02;
  const double det = V00 * c00 + V01 * c01 + V02 * c02, id = 1.0 / det;
  double Vi[9];
  Vi[0] = c00 * id; Vi[1] = (V02 * V12 - V01 * V22) * id; Vi[2] = (V01 * V12 - V02 * V11) * id;
  Vi[3] = c01 * id; Vi[4] = (V00 * V22 - V02 * V02) * id; Vi[5] = (V02 * V01 - V00 * V12) * id;
  Vi[6] = c02 * id; Vi[7] = (V01 * V02 - V00 * V12) * id; Vi[8] = (V00 * V11 - V01 * V01) * id;
#pragma unroll
  for (int k = 0; k < 9; ++k) P.Vinv[9 * (int64_t)pc + k] = Vi[k];
  // V^-1 = M M' (lower Cholesky of the symmetric positive definite inverse)
  const double m00 = sqrt(Vi[0]), m10 = Vi[3] / m00, m20 = Vi[6] / m00;
  const double m11 = sqrt(Vi[4] - m10 * m10), m21 = (Vi[7] - m20 * m10) / m11;
  const double m22 = sqrt(Vi[8] - m20 * m20 - m21 * m21);
  W.U[3 * (int64_t)p] = m00 * g[0] + m10 * g[1] + m20 * g[2];
  W.U[3 * (int64_t)p + 1] = m11 * g[1] + m21 * g[2];
  W.U[3 * (int64_t)p + 2] = m22 * g[2];
  for (int a = 0; a < L; ++a) {
    double r[2], Jc[2 * NC], Jp[6];
    int col[NC];
    eval_obs<LOSS>(P, o0 + a, p, loss_scale, r, Jc, Jp, col, pc);
    double* z = W.Z + (o0 + a) * 30;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double w0 = Jc[k] * Jp[0] + Jc[NC + k] * Jp[3];
      const double w1 = Jc[k] * Jp[1] + Jc[NC + k] * Jp[4];
      const double w2 = Jc[k] * Jp[2] + Jc[NC + k] * Jp[5];
      z[3 * k] = w0 * m00 + w1 * m10 + w2 * m20;
      z[3 * k + 1] = w1 * m11 + w2 * m21;
      z[3 * k + 2] = w2 * m22;
    }
  }
}

// ------------------------------------------------------------------ window product S_window -= Z Z'
template <int NLOC>
struct WinCfg {
  static constexpr int N = NLOC * NC;             // local columns
  static constexpr int NT8 = N / 8;               // 8 x 8 register tiles per side
  static constexpr int NTILES = NT8 * (NT8 + 1) / 2;
  static constexpr int THREADS = (NTILES + 31) / 32 * 32;
  static constexpr int KB = 3 * kWinBatch;        // k-panel: 3 columns per point
  // shared-memory row: every group of 8 columns (one register-tile side, 64 B) padded to 10 doubles, so that the 15 / 20
  // groups a warp's tiles read in one LDS.128 spread over all banks (unpadded, groups 0, 2, 4, ... share four banks: the
  // 57 M bank conflicts of the first capture, profiles/r2_ba_fused_ncu_full.txt)
  static constexpr int NP = NT8 * 10;
  __host__ __device__ static constexpr int pad(int e) { return (e >> 3) * 10 + (e & 7); }
};

template <int NLOC>
__global__ void __launch_bounds__(WinCfg<NLOC>::THREADS) schur_window_kernel(BaDev P, BaWin W, BaTiles T) {
  using C = WinCfg<NLOC>;
  constexpr int N = C::N, KB = C::KB;
  extern __shared__ double win_smem[];  // Zt[3 pl + m][pad(slot * 10 + k)] (above the 48 KB static limit for 16 slots)
  double (*Zt)[C::NP] = reinterpret_cast<double (*)[C::NP]>(win_smem);
  __shared__ double sU[KB];
  __shared__ int sCol[N];
  const int tid = threadIdx.x;
  // this thread's register tile (ti <= tj) of the upper triangle
  int ti = 0, tj = 0;
  {
    int rem = tid;
    while (ti < C::NT8 && rem >= C::NT8 - ti) { rem -= C::NT8 - ti; ++ti; }
    tj = ti + rem;
  }
  const bool owner = tid < C::NTILES;
  for (int chunk = blockIdx.x; chunk < W.n_chunks; chunk += gridDim.x) {
    for (int e = tid; e < N; e += C::THREADS) {
      const int slot = e / NC, k = e - NC * slot;
      const int img = W.chunk_img[chunk * NLOC + slot];
      sCol[e] = img < 0 ? -1 : (k < 6 ? P.pose_col[6 * img + k] : P.intr_col[KI * P.img_cam[img] + (k - 6)]);
    }
    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
    double racc = 0.0;
    const int p_lo = W.chunk_pt0[chunk], p_hi = W.chunk_pt0[chunk + 1];
    for (int b0 = p_lo; b0 < p_hi; b0 += kWinBatch) {
      const int nb = min(kWinBatch, p_hi - b0);
      for (int e = tid; e < 3 * nb * C::NP; e += C::THREADS) (&Zt[0][0])[e] = 0.0;
      __syncthreads();
      for (int pl = 0; pl < nb; ++pl) {
        const int p = W.pt_order[b0 + pl];
        const int64_t o0 = P.pt_start[p];
        const int L = (int)(P.pt_start[p + 1] - o0);
        for (int e = tid; e < L * 30; e += C::THREADS) {
          const int a = e / 30, q = e - 30 * a, k = q / 3, m = q - 3 * k;
          Zt[3 * pl + m][C::pad(W.obs_slot[o0 + a] * NC + k)] = W.Z[(o0 + a) * 30 + q];
        }
        if (tid < 3) sU[3 * pl + tid] = W.U[3 * (int64_t)p + tid];
      }
      __syncthreads();
      if (owner) {
        for (int kk = 0; kk < 3 * nb; ++kk) {
          double av[8], bv[8];
          const double2* pa = reinterpret_cast<const double2*>(&Zt[kk][ti * 10]);
          const double2* pb = reinterpret_cast<const double2*>(&Zt[kk][tj * 10]);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const double2 x = pa[h], y = pb[h];
            av[2 * h] = x.x; av[2 * h + 1] = x.y;
            bv[2 * h] = y.x; bv[2 * h + 1] = y.y;
          }
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 8; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
      }
      if (tid < N)
        for (int kk = 0; kk < 3 * nb; ++kk) racc = fma(Zt[kk][C::pad(tid)], sU[kk], racc);
      __syncthreads();
    }
    // ---- the window into the packed system: S -= Z Z' (upper entries), rhs -= Z u
    if (owner) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          const int r = ti * 8 + a, c = tj * 8 + b;
          if (r > c) continue;  // lower half of a diagonal register tile
          const int ca = sCol[r], cb = sCol[c];
          if (ca < 0 || cb < 0) continue;
          double v = acc[a][b];
          if (v == 0.0) continue;
          if (r != c && ca == cb) v = v + v;  // two slots sharing a camera: (r, c) and (c, r) land on one diagonal entry
          atomicAdd(tile_entry(T, min(ca, cb), max(ca, cb)), -v);
        }
      }
    }
    if (tid < N && sCol[tid] >= 0 && racc != 0.0) atomicAdd(P.rhs + sCol[tid], -racc);
    __syncthreads();  // sCol is rebuilt for the next chunk
  }
}

// ------------------------------------------------------------------ LM diagonal, rhs, gradient norm
__global__ void finish_kernel(BaDev P, BaTiles T, double radius, double min_diag, double max_diag) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= (int64_t)T.nt * kST) return;
  double* d = tile_entry(T, (int)j, (int)j);
  if (j >= P.D) {  // padding rows of the last tile: identity
    *d = 1.0;
    return;
  }
  *d += fmin(fmax(P.diag_c[j], min_diag), max_diag) / radius;
  P.rhs[j] += P.g_c[j];
  atomic_max_nonneg(P.gmax, grad_norm_term(P, j));
}

// ------------------------------------------------------------------ back-substitution + model cost + candidate points
// out: [2] |step|^2 of the points, [3] |x|^2 of the points, [4] model cost change
template <int LOSS>
__global__ void __launch_bounds__(128) backsub_kernel(BaDev P, double* out, double loss_scale) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0, st = 0, xs = 0;
  if (p < P.n_pts) {
    const int pc = P.pt_col[p];
    const int64_t o0 = P.pt_start[p];
    const int L = (int)(P.pt_start[p + 1] - o0);
    double dp[3] = {0, 0, 0};
    if (pc >= 0) {
      double s[3] = {P.g_p[3 * (int64_t)pc], P.g_p[3 * (int64_t)pc + 1], P.g_p[3 * (int64_t)pc + 2]};
      for (int a = 0; a < L; ++a) {
        double r[2], Jc[2 * NC], Jp[6];
        int col[NC];
        eval_obs<LOSS>(P, o0 + a, p, loss_scale, r, Jc, Jp, col, pc);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          double jd = 0;
#pragma unroll
          for (int k = 0; k < NC; ++k)
            if (col[k] >= 0) jd += Jc[NC * i + k] * P.dc[col[k]];
#pragma unroll
          for (int k = 0; k < 3; ++k) s[k] += Jp[3 * i + k] * jd;
        }
      }
      const double* Vi = P.Vinv + 9 * (int64_t)pc;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dp[k] = -(Vi[3 * k] * s[0] + Vi[3 * k + 1] * s[1] + Vi[3 * k + 2] * s[2]);
        P.dp[3 * (int64_t)pc + k] = dp[k];
      }
    }
    for (int a = 0; a < L; ++a) {
      double r[2], Jc[2 * NC], Jp[6];
      int col[NC];
      eval_obs<LOSS>(P, o0 + a, p, loss_scale, r, Jc, Jp, col, pc);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double m = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k)
          if (col[k] >= 0) m += Jc[NC * i + k] * P.dc[col[k]];
#pragma unroll
        for (int k = 0; k < 3; ++k) m += Jp[3 * i + k] * dp[k];
        acc -= m * (r[i] + m / 2.0);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double v = 0;
      if (pc >= 0) {
        v = dp[k] * P.scale_p[3 * (int64_t)pc + k];
        st += v * v;
      }
      const double x = P.xyz[3 * (int64_t)p + k];
      P.xyz_new[3 * (int64_t)p + k] = x + v;
      xs += x * x;
    }
  }
  acc = warp_sum(acc);
  st = warp_sum(st);
  xs = warp_sum(xs);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 4, acc);
    atomicAdd(out + 2, st);
    atomicAdd(out + 3, xs);
  }
}

}  // namespace baf

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

cudaError_t baf_launch_camera_terms(const BaDev& P, const BaIter& I, const BaTiles& T, int loss_type, double loss_scale,
                                    cudaStream_t s) {
  if (P.n_obs == 0 || P.n_img == 0) return cudaSuccess;
  if (loss_type == 1) baf::camera_terms_kernel<1><<<P.n_img, baf::kImageThreads, 0, s>>>(P, I, T, loss_scale);
  else if (loss_type == 2) baf::camera_terms_kernel<2><<<P.n_img, baf::kImageThreads, 0, s>>>(P, I, T, loss_scale);
  else baf::camera_terms_kernel<0><<<P.n_img, baf::kImageThreads, 0, s>>>(P, I, T, loss_scale);
  return cudaGetLastError();
}
cudaError_t baf_launch_schur_points(const BaDev& P, const BaWin& W, double radius, double min_diag, double max_diag,
                                    int loss_type, double loss_scale, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  const unsigned g = nblk(P.n_pts, 128);
  if (loss_type == 1) baf::schur_points_kernel<1><<<g, 128, 0, s>>>(P, W, radius, min_diag, max_diag, loss_scale);
  else if (loss_type == 2) baf::schur_points_kernel<2><<<g, 128, 0, s>>>(P, W, radius, min_diag, max_diag, loss_scale);
  else baf::schur_points_kernel<0><<<g, 128, 0, s>>>(P, W, radius, min_diag, max_diag, loss_scale);
  return cudaGetLastError();
}
cudaError_t baf_launch_schur_window(const BaDev& P, const BaWin& W, const BaTiles& T, int n_sm, cudaStream_t s) {
  if (W.n_chunks == 0) return cudaSuccess;
  const int grid = std::min(W.n_chunks, n_sm * 4);
  constexpr int smem12 = baf::WinCfg<12>::KB * baf::WinCfg<12>::NP * (int)sizeof(double);
  constexpr int smem16 = baf::WinCfg<16>::KB * baf::WinCfg<16>::NP * (int)sizeof(double);
  static bool attr_set = false;  // per process; the attribute is a property of the function
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(baf::schur_window_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem12);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(baf::schur_window_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (W.nloc == 12) baf::schur_window_kernel<12><<<grid, baf::WinCfg<12>::THREADS, smem12, s>>>(P, W, T);
  else baf::schur_window_kernel<16><<<grid, baf::WinCfg<16>::THREADS, smem16, s>>>(P, W, T);
  return cudaGetLastError();
}
cudaError_t baf_launch_finish(const BaDev& P, const BaTiles& T, double radius, double min_diag, double max_diag,
                              cudaStream_t s) {
  if (T.nt == 0) return cudaSuccess;
  baf::finish_kernel<<<nblk((int64_t)T.nt * kST, 256), 256, 0, s>>>(P, T, radius, min_diag, max_diag);
  return cudaGetLastError();
}
cudaError_t baf_launch_backsub(const BaDev& P, double* scal, int loss_type, double loss_scale, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  const unsigned g = nblk(P.n_pts, 128);
  if (loss_type == 1) baf::backsub_kernel<1><<<g, 128, 0, s>>>(P, scal, loss_scale);
  else if (loss_type == 2) baf::backsub_kernel<2><<<g, 128, 0, s>>>(P, scal, loss_scale);
  else baf::backsub_kernel<0><<<g, 128, 0, s>>>(P, scal, loss_scale);
  return cudaGetLastError();
}

}  // namespace b2
