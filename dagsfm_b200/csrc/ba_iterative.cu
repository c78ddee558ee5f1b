// ITERATIVE_SCHUR for the bundle adjuster: the kernels behind the conjugate-gradient solve of the
// reduced camera system that BundleAdjuster::Solve selects above 1000 images
// (src/optim/bundle_adjustment.cc:274-284: ITERATIVE_SCHUR + SCHUR_JACOBI, at most
// max_linear_solver_iterations = 100 inner iterations for the final BA,
// src/controllers/distributed_mapper_controller.cpp:529).  The arithmetic is Ceres' (external,
// 1.14): ImplicitSchurComplement, SchurJacobiPreconditioner, ConjugateGradientsSolver.
//
// S is never formed.  With E = point columns and F = camera-side columns of the (Jacobi-scaled)
// Jacobian held per observation in ObsJac,
//     S x = D_c^2 x + F'(F x - E z),   z_p = (E'E + D_p^2)_p^-1 (E'F x)_p
// is two streaming passes over the observations:
//   matvec_point_kernel   point-major (the CSR order of ObsJac), 8 lanes per point:   x -> z_p
//   image_pass_kernel<0>  image-major (one block per image, observations through the permutation
//                         img_obs): out_i = sum_o Jc_o'(Jc_o x_i - Jp_o z_p); the block owns the image's
//                         pose columns, so they are written without atomics; only the (possibly
//                         shared) intrinsics columns are added atomically.
// The same image pass with (r, t_p = V^-1 g_p) in place of (F x, z) gives the right-hand side
// F'(r - E t_p) together with g_c and diag(F'F) (image_pass_kernel<1>).
// Algorithmic HBM bytes per CG iteration: 2 x 224 B per observation (ObsJac read by both passes)
// + 12 B per observation of indices + 24 B per point (z_p write + read); the vectors of length D live in L2.
//
// SCHUR_JACOBI: the block diagonal of S, one block per Ceres parameter block -- rotation (3 local
// columns), variable tvec components, variable intrinsics of a camera -- accumulated per observation
// by precond_kernel and inverted by precond_invert_kernel (Cholesky, as
// BlockRandomAccessDiagonalMatrix::Invert does).
//
// All dot products of the CG loop are reduced in a fixed order (per-block partial sums, summed on the
// host in block order), so a solve is reproducible run to run and identical on every rank of a
// multi-GPU job, which keeps the ranks' control flow in step.
#include <cuda_runtime.h>

#include <cstdint>

#include "ba_common.cuh"

namespace b2 {
namespace bit {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kLanesPerPoint = 8;
constexpr int kImageThreads = 128;

__device__ __forceinline__ double group8_sum(double v) {
  v += __shfl_xor_sync(kFull, v, 4);
  v += __shfl_xor_sync(kFull, v, 2);
  v += __shfl_xor_sync(kFull, v, 1);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// One observation's Jacobian block into registers with 128-bit loads: sizeof(J) is a multiple of 16 (224 / 352 bytes)
// and the array comes from cudaMalloc, so every block is 16-byte aligned -- 14 (22) LDG.128 instead of 28 (44) LDG.64
// on the two streaming passes of a CG iteration, which are bound by these reads.
template <class J>
__device__ __forceinline__ void load_block(const J* __restrict__ src, J* dst) {
  static_assert(sizeof(J) % sizeof(double2) == 0, "Jacobian blocks are loaded as double2");
  const double2* s = reinterpret_cast<const double2*>(src);
  double2* d = reinterpret_cast<double2*>(dst);
#pragma unroll
  for (int k = 0; k < (int)(sizeof(J) / sizeof(double2)); ++k) d[k] = __ldg(s + k);
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// ------------------------------------------------------------------ per LM iteration
// Eight lanes per point: V = E'E + D_p^2, V^-1, g_p, diag_p, t_p = V^-1 g_p, gradient max norm of
// the point columns.  (The exact path computes the same quantities inside schur_kernel.)
template <class J>
__global__ void __launch_bounds__(256)
point_prepare_kernel(BaDev P, BaIter I, double radius, double min_diag, double max_diag) {
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t p = gt / kLanesPerPoint;
  const int sub = threadIdx.x & (kLanesPerPoint - 1);
  const int pc = (p < P.n_pts) ? P.pt_col[p] : -1;
  double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
  if (pc >= 0) {
    const int64_t o0 = P.pt_start[p];
    const int L = (int)(P.pt_start[p + 1] - o0);
    for (int a = sub; a < L; a += kLanesPerPoint) {
      const J& e = jac<J>(P)[o0 + a];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double j0 = e.Jp[3 * i], j1 = e.Jp[3 * i + 1], j2 = e.Jp[3 * i + 2], r = e.r[i];
        v[0] += j0 * j0; v[1] += j0 * j1; v[2] += j0 * j2; v[3] += j1 * j1; v[4] += j1 * j2; v[5] += j2 * j2;
        g[0] += j0 * r; g[1] += j1 * r; g[2] += j2 * r;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = group8_sum(v[k]);
#pragma unroll
  for (int k = 0; k < 3; ++k) g[k] = group8_sum(g[k]);
  if (pc < 0 || sub != 0) return;
  const int64_t c3 = 3 * (int64_t)pc;
  P.diag_p[c3] = v[0]; P.diag_p[c3 + 1] = v[3]; P.diag_p[c3 + 2] = v[5];
  P.g_p[c3] = g[0]; P.g_p[c3 + 1] = g[1]; P.g_p[c3 + 2] = g[2];
  double gm = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) gm = fmax(gm, fabs(g[k] / P.scale_p[c3 + k]));
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
#pragma unroll
  for (int k = 0; k < 3; ++k) I.tp[c3 + k] = Vi[3 * k] * g[0] + Vi[3 * k + 1] * g[1] + Vi[3 * k + 2] * g[2];
}

// One block per image, its observations through img_obs.
//   MODE 0: out  = sum_o Jc_o' (Jc_o x - Jp_o z_p)                        (the F'(F x - E z) part of S x)
//   MODE 1: out  = sum_o Jc_o' (r_o  - Jp_o t_p)  = rhs of the reduced system,
//           out2 = sum_o Jc_o' r_o = g_c,  out3 = diag(F'F)
// Pose columns belong to this block alone (plain stores into zeroed vectors); intrinsics may be
// shared between images and are added atomically.
template <int MODE, class J>
__global__ void __launch_bounds__(kImageThreads)
image_pass_kernel(BaDev P, BaIter I, const double* __restrict__ x, const double* __restrict__ zp,
                  double* __restrict__ out, double* __restrict__ out2, double* __restrict__ out3) {
  constexpr int KI = J::kKI, NC = J::kNC;
  constexpr int NV = (MODE == 0) ? NC : 3 * NC;
  static_assert(NV <= kImageThreads, "one thread per reduced value");
  __shared__ double sh[kImageThreads / 32][NV];
  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cm = P.img_cam[i];
  int col[NC];
  bool any = false;
#pragma unroll
  for (int k = 0; k < 6; ++k) { col[k] = P.pose_col[6 * i + k]; any |= col[k] >= 0; }
#pragma unroll
  for (int k = 0; k < KI; ++k) { col[6 + k] = P.intr_col[KI * cm + k]; any |= col[6 + k] >= 0; }
  if (!any) return;  // uniform over the block
  double xl[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) xl[k] = (MODE == 0 && col[k] >= 0) ? x[col[k]] : 0.0;
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  const int64_t s0 = I.img_start[i], s1 = I.img_start[i + 1];
  for (int64_t s = s0 + tid; s < s1; s += kImageThreads) {
    const int o = I.img_obs[s];
    J e;
    load_block(jac<J>(P) + o, &e);
    const int pc = P.pt_col[P.obs_pt[o]];
    double z0 = 0, z1 = 0, z2 = 0;
    if (pc >= 0) { z0 = zp[3 * (int64_t)pc]; z1 = zp[3 * (int64_t)pc + 1]; z2 = zp[3 * (int64_t)pc + 2]; }
    double u[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      double ua;
      if (MODE == 0) {
        ua = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k) ua += e.Jc[NC * a + k] * xl[k];
      } else {
        ua = e.r[a];
      }
      u[a] = ua - (e.Jp[3 * a] * z0 + e.Jp[3 * a + 1] * z1 + e.Jp[3 * a + 2] * z2);
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double j0 = e.Jc[k], j1 = e.Jc[NC + k];
      acc[k] += j0 * u[0] + j1 * u[1];
      if (MODE == 1) {
        acc[NC + k] += j0 * e.r[0] + j1 * e.r[1];
        acc[2 * NC + k] += j0 * j0 + j1 * j1;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) sh[warp][k] = acc[k];
  __syncthreads();
  if (tid < NV) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kImageThreads / 32; ++w) s += sh[w][tid];
    const int k = tid % NC, which = tid / NC;
    const int c = col[k];
    if (c >= 0) {
      double* dst = (which == 0) ? out : (which == 1) ? out2 : out3;
      if (k < 6) dst[c] = s;
      else atomicAdd(dst + c, s);
    }
  }
}

// lm_c = clamp(diag_c) / radius; gradient max norm over the camera columns
__global__ void cam_diag_kernel(BaDev P, BaIter I, double radius, double min_diag, double max_diag) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= P.D) return;
  I.lm_c[j] = fmin(fmax(P.diag_c[j], min_diag), max_diag) / radius;
  atomic_max_nonneg(P.gmax, grad_norm_term(P, j));
}

// Thread per observation a: the rows of the block diagonal of S that a touches,
//   M_blk += Jc_a' Jc_a |blk  -  W_a V^-1 (sum over the observations b of the same point that share
//   the parameter block: W_b)' |blk
// for its three camera-side parameter blocks (rotation, tvec, intrinsics).  Summed over a this is
// the (blk, blk) block of S = F'F - F'E (E'E)^-1 E'F; the LM diagonal is added by the inversion.
template <class J>
__global__ void __launch_bounds__(128) precond_kernel(BaDev P, BaIter I) {
  constexpr int KI = J::kKI, NC = J::kNC;
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (o >= P.n_obs) return;
  const int i = P.obs_img[o], cm = P.img_cam[i], p = P.obs_pt[o], pc = P.pt_col[p];
  J e;
  load_block(jac<J>(P) + o, &e);
  int col[NC];
#pragma unroll
  for (int k = 0; k < 6; ++k) col[k] = P.pose_col[6 * i + k];
#pragma unroll
  for (int k = 0; k < KI; ++k) col[6 + k] = P.intr_col[KI * cm + k];
  // A[k][l] for k, l inside one group: rotation 0..2, tvec 3..5, intrinsics 6..6+KI-1
  double A[NC][KI];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int g0 = (k < 3) ? 0 : (k < 6) ? 3 : 6, gn = (k < 6) ? 3 : KI;
#pragma unroll
    for (int l = 0; l < KI; ++l)
      A[k][l] = (l < gn) ? e.Jc[k] * e.Jc[g0 + l] + e.Jc[NC + k] * e.Jc[NC + g0 + l] : 0.0;
  }
  if (pc >= 0) {
    double T[NC][3];  // sum of W_b over the observations of this point in the same image (rows 0..5) / camera (6..)
#pragma unroll
    for (int k = 0; k < NC; ++k) T[k][0] = T[k][1] = T[k][2] = 0.0;
    const int64_t o0 = P.pt_start[p], o1 = P.pt_start[p + 1];
    for (int64_t b = o0; b < o1; ++b) {
      const int ib = P.obs_img[b];
      const bool same_img = ib == i, same_cam = same_img || P.img_cam[ib] == cm;
      if (!same_cam) continue;
      J eb;
      load_block(jac<J>(P) + b, &eb);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        if (k < 6 && !same_img) continue;
#pragma unroll
        for (int m = 0; m < 3; ++m) T[k][m] += eb.Jc[k] * eb.Jp[m] + eb.Jc[NC + k] * eb.Jp[3 + m];
      }
    }
    const double* Vi = P.Vinv + 9 * (int64_t)pc;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      double w[3], y[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) w[m] = e.Jc[k] * e.Jp[m] + e.Jc[NC + k] * e.Jp[3 + m];
#pragma unroll
      for (int m = 0; m < 3; ++m) y[m] = w[0] * Vi[m] + w[1] * Vi[3 + m] + w[2] * Vi[6 + m];
      const int g0 = (k < 3) ? 0 : (k < 6) ? 3 : 6, gn = (k < 6) ? 3 : KI;
#pragma unroll
      for (int l = 0; l < KI; ++l)
        if (l < gn) A[k][l] -= y[0] * T[g0 + l][0] + y[1] * T[g0 + l][1] + y[2] * T[g0 + l][2];
    }
  }
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int g0 = (g == 0) ? 0 : (g == 1) ? 3 : 6, gn = (g == 2) ? KI : 3;
    int first = -1;
#pragma unroll
    for (int l = 0; l < KI; ++l)
      if (l < gn && first < 0 && col[g0 + l] >= 0) first = col[g0 + l];  // columns of a block are consecutive
    if (first < 0) continue;
#pragma unroll
    for (int k = 0; k < KI; ++k) {
      if (k >= gn || col[g0 + k] < 0) continue;
#pragma unroll
      for (int l = 0; l < KI; ++l) {
        if (l >= gn || col[g0 + l] < 0) continue;
        atomicAdd(I.M + KI * (int64_t)col[g0 + k] + (col[g0 + l] - first), A[g0 + k][l]);
      }
    }
  }
}

// Thread per parameter block: M_blk + D_c^2 -> its inverse by Cholesky (BlockRandomAccessDiagonalMatrix::Invert).
// A block that is not positive definite raises *flag (the step is then treated as a linear-solver failure).
template <int KI>
__global__ void precond_invert_kernel(BaDev P, BaIter I) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= P.D || I.blk_first[f] != f) return;
  const int n = I.blk_size[f];
  double A[KI][KI], Lc[KI][KI], Inv[KI][KI];
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) { A[r][c] = I.M[KI * (f + r) + c]; Lc[r][c] = 0.0; }
  for (int r = 0; r < n; ++r) A[r][r] += I.lm_c[f + r];
  for (int j = 0; j < n; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= Lc[j][k] * Lc[j][k];
    if (!(d > 0)) { atomicExch(I.flag, 1); return; }
    Lc[j][j] = sqrt(d);
    for (int r = j + 1; r < n; ++r) {
      double v = A[r][j];
      for (int k = 0; k < j; ++k) v -= Lc[r][k] * Lc[j][k];
      Lc[r][j] = v / Lc[j][j];
    }
  }
  for (int c = 0; c < n; ++c) {
    double y[KI];
    for (int r = 0; r < n; ++r) {
      double v = (r == c) ? 1.0 : 0.0;
      for (int k = 0; k < r; ++k) v -= Lc[r][k] * y[k];
      y[r] = v / Lc[r][r];
    }
    for (int r = n - 1; r >= 0; --r) {
      double v = y[r];
      for (int k = r + 1; k < n; ++k) v -= Lc[k][r] * Inv[k][c];
      Inv[r][c] = v / Lc[r][r];
    }
  }
  for (int r = 0; r < n; ++r)
    for (int c = 0; c < n; ++c) I.M[KI * (f + r) + c] = Inv[r][c];
}


// ------------------------------------------------------------------ camera terms, image-major (experimental)
// B2_BA_CAMTERMS=image on the exact path: what camera_terms_kernel adds with ~75 FP64 atomics per observation
// (U = Jc'Jc into S, g_c, diag_c), summed per image first -- one block per image over the same image-major permutation
// the iterative solver uses, each thread keeping the 55 unique entries of its 10 x 10 block plus g and the diagonal in
// registers -- and written once per image: plain stores for the pose-pose entries (this block owns them), one atomic per
// entry that involves the (possibly shared) intrinsics.  The production kernel stays the default until this is timed.
template <class J>
__global__ void __launch_bounds__(kImageThreads) camera_terms_image_kernel(BaDev P, BaIter I) {
  constexpr int KI = J::kKI, NC = J::kNC, NU = NC * (NC + 1) / 2, NV = NU + 2 * NC;
  __shared__ double sh[kImageThreads / 32][NV];
  const int i = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cm = P.img_cam[i];
  __shared__ int scol[NC];
  if (tid < NC) scol[tid] = (tid < 6) ? P.pose_col[6 * i + tid] : P.intr_col[KI * cm + (tid - 6)];
  __syncthreads();
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  const int64_t s0 = I.img_start[i], s1 = I.img_start[i + 1];
  for (int64_t s = s0 + tid; s < s1; s += kImageThreads) {
    J e;
    load_block(jac<J>(P) + I.img_obs[s], &e);
    int u = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double a0 = e.Jc[k], a1 = e.Jc[NC + k];
#pragma unroll
      for (int l = k; l < NC; ++l) acc[u++] += a0 * e.Jc[l] + a1 * e.Jc[NC + l];
      acc[NU + k] += a0 * e.r[0] + a1 * e.r[1];
      acc[NU + NC + k] += a0 * a0 + a1 * a1;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; ++k) sh[warp][k] = acc[k];
  __syncthreads();
  const int64_t D = P.D;
  for (int v = tid; v < NV; v += kImageThreads) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kImageThreads / 32; ++w) t += sh[w][v];
    if (v < NU) {  // entry (k, l), k <= l, of the image's block: index v in row-major upper-triangular order
      int k = 0, rem = v;
      while (rem >= NC - k) { rem -= NC - k; ++k; }
      const int l = k + rem;
      const int ck = scol[k], cl = scol[l];
      if (ck < 0 || cl < 0) continue;
      const int r = min(ck, cl), c = max(ck, cl);
      if (l < 6) P.S[r * D + c] += t;  // pose x pose: only this block touches the entry (the Schur kernel runs afterwards)
      else atomicAdd(P.S + r * D + c, t);
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

// ------------------------------------------------------------------ per CG iteration
// x -> z_p = V^-1 sum_a Jp_a' (Jc_a x)   (eight lanes per point)
template <class J>
__global__ void __launch_bounds__(256)
matvec_point_kernel(BaDev P, const double* __restrict__ x, double* __restrict__ zp) {
  constexpr int KI = J::kKI, NC = J::kNC;
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t p = gt / kLanesPerPoint;
  const int sub = threadIdx.x & (kLanesPerPoint - 1);
  const int pc = (p < P.n_pts) ? P.pt_col[p] : -1;
  double y[3] = {0, 0, 0};
  if (pc >= 0) {
    const int64_t o0 = P.pt_start[p];
    const int L = (int)(P.pt_start[p + 1] - o0);
    for (int a = sub; a < L; a += kLanesPerPoint) {
      J e;
      load_block(jac<J>(P) + (o0 + a), &e);
      const int i = P.obs_img[o0 + a], cm = P.img_cam[i];
      double u0 = 0, u1 = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int c = P.pose_col[6 * i + k];
        if (c >= 0) { const double xv = x[c]; u0 += e.Jc[k] * xv; u1 += e.Jc[NC + k] * xv; }
      }
#pragma unroll
      for (int k = 0; k < KI; ++k) {
        const int c = P.intr_col[KI * cm + k];
        if (c >= 0) { const double xv = x[c]; u0 += e.Jc[6 + k] * xv; u1 += e.Jc[NC + 6 + k] * xv; }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) y[k] += e.Jp[k] * u0 + e.Jp[3 + k] * u1;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) y[k] = group8_sum(y[k]);
  if (pc < 0 || sub != 0) return;
  const double* Vi = P.Vinv + 9 * (int64_t)pc;
#pragma unroll
  for (int k = 0; k < 3; ++k) zp[3 * (int64_t)pc + k] = Vi[3 * k] * y[0] + Vi[3 * k + 1] * y[1] + Vi[3 * k + 2] * y[2];
}

// Sum of `v` over the block in a fixed order -> partial[blockIdx.x]
__device__ __forceinline__ void block_partial(double v, double* __restrict__ partial) {
  __shared__ double ws[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += ws[k];
    partial[blockIdx.x] = s;
  }
  __syncthreads();
}

// partial = sum a[j] * b[j]
__global__ void __launch_bounds__(256) cg_dot_kernel(int64_t D, const double* __restrict__ a, const double* __restrict__ b,
                                                      double* __restrict__ partial) {
  double s = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < D; j += (int64_t)gridDim.x * blockDim.x) s += a[j] * b[j];
  block_partial(s, partial);
}
// z = M^-1 r (block diagonal), partial = sum r z
__global__ void __launch_bounds__(256) cg_precond_kernel(int64_t D, BaIter I, const double* __restrict__ r,
                                                          double* __restrict__ z, double* __restrict__ partial) {
  double s = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < D; j += (int64_t)gridDim.x * blockDim.x) {
    const int f = I.blk_first[j], n = I.blk_size[j];
    double v = 0;
    for (int k = 0; k < n; ++k) v += I.M[I.m_stride * j + k] * r[f + k];
    z[j] = v;
    s += r[j] * v;
  }
  block_partial(s, partial);
}
// p = z (first) or z + beta p
__global__ void cg_update_p_kernel(int64_t D, const double* __restrict__ z, double* __restrict__ p, double beta, int first) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < D) p[j] = first ? z[j] : z[j] + beta * p[j];
}
// q += D_c^2 p (after the all-reduce of the F'(F p - E z) part); partial = sum p q
__global__ void __launch_bounds__(256) cg_finish_q_kernel(int64_t D, const double* __restrict__ lm_c, const double* __restrict__ p,
                                                           double* __restrict__ q, double* __restrict__ partial) {
  double s = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < D; j += (int64_t)gridDim.x * blockDim.x) {
    const double v = q[j] + lm_c[j] * p[j];
    q[j] = v;
    s += p[j] * v;
  }
  block_partial(s, partial);
}
// mode 0: x += alpha p, r -= alpha q          mode 1: x += alpha p only (a residual reset follows)
// mode 2: r = b - q (q = S x)                 modes 0, 2: partial[0..) = sum x (b + r), partial[stride..) = sum r r
__global__ void __launch_bounds__(256) cg_update_xr_kernel(int64_t D, double* __restrict__ x, const double* __restrict__ p,
                                                            double* __restrict__ r, const double* __restrict__ q,
                                                            const double* __restrict__ b, double alpha, int mode,
                                                            double* __restrict__ partial, int stride) {
  double sq = 0, sr = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < D; j += (int64_t)gridDim.x * blockDim.x) {
    double xv = x[j], rv = r[j];
    if (mode != 2) { xv = xv + alpha * p[j]; x[j] = xv; }
    if (mode == 0) rv = rv - alpha * q[j];
    if (mode == 2) rv = b[j] - q[j];
    if (mode != 1) { r[j] = rv; sq += xv * (b[j] + rv); sr += rv * rv; }
  }
  block_partial(sq, partial);
  block_partial(sr, partial + stride);
}

}  // namespace bit

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }
static inline int vec_grid(int64_t D) { return (int)std::max<int64_t>(1, std::min<int64_t>(kBaIterMaxPartials, (D + 255) / 256)); }

cudaError_t bai_launch_point_prepare(const BaDev& P, const BaIter& I, double radius, double min_diag, double max_diag, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  if (P.wide) bit::point_prepare_kernel<ObsJacW><<<nblk((int64_t)P.n_pts * bit::kLanesPerPoint, 256), 256, 0, s>>>(P, I, radius, min_diag, max_diag);
  else bit::point_prepare_kernel<ObsJac><<<nblk((int64_t)P.n_pts * bit::kLanesPerPoint, 256), 256, 0, s>>>(P, I, radius, min_diag, max_diag);
  return cudaGetLastError();
}
cudaError_t bai_launch_rhs(const BaDev& P, const BaIter& I, cudaStream_t s) {
  if (P.n_img == 0 || P.n_obs == 0) return cudaSuccess;
  if (P.wide) bit::image_pass_kernel<1, ObsJacW><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I, nullptr, I.tp, P.rhs, P.g_c, P.diag_c);
  else bit::image_pass_kernel<1, ObsJac><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I, nullptr, I.tp, P.rhs, P.g_c, P.diag_c);
  return cudaGetLastError();
}
cudaError_t bai_launch_camera_terms_image(const BaDev& P, const BaIter& I, cudaStream_t s) {
  if (P.n_img == 0 || P.n_obs == 0) return cudaSuccess;
  if (P.wide) bit::camera_terms_image_kernel<ObsJacW><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I);
  else bit::camera_terms_image_kernel<ObsJac><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I);
  return cudaGetLastError();
}
cudaError_t bai_launch_cam_diag(const BaDev& P, const BaIter& I, double radius, double min_diag, double max_diag, cudaStream_t s) {
  if (P.D == 0) return cudaSuccess;
  bit::cam_diag_kernel<<<nblk(P.D, 256), 256, 0, s>>>(P, I, radius, min_diag, max_diag);
  return cudaGetLastError();
}
cudaError_t bai_launch_precond(const BaDev& P, const BaIter& I, cudaStream_t s) {
  if (P.n_obs == 0) return cudaSuccess;
  if (P.wide) bit::precond_kernel<ObsJacW><<<nblk(P.n_obs, 128), 128, 0, s>>>(P, I);
  else bit::precond_kernel<ObsJac><<<nblk(P.n_obs, 128), 128, 0, s>>>(P, I);
  return cudaGetLastError();
}
cudaError_t bai_launch_precond_invert(const BaDev& P, const BaIter& I, cudaStream_t s) {
  if (P.D == 0) return cudaSuccess;
  if (P.wide) bit::precond_invert_kernel<12><<<nblk(P.D, 128), 128, 0, s>>>(P, I);
  else bit::precond_invert_kernel<4><<<nblk(P.D, 128), 128, 0, s>>>(P, I);
  return cudaGetLastError();
}
// out = F'(F x - E (E'E)^-1 E'F x): zeroes `out`, two kernels.  The caller all-reduces `out` and adds D_c^2 x.
cudaError_t bai_launch_matvec(const BaDev& P, const BaIter& I, const double* x, double* out, cudaStream_t s) {
  if (P.D == 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)P.D * sizeof(double), s);
  if (e != cudaSuccess) return e;
  if (P.n_obs == 0 || P.n_img == 0) return cudaSuccess;
  if (P.wide) {
    bit::matvec_point_kernel<ObsJacW><<<nblk((int64_t)P.n_pts * bit::kLanesPerPoint, 256), 256, 0, s>>>(P, x, I.zp);
    bit::image_pass_kernel<0, ObsJacW><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I, x, I.zp, out, nullptr, nullptr);
  } else {
    bit::matvec_point_kernel<ObsJac><<<nblk((int64_t)P.n_pts * bit::kLanesPerPoint, 256), 256, 0, s>>>(P, x, I.zp);
    bit::image_pass_kernel<0, ObsJac><<<P.n_img, bit::kImageThreads, 0, s>>>(P, I, x, I.zp, out, nullptr, nullptr);
  }
  return cudaGetLastError();
}
cudaError_t bai_launch_dot(int64_t D, const double* a, const double* b, double* partial, int* n_partial, cudaStream_t s) {
  *n_partial = vec_grid(D);
  bit::cg_dot_kernel<<<*n_partial, 256, 0, s>>>(D, a, b, partial);
  return cudaGetLastError();
}
cudaError_t bai_launch_cg_precond(int64_t D, const BaIter& I, const double* r, double* z, double* partial, int* n_partial, cudaStream_t s) {
  *n_partial = vec_grid(D);
  bit::cg_precond_kernel<<<*n_partial, 256, 0, s>>>(D, I, r, z, partial);
  return cudaGetLastError();
}
cudaError_t bai_launch_cg_update_p(int64_t D, const double* z, double* p, double beta, bool first, cudaStream_t s) {
  bit::cg_update_p_kernel<<<nblk(D, 256), 256, 0, s>>>(D, z, p, beta, first ? 1 : 0);
  return cudaGetLastError();
}
cudaError_t bai_launch_cg_finish_q(int64_t D, const double* lm_c, const double* p, double* q, double* partial, int* n_partial, cudaStream_t s) {
  *n_partial = vec_grid(D);
  bit::cg_finish_q_kernel<<<*n_partial, 256, 0, s>>>(D, lm_c, p, q, partial);
  return cudaGetLastError();
}
cudaError_t bai_launch_cg_update_xr(int64_t D, double* x, const double* p, double* r, const double* q, const double* b,
                                    double alpha, int mode, double* partial, int* n_partial, cudaStream_t s) {
  *n_partial = vec_grid(D);
  bit::cg_update_xr_kernel<<<*n_partial, 256, 0, s>>>(D, x, p, r, q, b, alpha, mode, partial, kBaIterMaxPartials);
  return cudaGetLastError();
}

}  // namespace b2
