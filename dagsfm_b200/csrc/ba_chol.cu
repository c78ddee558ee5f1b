// Tiled Cholesky of the reduced camera system in packed 64 x 64 tiles, and the two triangular solves -- the linear
// solve of the exact Schur step (what Ceres' SPARSE_SCHUR / DENSE_SCHUR hand to CHOLMOD / LAPACK for the reference,
// src/optim/bundle_adjustment.cc:274-284).  Hand-written; replaces the cuSOLVER potrf / potrs of the first build.
//
// Storage: upper block triangle, S = R'R, only the tiles that the camera graph and the symbolic fill (host, once per
// solve) make non-zero: a banded / block-sparse camera graph costs O(D b^2), a complete graph is an ordinary dense
// right-looking factorisation on all SMs.  Per tile row k:
//   panel_kernel   one CTA per tile of row k.  Every CTA factors the (already updated) diagonal tile A_kk = R_kk'R_kk in
//                  shared memory itself (64^3/3 flops, redundant but off the critical path of nobody); CTA 0 stores the
//                  inverse of R_kk (all the solves need), CTA c > 0 turns its tile into R_kj = R_kk^-T A_kj.
//   update_kernel  one CTA per pair (a <= b) of off-diagonal tiles of row k:  A_{ja, jb} -= R_{k, ja}' R_{k, jb}
//                  (4 x 4 register tiles, explicit DFMA, operands staged through shared memory in two k-halves).
// solve_kernel: one CTA walks the tile rows forward (R'y = b) and backward (R x = y); the diagonal solves are products
// with the stored inverse tiles, so a step is two 64-long dot products deep instead of a 64-step substitution chain.
#include <cuda_runtime.h>

#include <cstdint>

#include "ba_common.cuh"

namespace b2 {
namespace bac {

constexpr int TS = kST;        // 64
constexpr int LD = TS + 1;     // shared-memory row stride (bank-conflict padding)

// A (upper triangle valid) -> R in place: A = R'R.  All threads of the block take part (blockDim = 256, a 16 x 16 grid
// over the trailing block).  Elimination runs on UNSCALED rows (A = L D L' form: one barrier per column, every thread
// derives 1 / d_j itself), the rows are scaled by d_j^-1/2 at the end.
__device__ void factor_tile(double (*A)[LD], int* bad) {
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  for (int j = 0; j < TS - 1; ++j) {
    __syncthreads();
    const double d = A[j][j];
    const double inv_d = 1.0 / d;
    for (int r = j + 1 + tr; r < TS; r += 16) {
      const double m = A[j][r] * inv_d;
      for (int c = j + 1 + tc; c < TS; c += 16)
        if (c >= r) A[r][c] = fma(-m, A[j][c], A[r][c]);
    }
  }
  __syncthreads();
  for (int e = tid; e < TS * TS; e += blockDim.x) {
    const int r = e >> 6, c = e & 63;
    if (c < r) continue;
    const double d = A[r][r];
    if (c == r) continue;
    A[r][c] = A[r][c] / sqrt(d > 0.0 ? d : 1.0);
  }
  __syncthreads();
  if (tid < TS) {
    const double d = A[tid][tid];
    if (!(d > 0.0)) *bad = 1;
    A[tid][tid] = sqrt(d > 0.0 ? d : 1.0);
  }
  __syncthreads();
}

// X = R^-T B for the 64 columns of B (in place), R upper triangular in shared memory: forward substitution, one thread
// per column with the column in registers; row l of R is a broadcast read.
__device__ void trsm_tile(const double (*R)[LD], double (*B)[LD]) {
  const int c = threadIdx.x;
  if (c < TS) {
    double a[TS];
#pragma unroll
    for (int i = 0; i < TS; ++i) a[i] = B[i][c];
#pragma unroll
    for (int l = 0; l < TS; ++l) {
      const double x = a[l] / R[l][l];
      a[l] = x;
#pragma unroll
      for (int i = l + 1; i < TS; ++i) a[i] = fma(-R[l][i], x, a[i]);
    }
#pragma unroll
    for (int i = 0; i < TS; ++i) B[i][c] = a[i];
  }
}

constexpr int kPanelSmem = 2 * TS * LD * (int)sizeof(double);  // 66 560 B: above the 48 KB static limit, opted in at launch
__global__ void __launch_bounds__(256) panel_kernel(BaTiles T, int k) {
  extern __shared__ double panel_smem[];
  double (*sA)[LD] = reinterpret_cast<double (*)[LD]>(panel_smem);
  double (*sB)[LD] = reinterpret_cast<double (*)[LD]>(panel_smem + TS * LD);
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const int t0 = T.row_ptr[k];
  double* Akk = T.tiles + (size_t)t0 * (TS * TS);
  if (tid == 0) s_bad = 0;
  for (int e = tid; e < TS * TS; e += blockDim.x) sA[e >> 6][e & 63] = Akk[e];
  __syncthreads();
  factor_tile(sA, &s_bad);
  if (blockIdx.x == 0) {
    if (tid == 0 && s_bad) *T.info = 1;
    // A_kk itself stays untouched in global memory: the other CTAs of this launch are still reading it, and nothing
    // after this launch needs R_kk (the solves work with its inverse)
    // the inverse of R_kk: X = R^-T I = (R^-1)', stored transposed back as the upper triangular R^-1
    for (int e = tid; e < TS * TS; e += blockDim.x) sB[e >> 6][e & 63] = ((e >> 6) == (e & 63)) ? 1.0 : 0.0;
    __syncthreads();
    trsm_tile(sA, sB);
    __syncthreads();
    double* Ri = T.rinv + (size_t)k * (TS * TS);
    for (int e = tid; e < TS * TS; e += blockDim.x) {
      const int r = e >> 6, c = e & 63;
      Ri[e] = (c >= r) ? sB[c][r] : 0.0;
    }
  } else {
    double* Akj = T.tiles + (size_t)(t0 + blockIdx.x) * (TS * TS);
    for (int e = tid; e < TS * TS; e += blockDim.x) sB[e >> 6][e & 63] = Akj[e];
    __syncthreads();
    trsm_tile(sA, sB);
    __syncthreads();
    for (int e = tid; e < TS * TS; e += blockDim.x) Akj[e] = sB[e >> 6][e & 63];
  }
}

// A_{ja, jb} -= R_{k, ja}' R_{k, jb} for the pair (a <= b) = blockIdx.x of the off-diagonal tiles of row k.
__global__ void __launch_bounds__(256) update_kernel(BaTiles T, int k) {
  __shared__ __align__(16) double sA[32][TS];
  __shared__ __align__(16) double sB[32][TS];
  const int tid = threadIdx.x;
  const int t0 = T.row_ptr[k], n = T.row_ptr[k + 1] - t0 - 1;
  int a = 0, rem = blockIdx.x;
  while (rem >= n - a) { rem -= n - a; ++a; }
  const int b = a + rem;
  const int ja = T.row_col[t0 + 1 + a], jb = T.row_col[t0 + 1 + b];
  const double* Ra = T.tiles + (size_t)(t0 + 1 + a) * (TS * TS);
  const double* Rb = T.tiles + (size_t)(t0 + 1 + b) * (TS * TS);
  double* C = T.tiles + (size_t)T.tile_id[ja * T.nt + jb] * (TS * TS);
  const int r0 = (tid >> 4) * 4, c0 = (tid & 15) * 4;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int half = 0; half < 2; ++half) {
    for (int e = tid; e < 32 * TS; e += blockDim.x) {
      (&sA[0][0])[e] = Ra[half * 32 * TS + e];
      (&sB[0][0])[e] = Rb[half * 32 * TS + e];
    }
    __syncthreads();
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      const double2 a01 = *reinterpret_cast<const double2*>(&sA[l][r0]);
      const double2 a23 = *reinterpret_cast<const double2*>(&sA[l][r0 + 2]);
      const double2 b01 = *reinterpret_cast<const double2*>(&sB[l][c0]);
      const double2 b23 = *reinterpret_cast<const double2*>(&sB[l][c0 + 2]);
      const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool diag = (ja == jb);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + i, c = c0 + j;
      if (diag && c < r) continue;
      C[r * TS + c] -= acc[i][j];
    }
}

// R'R x = b in place (x = b on entry).  One CTA of 512 threads = 8 groups of 64.
__global__ void __launch_bounds__(512) solve_kernel(BaTiles T, double* __restrict__ x) {
  __shared__ double sv[TS];
  __shared__ double part[8][TS];
  const int tid = threadIdx.x, g = tid >> 6, c = tid & 63;
  // ---- forward: R' y = b
  for (int k = 0; k < T.nt; ++k) {
    const int t0 = T.row_ptr[k], n = T.row_ptr[k + 1] - t0 - 1;
    double* xk = x + (size_t)k * TS;
    if (tid < TS) sv[tid] = xk[tid];
    __syncthreads();
    if (tid < TS) {  // y_k = (R_kk^-1)' b_k
      const double* Ri = T.rinv + (size_t)k * (TS * TS);
      double s = 0;
      for (int l = 0; l <= tid; ++l) s = fma(Ri[l * TS + tid], sv[l], s);
      xk[tid] = s;
    }
    __syncthreads();
    if (tid < TS) sv[tid] = xk[tid];
    __syncthreads();
    for (int t = g; t < n; t += 8) {  // b_j -= R_kj' y_k
      const double* R = T.tiles + (size_t)(t0 + 1 + t) * (TS * TS);
      double s = 0;
      for (int l = 0; l < TS; ++l) s = fma(R[l * TS + c], sv[l], s);
      x[(size_t)T.row_col[t0 + 1 + t] * TS + c] -= s;
    }
    __syncthreads();
  }
  // ---- backward: R x = y
  for (int k = T.nt - 1; k >= 0; --k) {
    const int t0 = T.row_ptr[k], n = T.row_ptr[k + 1] - t0 - 1;
    double* xk = x + (size_t)k * TS;
    double s = 0;
    for (int t = g; t < n; t += 8) {  // sum_j R_kj x_j, row c of each tile
      const double* R = T.tiles + (size_t)(t0 + 1 + t) * (TS * TS) + c * TS;
      const double* xj = x + (size_t)T.row_col[t0 + 1 + t] * TS;
      for (int l = 0; l < TS; ++l) s = fma(R[l], xj[l], s);
    }
    part[g][c] = s;
    __syncthreads();
    if (tid < TS) {
      double t = xk[tid];
      for (int q = 0; q < 8; ++q) t -= part[q][tid];
      sv[tid] = t;
    }
    __syncthreads();
    if (tid < TS) {  // x_k = R_kk^-1 s
      const double* Ri = T.rinv + (size_t)k * (TS * TS) + tid * TS;
      double t = 0;
      for (int l = tid; l < TS; ++l) t = fma(Ri[l], sv[l], t);
      xk[tid] = t;
    }
    __syncthreads();
  }
}

}  // namespace bac

cudaError_t bac_factor(const BaTiles& T, const int32_t* h_row_ptr, const int32_t* h_row_col, const int32_t* h_tile_id,
                       cudaStream_t s, int* n_launches) {
  (void)h_row_col; (void)h_tile_id;
  int launches = 0;
  static bool attr_set = false;  // per process; the attribute is a property of the function, not of the handle
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bac::panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bac::kPanelSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  for (int k = 0; k < T.nt; ++k) {
    const int n = h_row_ptr[k + 1] - h_row_ptr[k];  // tiles in row k, the diagonal one included
    bac::panel_kernel<<<n, 256, bac::kPanelSmem, s>>>(T, k);
    ++launches;
    if (n > 1) {
      const int m = n - 1;
      bac::update_kernel<<<m * (m + 1) / 2, 256, 0, s>>>(T, k);
      ++launches;
    }
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}

cudaError_t bac_solve(const BaTiles& T, double* x, int64_t D, cudaStream_t s) {
  (void)D;
  if (T.nt == 0) return cudaSuccess;
  bac::solve_kernel<<<1, 512, 0, s>>>(T, x);
  return cudaGetLastError();
}

}  // namespace b2
