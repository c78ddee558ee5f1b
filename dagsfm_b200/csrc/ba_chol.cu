// Tiled Cholesky of the reduced camera system in packed 64 x 64 tiles and the two triangular solves -- the linear solve
// of the exact Schur step (what Ceres' SPARSE_SCHUR / DENSE_SCHUR hand to CHOLMOD / LAPACK for the reference,
// src/optim/bundle_adjustment.cc:274-284).  Hand-written; no library on this path.
//
// Storage: upper block triangle, S = R'R, only the tiles that the camera graph and the symbolic fill (host, once per
// solve) make non-zero: a banded / block-sparse camera graph costs O(D b^2), a complete graph is an ordinary dense
// factorisation.
//
// ONE persistent kernel runs the whole solve as a task graph (left-looking, one task per tile, every tile written once):
//   DIAG(i)    A_ii -= sum_k R_ki' R_ki over the tiles above it, factored in registers (A = U' D^-1 U elimination, one
//              barrier per column, rows scaled at the end), R_ii and 1 / diag(R_ii) published; then, off the critical
//              path, the inverse of R_ii (the two solves are products with it).
//   OFF(i, j)  A_ij -= sum_k R_ki' R_kj, then R_ij = R_ii^-T A_ij by forward substitution, four threads per column.
//   FWD(i)     y_i = R_ii^-T (b_i - sum_k R_ki' y_k)        (runs alongside the factorisation)
//   BACK(i)    x_i = R_ii^-1 (y_i - sum_j R_ij x_j)         (last tile row first, after the factorisation)
// Tasks are numbered so that every task depends on lower numbers only; CTAs draw task numbers from a global counter, so a
// CTA that waits (acquire-spin on a per-task flag) always waits for a task that is running or done -- no deadlock
// whatever the number of resident CTAs.  The dependency lists are laid out by the host once per solve: the structure of
// a bundle-adjustment problem does not change across LM iterations.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ba_common.cuh"

namespace b2 {
namespace bac {

constexpr int TS = kST;  // 64
constexpr int kThreads = 256;
enum { TASK_DIAG = 0, TASK_OFF = 1, TASK_FWD = 2, TASK_BACK = 3 };

// ---- memory-ordering helpers: data written by another CTA of this launch is read through L2 (never a stale L1 line)
__device__ __forceinline__ double ld_l2(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ double2 ld_l2(const double2* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}
__device__ __forceinline__ int ld_flag(const int* p) {
#ifdef __CUDA_ARCH__
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return *(const volatile int*)p;
#endif
}
// all threads of the CTA call; returns once *f != 0 and the producer's writes are visible to every thread
__device__ __forceinline__ void wait_flag(const int* f) {
  if (threadIdx.x == 0) {
    while (ld_flag(f) == 0) {
#ifdef __CUDA_ARCH__
      __nanosleep(32);
#else
      // host build of the kernel (CPU test-suite): blocks run one after the other, so a flag that is not set here never
      // will be -- a broken task order is reported instead of hanging
      fprintf(stderr, "ba_chol: task waits for a flag that no earlier task set (task order broken)\n");
      abort();
#endif
    }
  }
  __syncthreads();
}
// all threads of the CTA call after their last write of the published data
__device__ __forceinline__ void publish(int* f) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicExch(f, 1);
  }
}

// ---- acc (4 x 4 per thread: rows 4 tr .., columns 4 tc ..) -= A' B for two 64 x 64 row-major tiles in shared memory
__device__ __forceinline__ void tile_product(const double* __restrict__ sA, const double* __restrict__ sB, int r0, int c0,
                                             double (&acc)[4][4]) {
#pragma unroll 4
  for (int l = 0; l < TS; ++l) {
    const double2 a01 = *reinterpret_cast<const double2*>(sA + l * TS + r0);
    const double2 a23 = *reinterpret_cast<const double2*>(sA + l * TS + r0 + 2);
    const double2 b01 = *reinterpret_cast<const double2*>(sB + l * TS + c0);
    const double2 b23 = *reinterpret_cast<const double2*>(sB + l * TS + c0 + 2);
    const double av[4] = {a01.x, a01.y, a23.x, a23.y}, bv[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fma(-av[i], bv[j], acc[i][j]);
  }
}
__device__ __forceinline__ void load_tile(const double* __restrict__ g, double* __restrict__ s) {
  const double2* g2 = reinterpret_cast<const double2*>(g);
  double2* s2 = reinterpret_cast<double2*>(s);
  double2 v[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) v[m] = ld_l2(g2 + threadIdx.x + kThreads * m);
#pragma unroll
  for (int m = 0; m < 8; ++m) s2[threadIdx.x + kThreads * m] = v[m];
}

// X = R^-T B for the 64 columns of B: forward substitution, four threads per column (thread q of a column holds rows
// q, q + 4, ...; the pivot entry travels by shuffle inside the quad).  sR: R (upper) row-major, sd: 1 / R_ll.
__device__ __forceinline__ void quad_trsm(const double* __restrict__ sR, const double* __restrict__ sd, double (&a)[16], int q) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int l = 0; l < TS; ++l) {
    const double mine = a[l >> 2] * sd[l];
    const double x = __shfl_sync(0xffffffffu, mine, (int)((lane & ~3u) | (unsigned)(l & 3)));
    if (q == (l & 3)) a[l >> 2] = x;
#pragma unroll
    for (int m = l >> 2; m < 16; ++m) {
      const int i = 4 * m + q;
      if (i > l) a[m] = fma(-sR[l * TS + i], x, a[m]);
    }
  }
}

// 1 / d to <= 1 ulp without the IEEE division's dependent chain (d a positive pivot of Jacobi-scaled normal equations;
// anything outside the single-precision seed's range takes the division)
__device__ __forceinline__ double fast_rcp(double d) {
#ifdef __CUDA_ARCH__
  if (d > 0x1p-100 && d < 0x1p100) {
    double y = (double)__frcp_rn((float)d);
    y = fma(y, fma(-d, y, 1.0), y);
    y = fma(y, fma(-d, y, 1.0), y);
    return y;
  }
#endif
  return 1.0 / d;
}

struct Plan {
  const int32_t* task;      // [n_tasks][4]: kind, i, j (OFF), target tile / -1
  const int32_t* dep_ptr;   // [n_tasks + 1]
  const int32_t* dep;       // [..][2]: DIAG/OFF (tile(k,i), tile(k,j)); FWD (tile(k,i), k); BACK (tile(i,j), j)
  int32_t n_tasks;
  int* flags;               // [n_tiles] tile | [nt] inverse | [nt] y | [nt] x | [1] task counter
  double* rdiag;            // [nt][64] 1 / diag(R_ii)
  unsigned long long* trace;  // optional [n_tasks][8] %globaltimer stamps: picked, dependencies consumed, published, done, + 4 phase marks
};
__device__ __forceinline__ void stamp(const Plan& Pn, int t, int k) {
#ifdef __CUDA_ARCH__
  if (Pn.trace && threadIdx.x == 0) {
    unsigned long long v;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
    Pn.trace[8 * (size_t)t + k] = v;
  }
#endif
}

__global__ void __launch_bounds__(kThreads, 2) solve_graph_kernel(BaTiles T, Plan Pn, double* __restrict__ x) {
  extern __shared__ double chol_smem[];
  double* sA = chol_smem;            // 64 x 64
  double* sB = chol_smem + TS * TS;  // 64 x 64
  __shared__ __align__(16) double ublk[2][4][TS];
  __shared__ double dblk[2][20];
  __shared__ double sd[TS], sv[TS], part[4][TS];
  __shared__ int s_task;
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15, r0 = 4 * tr, c0 = 4 * tc;
  int* f_tile = Pn.flags;
  int* f_inv = Pn.flags + T.n_tiles;
  int* f_y = f_inv + T.nt;
  int* f_x = f_y + T.nt;
  int* counter = f_x + T.nt;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_task;
    if (t >= Pn.n_tasks) break;
    const int kind = Pn.task[4 * t], i = Pn.task[4 * t + 1], j = Pn.task[4 * t + 2], tile = Pn.task[4 * t + 3];
    const int d0 = Pn.dep_ptr[t], d1 = Pn.dep_ptr[t + 1];
    stamp(Pn, t, 0);
    if (kind == TASK_DIAG || kind == TASK_OFF) {
      double* Aij = T.tiles + (size_t)tile * (TS * TS);
      double acc[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const double2 v01 = ld_l2(reinterpret_cast<const double2*>(Aij + (r0 + a) * TS + c0));
        const double2 v23 = ld_l2(reinterpret_cast<const double2*>(Aij + (r0 + a) * TS + c0 + 2));
        acc[a][0] = v01.x; acc[a][1] = v01.y; acc[a][2] = v23.x; acc[a][3] = v23.y;
      }
      for (int d = d0; d < d1; ++d) {
        const int ta = Pn.dep[2 * d], tb = Pn.dep[2 * d + 1];
        wait_flag(f_tile + ta);
        if (tb != ta) wait_flag(f_tile + tb);
        if (d == d1 - 1) stamp(Pn, t, 6);
        load_tile(T.tiles + (size_t)ta * (TS * TS), sA);
        if (tb != ta) load_tile(T.tiles + (size_t)tb * (TS * TS), sB);
        __syncthreads();
        if (d == d1 - 1) stamp(Pn, t, 7);
        tile_product(sA, tb != ta ? sB : sA, r0, c0, acc);
        __syncthreads();
      }
      stamp(Pn, t, 1);
      if (kind == TASK_DIAG) {
        // ---- A = U' D^-1 U by blocked elimination on the register blocks, four pivots per block barrier: the thread that
        // holds the 4 x 4 diagonal block eliminates it and hands its multipliers to the 15 threads holding the rest of
        // the four pivot rows (same warp: a warp barrier), they finish their rows, and the whole CTA applies the four
        // rank-1 updates at once.  The pivots' reciprocals come from a single-precision seed + two Newton steps (<= 1
        // ulp): the IEEE division is ~200 cycles of dependent instructions, 64 times on this kernel's critical path.
        bool bad = false;
#pragma unroll 1
        for (int p = 0; p < 16; ++p) {
          double (*ub)[TS] = ublk[p & 1];
          double* db = dblk[p & 1];  // [0, 16): the eliminated diagonal block, row-major; [16, 20): 1 / pivot
          if ((tr >> 1) == (p >> 1)) {  // the warp that holds block row p (two block rows per warp)
            if (tr == p && tc == p) {
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                double d = acc[a][a];
                if (!(d > 0.0)) { bad = true; d = 1.0; acc[a][a] = 1.0; }
                const double inv = fast_rcp(d);
                db[16 + a] = inv;
                sd[4 * p + a] = d;
#pragma unroll
                for (int a2 = a + 1; a2 < 4; ++a2) {
                  const double m = acc[a][a2] * inv;
#pragma unroll
                  for (int b2 = a2; b2 < 4; ++b2) acc[a2][b2] = fma(-m, acc[a][b2], acc[a2][b2]);
                }
              }
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) db[4 * a + b2] = acc[a][b2];
            }
            __syncwarp();
            if (tr == p && tc > p) {
#pragma unroll
              for (int a = 0; a < 3; ++a) {
                const double inv = db[16 + a];
#pragma unroll
                for (int a2 = a + 1; a2 < 4; ++a2) {
                  const double m = db[4 * a + a2] * inv;
#pragma unroll
                  for (int b2 = 0; b2 < 4; ++b2) acc[a2][b2] = fma(-m, acc[a][b2], acc[a2][b2]);
                }
              }
            }
            if (tr == p) {
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) ub[a][c0 + b2] = acc[a][b2];
            }
          }
          __syncthreads();
          if (tr > p) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const double inv = db[16 + k];
              const double2 r01 = *reinterpret_cast<const double2*>(&ub[k][r0]), r23 = *reinterpret_cast<const double2*>(&ub[k][r0 + 2]);
              const double2 c01 = *reinterpret_cast<const double2*>(&ub[k][c0]), c23 = *reinterpret_cast<const double2*>(&ub[k][c0 + 2]);
              const double mr[4] = {r01.x * inv, r01.y * inv, r23.x * inv, r23.y * inv}, uc[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = fma(-mr[a], uc[b2], acc[a][b2]);
            }
          }
        }
        __syncthreads();
        stamp(Pn, t, 4);
        if (bad) *T.info = 1;
        // R = D^-1/2 U; the tile in global memory and a copy in sA for the inverse
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int r = r0 + a;
          const double sq = sqrt(sd[r]), rsq = 1.0 / sq;
#pragma unroll
          for (int b2 = 0; b2 < 4; ++b2) {
            const int c = c0 + b2;
            const double v = (c > r) ? acc[a][b2] * rsq : (c == r ? sq : 0.0);
            sA[r * TS + c] = v;
            Aij[r * TS + c] = v;
          }
        }
        __syncthreads();
        if (tid < TS) {
          const double inv = 1.0 / sA[tid * TS + tid];
          sv[tid] = inv;
          Pn.rdiag[(size_t)i * TS + tid] = inv;
        }
        stamp(Pn, t, 5);
        publish(f_tile + tile);  // (contains the barrier that orders sv)
        stamp(Pn, t, 2);
        // ---- off the critical path: Y = R^-T I, stored transposed = R^-1 (upper)
        {
          const int c = tid >> 2, q = tid & 3;
          double a[16];
#pragma unroll
          for (int m = 0; m < 16; ++m) a[m] = (4 * m + q == c) ? 1.0 : 0.0;
          quad_trsm(sA, sv, a, q);
          double* Ri = T.rinv + (size_t)i * (TS * TS);
#pragma unroll
          for (int m = 0; m < 16; ++m) Ri[c * TS + 4 * m + q] = a[m];  // Rinv[c][r] = Y[r][c]; zero below the diagonal
        }
        publish(f_inv + i);
        stamp(Pn, t, 3);
      } else {
        // ---- R_ij = R_ii^-T A_ij
        const int tdiag = T.row_ptr[i];
        wait_flag(f_tile + tdiag);
        stamp(Pn, t, 2);  // (for an OFF task: the diagonal factor has arrived)
        load_tile(T.tiles + (size_t)tdiag * (TS * TS), sA);
        if (tid < TS) sv[tid] = ld_l2(Pn.rdiag + (size_t)i * TS + tid);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) sB[(r0 + a) * TS + c0 + b] = acc[a][b];
        __syncthreads();
        const int c = tid >> 2, q = tid & 3;
        double a[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = sB[(4 * m + q) * TS + c];
        stamp(Pn, t, 4);
        quad_trsm(sA, sv, a, q);
        __syncthreads();
        stamp(Pn, t, 5);
#pragma unroll
        for (int m = 0; m < 16; ++m) sB[(4 * m + q) * TS + c] = a[m];
        __syncthreads();
        {  // coalesced store of the finished tile
          double2* g2 = reinterpret_cast<double2*>(Aij);
          const double2* s2 = reinterpret_cast<const double2*>(sB);
#pragma unroll
          for (int m = 0; m < 8; ++m) g2[tid + kThreads * m] = s2[tid + kThreads * m];
        }
        publish(f_tile + tile);
        stamp(Pn, t, 3);
      }
      (void)j;
    } else if (kind == TASK_FWD) {
      // ---- y_i = Rinv_i' (b_i - sum_k R_ki' y_k)
      const int c = tid & 63, q = tid >> 6;
      double t_own = 0.0;  // thread (q, c): partial of sum_k (R_ki' y_k)[c] over l = 16 q ..
      for (int d = d0; d < d1; ++d) {
        const int ta = Pn.dep[2 * d], k = Pn.dep[2 * d + 1];
        wait_flag(f_tile + ta);
        wait_flag(f_y + k);
        if (tid < TS) sv[tid] = ld_l2(x + (size_t)k * TS + tid);
        __syncthreads();
        const double* R = T.tiles + (size_t)ta * (TS * TS);
#pragma unroll 4
        for (int l = 16 * q; l < 16 * q + 16; ++l) t_own = fma(ld_l2(R + l * TS + c), sv[l], t_own);
        __syncthreads();
      }
      part[q][c] = t_own;
      __syncthreads();
      if (tid < TS) sv[tid] = ld_l2(x + (size_t)i * TS + tid) - (part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid]);
      wait_flag(f_inv + i);  // (barrier inside orders sv)
      const double* Ri = T.rinv + (size_t)i * (TS * TS);
      double s = 0.0;
#pragma unroll 4
      for (int l = 16 * q; l < 16 * q + 16; ++l)
        if (l <= c) s = fma(ld_l2(Ri + l * TS + c), sv[l], s);
      part[q][c] = s;
      __syncthreads();
      if (tid < TS) x[(size_t)i * TS + tid] = part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid];
      publish(f_y + i);
      stamp(Pn, t, 3);
    } else {
      // ---- x_i = Rinv_i (y_i - sum_j R_ij x_j): a warp per eight rows, lanes across the columns
      const int lane = tid & 31, w = tid >> 5;
      double racc[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) racc[a] = 0.0;
      wait_flag(f_y + i);
      for (int d = d0; d < d1; ++d) {
        const int ta = Pn.dep[2 * d], jj = Pn.dep[2 * d + 1];
        wait_flag(f_x + jj);
        const double x0 = ld_l2(x + (size_t)jj * TS + lane), x1 = ld_l2(x + (size_t)jj * TS + lane + 32);
        const double* R = T.tiles + (size_t)ta * (TS * TS);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int r = 8 * w + a;
          racc[a] = fma(ld_l2(R + r * TS + lane), x0, fma(ld_l2(R + r * TS + lane + 32), x1, racc[a]));
        }
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        double v = racc[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sv[8 * w + a] = v;
      }
      __syncthreads();
      if (tid < TS) sd[tid] = ld_l2(x + (size_t)i * TS + tid) - sv[tid];
      __syncthreads();
      const double* Ri = T.rinv + (size_t)i * (TS * TS);
      const double s0 = sd[lane], s1 = sd[lane + 32];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int r = 8 * w + a;
        double v = fma(ld_l2(Ri + r * TS + lane), s0, ld_l2(Ri + r * TS + lane + 32) * s1);  // zero below the diagonal
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sv[r] = v;
      }
      __syncthreads();
      if (tid < TS) x[(size_t)i * TS + tid] = sv[tid];
      publish(f_x + i);
      stamp(Pn, t, 3);
    }
  }
}

}  // namespace bac

// Task graph of the solve for a tile pattern (closed under the symbolic fill): host, once per b2_ba_solve.
void bac_build_graph(int nt, const int32_t* tile_id, const int32_t* row_ptr, const int32_t* row_col, BaCholGraph* G) {
  G->task.clear();
  G->dep_ptr.assign(1, 0);
  G->dep.clear();
  auto tid_of = [&](int r, int c) { return tile_id[(size_t)r * nt + c]; };
  for (int i = 0; i < nt; ++i) {
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
      const int j = row_col[e];
      G->task.insert(G->task.end(), {j == i ? bac::TASK_DIAG : bac::TASK_OFF, i, j, e});
      for (int k = 0; k < i; ++k) {
        const int ta = tid_of(k, i), tb = tid_of(k, j);
        if (ta >= 0 && tb >= 0) G->dep.insert(G->dep.end(), {ta, tb});
      }
      G->dep_ptr.push_back((int32_t)(G->dep.size() / 2));
      if (j == i) {  // the forward solve of row i right behind its diagonal tile
        G->task.insert(G->task.end(), {bac::TASK_FWD, i, i, -1});
        for (int k = 0; k < i; ++k)
          if (tid_of(k, i) >= 0) G->dep.insert(G->dep.end(), {tid_of(k, i), k});
        G->dep_ptr.push_back((int32_t)(G->dep.size() / 2));
      }
    }
  }
  for (int i = nt - 1; i >= 0; --i) {
    G->task.insert(G->task.end(), {bac::TASK_BACK, i, i, -1});
    for (int e = row_ptr[i + 1] - 1; e > row_ptr[i]; --e) G->dep.insert(G->dep.end(), {e, row_col[e]});  // farthest column first: x_{i+1}, the last one written, is waited for last
    G->dep_ptr.push_back((int32_t)(G->dep.size() / 2));
  }
  G->n_tasks = (int32_t)(G->task.size() / 4);
}

size_t bac_flag_count(const BaTiles& T) { return (size_t)T.n_tiles + 3 * (size_t)T.nt + 1; }

// S x = b in place (x = b on entry, padded to whole tiles); S is overwritten by its factor.
cudaError_t bac_solve_system(const BaTiles& T, const BaCholDev& G, double* x, int n_sm, cudaStream_t s) {
  if (T.nt == 0) return cudaSuccess;
  constexpr int kSmem = 2 * bac::TS * bac::TS * (int)sizeof(double);  // 65 536 B: above the 48 KB static limit
  static bool attr_set = false;  // per process; the attribute is a property of the function, not of the handle
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(bac::solve_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaError_t e = cudaMemsetAsync(G.flags, 0, bac_flag_count(T) * sizeof(int), s);
  if (e != cudaSuccess) return e;
  bac::Plan Pn;
  Pn.task = G.task; Pn.dep_ptr = G.dep_ptr; Pn.dep = G.dep; Pn.n_tasks = G.n_tasks; Pn.flags = G.flags; Pn.rdiag = G.rdiag;
  Pn.trace = G.trace;
  const int grid = std::min(G.n_tasks, 2 * std::max(n_sm, 1));
  bac::solve_graph_kernel<<<grid, bac::kThreads, kSmem, s>>>(T, Pn, x);
  return cudaGetLastError();
}

}  // namespace b2
