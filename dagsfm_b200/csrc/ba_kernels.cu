// Bundle adjustment kernels: reprojection residual + analytic Jacobian, camera normal
// blocks, the Schur-complement reduce-to-camera-block, back-substitution, step
// application and cost evaluation.  FP64 throughout, compiled with --fmad=false.
//
// Reference arithmetic restated (file:line under the reference tree):
//   BundleAdjustmentCostFunction::operator()   src/base/cost_functions.h:57-84
//   SimpleRadialCameraModel::WorldToImage      src/base/camera_models.h:714-732,748-757
//   ceres::UnitQuaternionRotatePoint, QuaternionParameterization (Ceres 1.14, external)
// The reference evaluates the Jacobian by Ceres autodiff of these formulas and leaves the
// Schur elimination to Ceres' SchurEliminator; here both are explicit.
//
// HBM layout (SoA over observations sorted by point = CSR):
//   obs_img int32[n_obs], obs_xy double2[n_obs], pt_start int64[n_pts+1]
//   qvec double[n_img*4], tvec double[n_img*3], cam_params double[n_cam*4], xyz double[n_pts*3]
//   pose_col int32[n_img*6], intr_col int32[n_cam*4]: column in the reduced system or -1
//   ObsJac[n_obs]: residual, 2x10 camera Jacobian, 2x3 point Jacobian (Jacobi-scaled)
//   S double[D*D] (row-major, entries with col_row <= col_col), rhs / g_c / diag_c double[D]
#include <cuda_runtime.h>

#include <cfloat>

#include <cstdint>

#include "ba_common.cuh"
#include "ba_eval.cuh"
#include "ba_loss.cuh"

namespace b2 {
namespace bak {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

// ---------------------------------------------------------------- Jacobian
// Thread per observation: residual, Jacobian (optionally Jacobi-scaled), cost.
// mode 0: write ObsJac; mode 1: residual only at candidate parameters (cost);
// mode 2: unscaled Jacobian -> squared column norms (Jacobi scaling set-up).
//
// LOSS (BundleAdjustmentOptions::LossFunctionType, bundle_adjustment.cc:53-68): 0 TRIVIAL, 1 SOFT_L1,
// 2 CAUCHY with scale a.  Ceres' ResidualBlock::Evaluate corrects the block's Jacobians and
// residuals with Corrector (corrector.cc): for rho'' <= 0 -- always true for these two losses --
// both are scaled by sqrt(rho'(s)), s = |r|^2, and the block's cost is rho(s) / 2.  LOSS == 0
// compiles to exactly the code it was before the template parameter existed.
template <int LOSS, class J>
__global__ void __launch_bounds__(256)
jacobian_kernel(BaDev P, const double* __restrict__ q, const double* __restrict__ t, const double* __restrict__ kp,
                const double* __restrict__ X, int mode, double* __restrict__ cost_out, double loss_scale) {
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double c = 0;
  if (o < P.n_obs) {
    const int i = P.obs_img[o], p = P.obs_pt[o], cm = P.img_cam[i];
    const double2 xy = P.obs_xy[o];
    constexpr int KI = J::kKI, NC = J::kNC;
    double r[2], Jc[2 * NC], Jp[6];
    double rho0 = 0;
    if (mode == 1) {
      evaluate<KI>(P.cam_model[cm], q + 4 * i, t + 3 * i, X + 3 * (int64_t)p, kp + KI * cm, xy.x, xy.y, r, nullptr, nullptr);
      if (LOSS != 0) {
        double w;
        loss_eval<LOSS>(loss_scale, r[0] * r[0] + r[1] * r[1], &rho0, &w);
      }
    } else {
      evaluate<KI>(P.cam_model[cm], q + 4 * i, t + 3 * i, X + 3 * (int64_t)p, kp + KI * cm, xy.x, xy.y, r, Jc, Jp);
      if (LOSS != 0) {
        double w;
        loss_eval<LOSS>(loss_scale, r[0] * r[0] + r[1] * r[1], &rho0, &w);
        r[0] *= w;
        r[1] *= w;
#pragma unroll
        for (int k = 0; k < 2 * NC; ++k) Jc[k] *= w;
#pragma unroll
        for (int k = 0; k < 6; ++k) Jp[k] *= w;
      }
      int col[NC];
#pragma unroll
      for (int k = 0; k < 6; ++k) col[k] = P.pose_col[6 * i + k];
#pragma unroll
      for (int k = 0; k < KI; ++k) col[6 + k] = P.intr_col[KI * cm + k];
      const int pc = P.pt_col[p];
      if (mode == 2) {
#pragma unroll
        for (int k = 0; k < NC; ++k)
          if (col[k] >= 0) atomicAdd(P.colnorm_c + col[k], Jc[k] * Jc[k] + Jc[NC + k] * Jc[NC + k]);
        if (pc >= 0)
#pragma unroll
          for (int k = 0; k < 3; ++k) atomicAdd(P.colnorm_p + 3 * (int64_t)pc + k, Jp[k] * Jp[k] + Jp[3 + k] * Jp[3 + k]);
      } else {
        J& e = jac<J>(P)[o];
        e.r[0] = r[0];
        e.r[1] = r[1];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          const double s = (col[k] >= 0) ? P.scale_c[col[k]] : 0.0;
          e.Jc[k] = Jc[k] * s;
          e.Jc[NC + k] = Jc[NC + k] * s;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double s = (pc >= 0) ? P.scale_p[3 * (int64_t)pc + k] : 0.0;
          e.Jp[k] = Jp[k] * s;
          e.Jp[3 + k] = Jp[3 + k] * s;
        }
      }
    }
    c = (LOSS != 0) ? rho0 : r[0] * r[0] + r[1] * r[1];
  }
  c = warp_sum(c);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += ws[k];
    atomicAdd(cost_out, 0.5 * s);
  }
}

// ------------------------------------------------------------ camera terms
// Thread per observation: U = Jc^T Jc into S (upper entries), g_c = Jc^T r, diag_c.
template <class J>
__global__ void __launch_bounds__(256) camera_terms_kernel(BaDev P) {
  constexpr int KI = J::kKI, NC = J::kNC;
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (o >= P.n_obs) return;
  const int i = P.obs_img[o], cm = P.img_cam[i];
  int col[NC];
#pragma unroll
  for (int k = 0; k < 6; ++k) col[k] = P.pose_col[6 * i + k];
#pragma unroll
  for (int k = 0; k < KI; ++k) col[6 + k] = P.intr_col[KI * cm + k];
  const J& e = jac<J>(P)[o];
  const int64_t D = P.D;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    if (col[k] < 0) continue;
    const double a0 = e.Jc[k], a1 = e.Jc[NC + k];
    atomicAdd(P.g_c + col[k], a0 * e.r[0] + a1 * e.r[1]);
    atomicAdd(P.diag_c + col[k], a0 * a0 + a1 * a1);
#pragma unroll
    for (int l = 0; l < NC; ++l) {
      if (col[l] < 0 || col[k] > col[l]) continue;
      atomicAdd(P.S + col[k] * D + col[l], a0 * e.Jc[l] + a1 * e.Jc[NC + l]);
    }
  }
}

// ----------------------------------------------------------------- Schur
// One block per point (grid-stride): V = sum Jp^T Jp + D_p, V^-1, t_p = V^-1 g_p,
// W_a = Jc_a^T Jp_a, Y_a = W_a V^-1; rhs -= W_a t_p; S -= Y_a W_b^T for all pairs (a,b)
// of the point's observations (only entries with col_a <= col_b: one triangle feeds the
// Cholesky).  Tracks longer than kTile are processed in kTile x kTile tiles.
constexpr int kTile = 32;
constexpr int kSchurThreads = 128;

// PM (pair-major variant, B2_BA_SCHUR=blocks): the kernel stops after W_a / Y_a, stores them per
// observation (Wg, Yg: [n_obs][30]) and leaves S -= Y_a W_b^T to pm_blocks_kernel, which sums the
// contributions of one (image, image) block over all points before touching S.  PM == false is the
// production kernel, unchanged.
template <bool PM, class J>
__global__ void __launch_bounds__(kSchurThreads) schur_kernel(BaDev P, double radius, double min_diag, double max_diag,
                                                              double* __restrict__ Wg, double* __restrict__ Yg) {
  constexpr int KI = J::kKI, NC = J::kNC, NW = 3 * NC;
  __shared__ double sWa[kTile][NW], sYa[kTile][NW], sWb[kTile][NW];
  __shared__ int sCa[kTile][NC], sCb[kTile][NC];
  __shared__ double sV[9], sT[3], sG[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t D = P.D;
  for (int p = blockIdx.x; p < P.n_pts; p += gridDim.x) {
    const int pc = P.pt_col[p];
    if (pc < 0) continue;  // constant point: contributes camera terms only
    const int64_t o0 = P.pt_start[p];
    const int L = (int)(P.pt_start[p + 1] - o0);
    // ---- V, g_p (warp 0)
    if (warp == 0) {
      double v[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
      for (int a = lane; a < L; a += 32) {
        const J& e = jac<J>(P)[o0 + a];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const double j0 = e.Jp[3 * i], j1 = e.Jp[3 * i + 1], j2 = e.Jp[3 * i + 2], r = e.r[i];
          v[0] += j0 * j0; v[1] += j0 * j1; v[2] += j0 * j2; v[3] += j1 * j1; v[4] += j1 * j2; v[5] += j2 * j2;
          g[0] += j0 * r; g[1] += j1 * r; g[2] += j2 * r;
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
#pragma unroll
      for (int k = 0; k < 3; ++k) g[k] = warp_sum(g[k]);
      if (lane == 0) {
        P.diag_p[3 * (int64_t)pc] = v[0];
        P.diag_p[3 * (int64_t)pc + 1] = v[3];
        P.diag_p[3 * (int64_t)pc + 2] = v[5];
        P.g_p[3 * (int64_t)pc] = g[0];
        P.g_p[3 * (int64_t)pc + 1] = g[1];
        P.g_p[3 * (int64_t)pc + 2] = g[2];
        // gradient max norm of the unscaled problem
        double gm = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) gm = fmax(gm, fabs(g[k] / P.scale_p[3 * (int64_t)pc + k]));
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
        for (int k = 0; k < 9; ++k) { sV[k] = Vi[k]; P.Vinv[9 * (int64_t)pc + k] = Vi[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) { sT[k] = Vi[3 * k] * g[0] + Vi[3 * k + 1] * g[1] + Vi[3 * k + 2] * g[2]; sG[k] = g[k]; }
      }
    }
    __syncthreads();
    for (int a0 = 0; a0 < L; a0 += kTile) {
      const int na = min(kTile, L - a0);
      // W_a, Y_a, columns of the a-tile; rhs -= W_a t_p
      for (int e = tid; e < na * NC; e += kSchurThreads) {
        const int a = e / NC, k = e % NC;
        const J& ob = jac<J>(P)[o0 + a0 + a];
        const int i = P.obs_img[o0 + a0 + a];
        const int col = (k < 6) ? P.pose_col[6 * i + k] : P.intr_col[KI * P.img_cam[i] + (k - 6)];
        sCa[a][k] = col;
        double w[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) w[l] = ob.Jc[k] * ob.Jp[l] + ob.Jc[NC + k] * ob.Jp[3 + l];
#pragma unroll
        for (int l = 0; l < 3; ++l) {
          sWa[a][3 * k + l] = w[l];
          sYa[a][3 * k + l] = w[0] * sV[l] + w[1] * sV[3 + l] + w[2] * sV[6 + l];
        }
        if (col >= 0) atomicAdd(P.rhs + col, -(w[0] * sT[0] + w[1] * sT[1] + w[2] * sT[2]));
        if (PM) {
#pragma unroll
          for (int l = 0; l < 3; ++l) {
            Wg[(o0 + a0 + a) * NW + 3 * k + l] = w[l];
            Yg[(o0 + a0 + a) * NW + 3 * k + l] = w[0] * sV[l] + w[1] * sV[3 + l] + w[2] * sV[6 + l];
          }
        }
      }
      if (PM) continue;
      __syncthreads();
      // Unordered observation pairs a <= b only: Y_a W_b^T and Y_b W_a^T are transposes of each other
      // (V^-1 is symmetric), so every entry of the block is written once, at (min col, max col) --
      // the triangle the Cholesky reads.  Thread (k, l) owns one entry of the 10 x 10 block and walks
      // the pairs: no integer division in the loop, Y_a stays in registers across b.
      for (int b0 = a0; b0 < L; b0 += kTile) {
        const int nb = min(kTile, L - b0);
        if (b0 == a0) {
          for (int e = tid; e < nb * NW; e += kSchurThreads) sWb[e / NW][e % NW] = sWa[e / NW][e % NW];
          for (int e = tid; e < nb * NC; e += kSchurThreads) sCb[e / NC][e % NC] = sCa[e / NC][e % NC];
        } else {
          for (int e = tid; e < nb * NC; e += kSchurThreads) {
            const int b = e / NC, k = e % NC;
            const J& ob = jac<J>(P)[o0 + b0 + b];
            const int i = P.obs_img[o0 + b0 + b];
            sCb[b][k] = (k < 6) ? P.pose_col[6 * i + k] : P.intr_col[KI * P.img_cam[i] + (k - 6)];
#pragma unroll
            for (int l = 0; l < 3; ++l) sWb[b][3 * k + l] = ob.Jc[k] * ob.Jp[l] + ob.Jc[NC + k] * ob.Jp[3 + l];
          }
        }
        __syncthreads();
        // NC = 10: the first 100 threads own one entry each; NC = 18: 324 entries, up to three per thread
        for (int ent = tid; ent < NC * NC; ent += kSchurThreads) {
          const int k = ent / NC, l = ent - NC * k;
          for (int a = 0; a < na; ++a) {
            const int ca = sCa[a][k];
            if (ca < 0) continue;
            const double y0 = sYa[a][3 * k], y1 = sYa[a][3 * k + 1], y2 = sYa[a][3 * k + 2];
            const bool same_tile = (b0 == a0);
            for (int b = same_tile ? a : 0; b < nb; ++b) {
              const int cb = sCb[b][l];
              if (cb < 0) continue;
              const bool diag_pair = same_tile && (b == a);
              if (diag_pair && ca > cb) continue;  // within one observation keep the upper entries only
              double v = y0 * sWb[b][3 * l] + y1 * sWb[b][3 * l + 1] + y2 * sWb[b][3 * l + 2];
              if (!diag_pair && ca == cb) v = v + v;  // shared intrinsics: (a,b) and (b,a) hit one diagonal entry
              const int r = min(ca, cb), c = max(ca, cb);
              atomicAdd(P.S + r * D + c, -v);
            }
          }
        }
        __syncthreads();
      }
    }
    if (PM) __syncthreads();  // sV / sT are rewritten by warp 0 for the next point
  }
}

// ------------------------------------------------------------ pair-major Schur (experimental)
// The observation pairs (a <= b) of every variable point, keyed by their (image, image) block with
// the smaller image index first.  FILL == false counts the tuples of each block, FILL == true writes
// them behind the block's start offset (counting sort; the order inside a block is arbitrary).  The
// structure does not change across LM iterations, so both run once per solve.  One warp per point.
template <bool FILL>
__global__ void __launch_bounds__(256)
pm_enumerate_kernel(BaDev P, int n_img, uint32_t* __restrict__ count_or_cursor, const uint32_t* __restrict__ start,
                    int2* __restrict__ tuples) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp0; p < P.n_pts; p += n_warps) {
    const bool variable = P.pt_col[p] >= 0;
    const int64_t o0 = P.pt_start[p];
    const int L = (int)(P.pt_start[p + 1] - o0);
    for (int a = 0; a < L; ++a) {
      const int ia = P.obs_img[o0 + a];
      // a constant point has no Schur term (W = 0) but its observations still carry camera terms:
      // only the diagonal tuple (a, a)
      for (int b = a + lane; b < (variable ? L : a + 1); b += 32) {
        const int ib = P.obs_img[o0 + b];
        const bool swap = ib < ia;
        const int lo_img = swap ? ib : ia, hi_img = swap ? ia : ib;
        const int lo_obs = (int)(o0 + (swap ? b : a)), hi_obs = (int)(o0 + (swap ? a : b));
        const int64_t key = (int64_t)lo_img * n_img + hi_img;
        const uint32_t pos = atomicAdd(count_or_cursor + key, 1u);
        if (FILL) tuples[start[key] + pos] = make_int2(lo_obs, hi_obs);
      }
    }
  }
}

// One warp per (image i <= image j) block: B = sum over the block's tuples of Y_lo W_hi^T (10 x 10, lane
// e owns entries e, e + 32, e + 64, e + 96), then one atomicAdd per entry into S instead of one per tuple
// and entry.  Same-image blocks are symmetric: a tuple of two different observations of one image adds
// its transpose too, and only the upper entries are written; their diagonal tuples (a, a) also carry the
// observation's camera terms U = Jc^T Jc, g_c, diag_c, so camera_terms_kernel is not run in this mode.
// Different images that share a camera:
// the intrinsics entries (k, l) and (l, k) land on the same upper slot, the diagonal ones are doubled
// (their (hi, lo) counterpart is not enumerated) -- as in schur_kernel's pair loop.
__global__ void __launch_bounds__(256)
pm_blocks_kernel(BaDev P, int n_img, const uint32_t* __restrict__ start, const int2* __restrict__ tuples,
                 const double* __restrict__ Wg, const double* __restrict__ Yg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_keys = (int64_t)n_img * n_img;
  const int64_t D = P.D;
  int kk[4], ll[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int e = lane + 32 * u;
    kk[u] = e / 10;
    ll[u] = e - 10 * kk[u];
  }
  for (int64_t key = warp0; key < n_keys; key += n_warps) {
    const uint32_t t0 = start[key], t1 = start[key + 1];
    if (t0 == t1) continue;
    const int i = (int)(key / n_img), j = (int)(key - (int64_t)i * n_img);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double gsum = 0.0, dsum = 0.0;  // lanes 0..9 of an (i, i) block: g_c and diag_c of column k = lane
    for (uint32_t t = t0; t < t1; ++t) {
      const int2 tp = tuples[t];
      if (tp.x == tp.y) {
        // the observation's own camera terms (what camera_terms_kernel adds with one atomic per entry):
        // U = Jc^T Jc enters S with the opposite sign of the Schur term, g_c = Jc^T r, diag_c = diag(U)
        const ObsJac& e = P.J[tp.x];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (lane + 32 * u >= 100) continue;
          const int k = kk[u], l = ll[u];
          acc[u] -= e.Jc[k] * e.Jc[l] + e.Jc[10 + k] * e.Jc[10 + l];
        }
        if (lane < 10) {
          gsum += e.Jc[lane] * e.r[0] + e.Jc[10 + lane] * e.r[1];
          dsum += e.Jc[lane] * e.Jc[lane] + e.Jc[10 + lane] * e.Jc[10 + lane];
        }
      }
      const double* y = Yg + (int64_t)tp.x * 30;
      const double* w = Wg + (int64_t)tp.y * 30;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (lane + 32 * u >= 100) continue;
        const int k = kk[u], l = ll[u];
        acc[u] += y[3 * k] * w[3 * l] + y[3 * k + 1] * w[3 * l + 1] + y[3 * k + 2] * w[3 * l + 2];
      }
      if (i == j && tp.x != tp.y) {
        const double* y2 = Yg + (int64_t)tp.y * 30;
        const double* w2 = Wg + (int64_t)tp.x * 30;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (lane + 32 * u >= 100) continue;
          const int k = kk[u], l = ll[u];
          acc[u] += y2[3 * k] * w2[3 * l] + y2[3 * k + 1] * w2[3 * l + 1] + y2[3 * k + 2] * w2[3 * l + 2];
        }
      }
    }
    const int cam_i = P.img_cam[i], cam_j = P.img_cam[j];
    if (i == j && lane < 10) {
      const int col = (lane < 6) ? P.pose_col[6 * i + lane] : P.intr_col[4 * cam_i + (lane - 6)];
      if (col >= 0) {
        atomicAdd(P.g_c + col, gsum);
        atomicAdd(P.diag_c + col, dsum);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (lane + 32 * u >= 100) continue;
      const int k = kk[u], l = ll[u];
      const int ca = (k < 6) ? P.pose_col[6 * i + k] : P.intr_col[4 * cam_i + (k - 6)];
      const int cb = (l < 6) ? P.pose_col[6 * j + l] : P.intr_col[4 * cam_j + (l - 6)];
      if (ca < 0 || cb < 0) continue;
      double v = acc[u];
      if (i == j) {
        if (ca > cb) continue;
      } else if (ca == cb) {
        v = v + v;
      }
      const int r = min(ca, cb), c = max(ca, cb);
      atomicAdd(P.S + r * D + c, -v);
    }
  }
}

// S_jj += clamp(diag_c) / radius, rhs = g_c - sum W V^-1 g_p, gradient max norm over camera columns.
__global__ void add_diag_kernel(BaDev P, double radius, double min_diag, double max_diag) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= P.D) return;
  P.S[j * P.D + j] += fmin(fmax(P.diag_c[j], min_diag), max_diag) / radius;
  P.rhs[j] += P.g_c[j];  // rhs held only the Schur part -sum W V^-1 g_p so far
  atomic_max_nonneg(P.gmax, grad_norm_term(P, j));
}

// ----------------------------------------------------------- back-substitution
// Warp per point: dp = -V^-1 (g_p + sum_a Jp_a^T (Jc_a dc)).
template <class J>
__global__ void __launch_bounds__(256) backsub_kernel(BaDev P) {
  constexpr int KI = J::kKI, NC = J::kNC;
  const int lane = threadIdx.x & 31;
  const int64_t wid = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (wid >= P.n_pts) return;
  const int p = (int)wid, pc = P.pt_col[p];
  if (pc < 0) return;
  const int64_t o0 = P.pt_start[p];
  const int L = (int)(P.pt_start[p + 1] - o0);
  double s[3] = {0, 0, 0};
  for (int a = lane; a < L; a += 32) {
    const J& e = jac<J>(P)[o0 + a];
    const int i = P.obs_img[o0 + a], cm = P.img_cam[i];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double jd = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) { const int c = P.pose_col[6 * i + k]; if (c >= 0) jd += e.Jc[NC * r + k] * P.dc[c]; }
#pragma unroll
      for (int k = 0; k < KI; ++k) { const int c = P.intr_col[KI * cm + k]; if (c >= 0) jd += e.Jc[NC * r + 6 + k] * P.dc[c]; }
#pragma unroll
      for (int k = 0; k < 3; ++k) s[k] += e.Jp[3 * r + k] * jd;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) s[k] = warp_sum(s[k]) + P.g_p[3 * (int64_t)pc + k];
  if (lane < 3) {
    const double* Vi = P.Vinv + 9 * (int64_t)pc;
    P.dp[3 * (int64_t)pc + lane] = -(Vi[3 * lane] * s[0] + Vi[3 * lane + 1] * s[1] + Vi[3 * lane + 2] * s[2]);
  }
}

// model_cost_change = -sum m (r + m/2), m = J delta (thread per observation)
template <class J>
__global__ void __launch_bounds__(256) model_cost_kernel(BaDev P, double* out) {
  constexpr int KI = J::kKI, NC = J::kNC;
  const int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double acc = 0;
  if (o < P.n_obs) {
    const J& e = jac<J>(P)[o];
    const int i = P.obs_img[o], cm = P.img_cam[i], pc = P.pt_col[P.obs_pt[o]];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      double m = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) { const int c = P.pose_col[6 * i + k]; if (c >= 0) m += e.Jc[NC * r + k] * P.dc[c]; }
#pragma unroll
      for (int k = 0; k < KI; ++k) { const int c = P.intr_col[KI * cm + k]; if (c >= 0) m += e.Jc[NC * r + 6 + k] * P.dc[c]; }
      if (pc >= 0)
#pragma unroll
        for (int k = 0; k < 3; ++k) m += e.Jp[3 * r + k] * P.dp[3 * (int64_t)pc + k];
      acc -= m * (e.r[r] + m / 2.0);
    }
  }
  acc = warp_sum(acc);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += ws[k];
    atomicAdd(out, s);
  }
}

// candidate parameters x + scale * delta; accumulates |step|^2 and |x|^2 in out[0], out[1]
template <int KI>
__global__ void candidate_cameras_kernel(BaDev P, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double st = 0, xs = 0;
  if (i < P.n_img) {
    double d[3] = {0, 0, 0};
    bool any = false;
    for (int k = 0; k < 3; ++k) {
      const int c = P.pose_col[6 * i + k];
      if (c >= 0) { d[k] = P.dc[c] * P.scale_c[c]; any = true; st += d[k] * d[k]; }
    }
    const double* x = P.qvec + 4 * i;
    double* qo = P.qvec_new + 4 * i;
    const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (any && nd > 0.0) {  // QuaternionParameterization::Plus
      const double s = sin(nd) / nd;
      const double q0 = cos(nd), q1 = s * d[0], q2 = s * d[1], q3 = s * d[2];
      qo[0] = q0 * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
      qo[1] = q0 * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
      qo[2] = q0 * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
      qo[3] = q0 * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
    } else {
      for (int k = 0; k < 4; ++k) qo[k] = x[k];
    }
    for (int k = 0; k < 4; ++k) xs += x[k] * x[k];
    for (int k = 0; k < 3; ++k) {
      const int c = P.pose_col[6 * i + 3 + k];
      double v = 0;
      if (c >= 0) { v = P.dc[c] * P.scale_c[c]; st += v * v; }
      P.tvec_new[3 * i + k] = P.tvec[3 * i + k] + v;
      xs += P.tvec[3 * i + k] * P.tvec[3 * i + k];
    }
  }
  if (i < P.n_cam) {
    for (int k = 0; k < KI; ++k) {
      const int c = P.intr_col[KI * i + k];
      double v = 0;
      if (c >= 0) { v = P.dc[c] * P.scale_c[c]; st += v * v; }
      P.cam_new[KI * i + k] = P.cam_params[KI * i + k] + v;
      xs += P.cam_params[KI * i + k] * P.cam_params[KI * i + k];
    }
  }
  st = warp_sum(st);
  xs = warp_sum(xs);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, st);
    atomicAdd(out + 1, xs);
  }
}
__global__ void candidate_points_kernel(BaDev P, double* out) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // coordinate index
  double st = 0, xs = 0;
  if (j < 3 * (int64_t)P.n_pts) {
    const int p = (int)(j / 3), k = (int)(j % 3), pc = P.pt_col[p];
    double v = 0;
    if (pc >= 0) { v = P.dp[3 * (int64_t)pc + k] * P.scale_p[3 * (int64_t)pc + k]; st = v * v; }
    P.xyz_new[j] = P.xyz[j] + v;
    xs = P.xyz[j] * P.xyz[j];
  }
  st = warp_sum(st);
  xs = warp_sum(xs);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out + 2, st);
    atomicAdd(out + 3, xs);
  }
}
__global__ void make_scale_kernel(const double* colnorm, double* scale, int64_t n) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) scale[j] = 1.0 / (1.0 + sqrt(colnorm[j]));
}
__global__ void negate_kernel(double* v, int64_t n) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) v[j] = -v[j];
}

}  // namespace bak

static inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

cudaError_t ba_launch_jacobian(const BaDev& P, const double* q, const double* t, const double* k, const double* X,
                               int mode, double* cost_out, cudaStream_t s, int loss_type, double loss_scale) {
  if (P.n_obs == 0) return cudaSuccess;
  if (P.wide) {  // problems with a camera model beyond the three 4-slot ones
    if (loss_type == 1)
      bak::jacobian_kernel<1, ObsJacW><<<nblk(P.n_obs, 128), 128, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
    else if (loss_type == 2)
      bak::jacobian_kernel<2, ObsJacW><<<nblk(P.n_obs, 128), 128, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
    else
      bak::jacobian_kernel<0, ObsJacW><<<nblk(P.n_obs, 128), 128, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
    return cudaGetLastError();
  }
  if (loss_type == 1)
    bak::jacobian_kernel<1, ObsJac><<<nblk(P.n_obs, 256), 256, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
  else if (loss_type == 2)
    bak::jacobian_kernel<2, ObsJac><<<nblk(P.n_obs, 256), 256, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
  else
    bak::jacobian_kernel<0, ObsJac><<<nblk(P.n_obs, 256), 256, 0, s>>>(P, q, t, k, X, mode, cost_out, loss_scale);
  return cudaGetLastError();
}
cudaError_t ba_launch_camera_terms(const BaDev& P, cudaStream_t s) {
  if (P.n_obs == 0) return cudaSuccess;
  if (P.wide) bak::camera_terms_kernel<ObsJacW><<<nblk(P.n_obs, 256), 256, 0, s>>>(P);
  else bak::camera_terms_kernel<ObsJac><<<nblk(P.n_obs, 256), 256, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t ba_launch_schur(const BaDev& P, double radius, double min_diag, double max_diag, int n_sm, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  const int grid = (int)std::min<int64_t>(P.n_pts, (int64_t)n_sm * 8);
  if (P.wide) bak::schur_kernel<false, ObsJacW><<<grid, bak::kSchurThreads, 0, s>>>(P, radius, min_diag, max_diag, nullptr, nullptr);
  else bak::schur_kernel<false, ObsJac><<<grid, bak::kSchurThreads, 0, s>>>(P, radius, min_diag, max_diag, nullptr, nullptr);
  return cudaGetLastError();
}
// pair-major variant: structure once per solve ...
cudaError_t ba_launch_pm_enumerate(const BaDev& P, int n_img, bool fill, uint32_t* count_or_cursor, const uint32_t* start,
                                   void* tuples, int n_sm, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  const int grid = (int)std::min<int64_t>((P.n_pts + 7) / 8, (int64_t)n_sm * 8);
  if (fill)
    bak::pm_enumerate_kernel<true><<<grid, 256, 0, s>>>(P, n_img, count_or_cursor, start, (int2*)tuples);
  else
    bak::pm_enumerate_kernel<false><<<grid, 256, 0, s>>>(P, n_img, count_or_cursor, start, (int2*)tuples);
  return cudaGetLastError();
}
// ... and per LM iteration: W / Y per observation (point-major), then the block sums
cudaError_t ba_launch_schur_pm(const BaDev& P, double radius, double min_diag, double max_diag, int n_img,
                               const uint32_t* start, const void* tuples, double* Wg, double* Yg, int n_sm,
                               cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  const int grid = (int)std::min<int64_t>(P.n_pts, (int64_t)n_sm * 8);
  bak::schur_kernel<true, ObsJac><<<grid, bak::kSchurThreads, 0, s>>>(P, radius, min_diag, max_diag, Wg, Yg);
  bak::pm_blocks_kernel<<<n_sm * 8, 256, 0, s>>>(P, n_img, start, (const int2*)tuples, Wg, Yg);
  return cudaGetLastError();
}
cudaError_t ba_launch_add_diag(const BaDev& P, double radius, double min_diag, double max_diag, cudaStream_t s) {
  if (P.D == 0) return cudaSuccess;
  bak::add_diag_kernel<<<nblk(P.D, 256), 256, 0, s>>>(P, radius, min_diag, max_diag);
  return cudaGetLastError();
}
cudaError_t ba_launch_backsub(const BaDev& P, cudaStream_t s) {
  if (P.n_pts == 0) return cudaSuccess;
  if (P.wide) bak::backsub_kernel<ObsJacW><<<nblk((int64_t)P.n_pts * 32, 256), 256, 0, s>>>(P);
  else bak::backsub_kernel<ObsJac><<<nblk((int64_t)P.n_pts * 32, 256), 256, 0, s>>>(P);
  return cudaGetLastError();
}
cudaError_t ba_launch_model_cost(const BaDev& P, double* out, cudaStream_t s) {
  if (P.n_obs == 0) return cudaSuccess;
  if (P.wide) bak::model_cost_kernel<ObsJacW><<<nblk(P.n_obs, 256), 256, 0, s>>>(P, out);
  else bak::model_cost_kernel<ObsJac><<<nblk(P.n_obs, 256), 256, 0, s>>>(P, out);
  return cudaGetLastError();
}
cudaError_t ba_launch_candidate(const BaDev& P, double* out, bool cameras, cudaStream_t s) {
  if (cameras) {
    const int n = std::max(P.n_img, P.n_cam);
    if (n > 0 && P.wide) bak::candidate_cameras_kernel<12><<<nblk(n, 128), 128, 0, s>>>(P, out);
    else if (n > 0) bak::candidate_cameras_kernel<4><<<nblk(n, 128), 128, 0, s>>>(P, out);
  } else if (P.n_pts > 0) {
    bak::candidate_points_kernel<<<nblk(3 * (int64_t)P.n_pts, 256), 256, 0, s>>>(P, out);
  }
  return cudaGetLastError();
}
cudaError_t ba_launch_make_scale(const double* colnorm, double* scale, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  bak::make_scale_kernel<<<nblk(n, 256), 256, 0, s>>>(colnorm, scale, n);
  return cudaGetLastError();
}
cudaError_t ba_launch_negate(double* v, int64_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  bak::negate_kernel<<<nblk(n, 256), 256, 0, s>>>(v, n);
  return cudaGetLastError();
}

}  // namespace b2
