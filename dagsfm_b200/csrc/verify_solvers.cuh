// Per-lane FP64 solvers of the two-view verifier (one CUDA thread solves one
// RANSAC hypothesis; 32 hypotheses per warp in flight).
//
// Reference functions restated for the device (file:line under the reference tree):
//   FundamentalMatrixSevenPointEstimator::Estimate   src/estimators/fundamental_matrix.cc:47-142
//   FundamentalMatrixEightPointEstimator::Estimate   src/estimators/fundamental_matrix.cc:150-192
//   EssentialMatrixFivePointEstimator::Estimate      src/estimators/essential_matrix.cc:46-150
//   HomographyMatrixEstimator::Estimate              src/estimators/homography_matrix.cc:44-92
//   CenterAndNormalizeImagePoints                    src/estimators/utils.cc:38-85
//   FindPolynomialRootsCompanionMatrix               src/base/polynomial.cc:208-279
// Eigen's JacobiSVD / PartialPivLU / EigenSolver are replaced by a one-sided Jacobi SVD,
// Gaussian elimination with partial pivoting and a Francis double-shift QR iteration on
// the companion matrix; the 2115 generated lines of essential_matrix_poly.h by a
// table-driven polynomial expansion.  Compiled with --fmad=false: every operation is an
// IEEE-754 double add/mul/div/sqrt exactly as written.
#pragma once
#include <cfloat>
#include <cstdint>

namespace b2 {
namespace vf {

constexpr double kEps = 2.220446049250313e-16;

// The solvers are deliberately __noinline__: the production kernel and the test seam
// (b2_verify_debug_solve) then execute the SAME machine code, so the parity tests of the seam
// cover what production runs.  (Two optimiser-dependent miscompilations of inlined copies were
// observed with nvcc 12.9 -- an insertion sort over a small local array, and a solver inlined
// into a second kernel -- both absent at -G; see DESIGN.md.)

// ------------------------------------------------------------------ one-sided Jacobi
// G: m x n row-major (destroyed), V: n x n row-major out (columns = right singular
// vectors), sig[n]: singular values, both sorted descending (stable).
// Per-lane working arrays live in SHARED memory, element e of lane l at ws[e * STRIDE + l]
// (bank-conflict free across the lanes of a warp); STRIDE == 1 is an ordinary array.
template <int STRIDE>
struct View {
  double* base;
  __device__ __forceinline__ double& operator[](int i) const { return base[i * STRIDE]; }
  __device__ __forceinline__ View operator+(int off) const { return View{base + off * STRIDE}; }
};
constexpr int kLaneWorkDoubles = 72;  // transposed constraint matrix of a minimal solver (<= 9 x 8)

template <int NMAX, typename GV, typename VV>
__device__ __noinline__ void jacobi_svd(GV G, int m, int n, VV V, double* sig) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  // numerically-zero (null space) column pairs are not rotated: alpha * beta would underflow
  double frob2 = 0;
  for (int i = 0; i < m * n; ++i) frob2 += G[i] * G[i];
  const double tiny = frob2 * 1e-40;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < m; ++i) {
          const double gp = G[i * n + p], gq = G[i * n + q];
          alpha += gp * gp;
          beta += gq * gq;
          gamma += gp * gq;
        }
        if (gamma == 0.0 || (alpha <= tiny || beta <= tiny) || fabs(gamma) <= kEps * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double s = c * t;
        for (int i = 0; i < m; ++i) {
          const double gp = G[i * n + p], gq = G[i * n + q];
          G[i * n + p] = c * gp - s * gq;
          G[i * n + q] = s * gp + c * gq;
        }
        for (int i = 0; i < n; ++i) {
          const double wp = V[i * n + p], wq = V[i * n + q];
          V[i * n + p] = c * wp - s * wq;
          V[i * n + q] = s * wp + c * wq;
        }
      }
    }
    if (!rotated) break;
  }
  double nrm[NMAX];
  int order[NMAX];
  for (int j = 0; j < n; ++j) {
    double s = 0;
    for (int i = 0; i < m; ++i) s += G[i * n + j] * G[i * n + j];
    nrm[j] = sqrt(s);
  }
  // stable descending order by rank counting (no data-dependent shifting loops)
  for (int j = 0; j < n; ++j) {
    int rank = 0;
    for (int i = 0; i < n; ++i) rank += (nrm[i] > nrm[j] || (nrm[i] == nrm[j] && i < j)) ? 1 : 0;
    order[rank] = j;
  }
  double tmp[NMAX * NMAX];
  for (int i = 0; i < n * n; ++i) tmp[i] = V[i];
  for (int j = 0; j < n; ++j) {
    const int o = order[j];
    sig[j] = nrm[o];
    for (int i = 0; i < n; ++i) V[i * n + j] = tmp[i * n + o];
  }
  if (m <= NMAX) {  // small case: also permute G (= U * Sigma), used by svd3_rebuild
    double tmpg[NMAX * NMAX];
    for (int i = 0; i < m * n; ++i) tmpg[i] = G[i];
    for (int j = 0; j < n; ++j) {
      const int o = order[j];
      for (int i = 0; i < m; ++i) G[i * n + j] = tmpg[i * n + o];
    }
  }
}

// Orthonormal basis of the null space of an M x 9 constraint matrix A (M < 9): Householder QR of
// B = A^T (9 x M, element (r, c) at B[r * M + c], destroyed); the last 9 - M columns of Q span
// null(A) whatever the rank of A.  The reference takes the same subspace from the trailing columns
// of Eigen::JacobiSVD's V (fundamental_matrix.cc:55-60, homography_matrix.cc:84-90,
// essential_matrix.cc:80-84); the models derived from it are normalised afterwards (F / F(2,2),
// E / |E|, H homogeneous), so they do not depend on the basis, and the QR costs ~1/50 of a Jacobi
// SVD per lane.  Everything is unrolled: the workspace offsets are immediates.
template <int M, typename BV>
__device__ __forceinline__ void null_space_qr(BV B, double* out) {
  double beta[M];
#pragma unroll
  for (int k = 0; k < M; ++k) {
    const double x0 = B[k * M + k];
    double s = 0;
#pragma unroll
    for (int r = k + 1; r < 9; ++r) s += B[r * M + k] * B[r * M + k];
    const double nrm = sqrt(x0 * x0 + s);
    if (nrm == 0.0) {
      beta[k] = 0.0;
      continue;
    }
    const double v0 = x0 + (x0 >= 0 ? nrm : -nrm);
    B[k * M + k] = v0;
    beta[k] = 2.0 / (v0 * v0 + s);
#pragma unroll
    for (int j = k + 1; j < M; ++j) {
      double w = v0 * B[k * M + j];
#pragma unroll
      for (int r = k + 1; r < 9; ++r) w += B[r * M + k] * B[r * M + j];
      w *= beta[k];
      B[k * M + j] -= w * v0;
#pragma unroll
      for (int r = k + 1; r < 9; ++r) B[r * M + j] -= w * B[r * M + k];
    }
  }
#pragma unroll
  for (int j = M; j < 9; ++j) {
    double q[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) q[r] = (r == j) ? 1.0 : 0.0;
#pragma unroll
    for (int k = M - 1; k >= 0; --k) {
      double w = 0;
#pragma unroll
      for (int r = k; r < 9; ++r) w += B[r * M + k] * q[r];
      w *= beta[k];
#pragma unroll
      for (int r = k; r < 9; ++r) q[r] -= w * B[r * M + k];
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) out[(j - M) * 9 + r] = q[r];
  }
}

// R = U diag(new sigma) V^T for a 3x3 A (row-major); mode 0: sigma2 = 0 (F), 1: also
// sigma0 = sigma1 = mean (E).
__device__ inline void svd3_rebuild(const double* A, int mode, double* R) {
  double G[9], V[9], sig[3];
  for (int i = 0; i < 9; ++i) G[i] = A[i];
  jacobi_svd<3>(View<1>{G}, 3, 3, View<1>{V}, sig);
  double sn[3] = {sig[0], sig[1], 0.0};
  if (mode == 1) {
    sn[0] = (sig[0] + sig[1]) / 2.0;
    sn[1] = sn[0];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) {
        if (sn[k] == 0.0 || sig[k] == 0.0) continue;
        s += (G[i * 3 + k] / sig[k]) * sn[k] * V[j * 3 + k];
      }
      R[i * 3 + j] = s;
    }
}

__device__ inline void mat3_mul(const double* a, const double* b, double* r) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
      r[3 * i + j] = s;
    }
}
__device__ inline void mat3_transpose(const double* a, double* r) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r[3 * i + j] = a[3 * j + i];
}
__device__ inline void mat3_inverse(const double* a, double* r) {
  const double c00 = a[4] * a[8] - a[5] * a[7];
  const double c01 = a[5] * a[6] - a[3] * a[8];
  const double c02 = a[3] * a[7] - a[4] * a[6];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double id = 1.0 / det;
  r[0] = c00 * id;
  r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
  r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
  r[3] = c01 * id;
  r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
  r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
  r[6] = c02 * id;
  r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
  r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// ------------------------------------------------ companion-matrix eigenvalues (hqr)
// a: n x n row-major upper Hessenberg (destroyed).  Returns false on non-convergence.
inline __device__ __noinline__ bool hqr(double* a, int n, double* wr, double* wi) {
#define HA(i, j) a[(i) * n + (j)]
  int nn = n - 1;
  double t = 0.0, p = 0, q = 0, r = 0, s = 0, w, x, y, z;
  double anorm = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(HA(i, j));
  while (nn >= 0) {
    int its = 0, l;
    do {
      for (l = nn; l >= 1; --l) {
        s = fabs(HA(l - 1, l - 1)) + fabs(HA(l, l));
        if (s == 0.0) s = anorm;
        if (fabs(HA(l, l - 1)) + s == s) {
          HA(l, l - 1) = 0.0;
          break;
        }
      }
      x = HA(nn, nn);
      if (l == nn) {
        wr[nn] = x + t;
        wi[nn--] = 0.0;
      } else {
        y = HA(nn - 1, nn - 1);
        w = HA(nn, nn - 1) * HA(nn - 1, nn);
        if (l == nn - 1) {
          p = 0.5 * (y - x);
          q = p * p + w;
          z = sqrt(fabs(q));
          x += t;
          if (q >= 0.0) {
            z = p + (p >= 0 ? fabs(z) : -fabs(z));
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn] = z;
            wi[nn - 1] = -z;
          }
          nn -= 2;
        } else {
          if (its == 60) return false;
          if (its == 10 || its == 20) {
            t += x;
            for (int i = 0; i <= nn; ++i) HA(i, i) -= x;
            s = fabs(HA(nn, nn - 1)) + fabs(HA(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          int m;
          for (m = nn - 2; m >= l; --m) {
            z = HA(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / HA(m + 1, m) + HA(m, m + 1);
            q = HA(m + 1, m + 1) - z - r - s;
            r = HA(m + 2, m + 1);
            s = fabs(p) + fabs(q) + fabs(r);
            p /= s;
            q /= s;
            r /= s;
            if (m == l) break;
            const double u = fabs(HA(m, m - 1)) * (fabs(q) + fabs(r));
            const double v = fabs(p) * (fabs(HA(m - 1, m - 1)) + fabs(z) + fabs(HA(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (int i = m + 2; i <= nn; ++i) {
            HA(i, i - 2) = 0.0;
            if (i != m + 2) HA(i, i - 3) = 0.0;
          }
          for (int k = m; k <= nn - 1; ++k) {
            if (k != m) {
              p = HA(k, k - 1);
              q = HA(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = HA(k + 2, k - 1);
              if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) {
                p /= x;
                q /= x;
                r /= x;
              }
            }
            const double sg = sqrt(p * p + q * q + r * r);
            s = (p >= 0 ? sg : -sg);
            if (s != 0.0) {
              if (k == m) {
                if (l != m) HA(k, k - 1) = -HA(k, k - 1);
              } else {
                HA(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s;
              y = q / s;
              z = r / s;
              q /= p;
              r /= p;
              for (int j = k; j <= nn; ++j) {
                p = HA(k, j) + q * HA(k + 1, j);
                if (k != nn - 1) {
                  p += r * HA(k + 2, j);
                  HA(k + 2, j) -= p * z;
                }
                HA(k + 1, j) -= p * y;
                HA(k, j) -= p * x;
              }
              const int mmin = nn < k + 3 ? nn : k + 3;
              for (int i = l; i <= mmin; ++i) {
                p = x * HA(i, k) + y * HA(i, k + 1);
                if (k != nn - 1) {
                  p += z * HA(i, k + 2);
                  HA(i, k + 2) -= p * r;
                }
                HA(i, k + 1) -= p * q;
                HA(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
#undef HA
  return true;
}

// Real roots (|imag| <= 1e-10) of the polynomial coeffs[0..nc) (highest power first),
// ascending.  Follows polynomial.cc:208-279 incl. leading / trailing zero handling.
// Returns the number of real roots, or -1 if the reference would return false.
template <int MAXDEG>
__device__ __noinline__ int real_roots(const double* coeffs_all, int nc, double* roots) {
  int lead = 0;
  while (lead < nc && coeffs_all[lead] == 0) ++lead;
  const double* coeffs = coeffs_all + lead;
  int n = nc - lead;  // number of coefficients
  const int degree = n - 1;
  if (degree <= 0) return -1;
  if (degree == 1) {
    roots[0] = -coeffs[1] / coeffs[0];
    return 1;
  }
  if (degree == 2) {
    const double a = coeffs[0], b = coeffs[1], c = coeffs[2];
    if (b == 0 && c == 0) {
      roots[0] = 0;
      return 1;
    }
    const double d = b * b - 4 * a * c;
    if (d >= 0) {
      const double sd = sqrt(d);
      double r0, r1;
      if (b >= 0) {
        r0 = (-b - sd) / (2 * a);
        r1 = (2 * c) / (-b - sd);
      } else {
        r0 = (2 * c) / (-b + sd);
        r1 = (-b + sd) / (2 * a);
      }
      roots[0] = r0 < r1 ? r0 : r1;
      roots[1] = r0 < r1 ? r1 : r0;
      return 2;
    }
    if (fabs(sqrt(-d) / (2 * a)) <= 1e-10) {
      roots[0] = roots[1] = -b / (2 * a);
      return 2;
    }
    return 0;
  }
  int trail = 0;
  while (trail < n && coeffs[n - 1 - trail] == 0) ++trail;
  n -= trail;
  if (n == 1) {
    roots[0] = 0;
    return 1;
  }
  const int d = n - 1;
  double C[MAXDEG * MAXDEG], wr[MAXDEG], wi[MAXDEG];
  for (int i = 0; i < d * d; ++i) C[i] = 0.0;
  for (int i = 1; i < d; ++i) C[i * d + i - 1] = 1;
  for (int j = 0; j < d; ++j) C[j] = -coeffs[j + 1] / coeffs[0];
  if (!hqr(C, d, wr, wi)) return -1;
  int nr = 0;
  for (int k = 0; k < d; ++k)
    if (fabs(wi[k]) <= 1e-10) roots[nr++] = wr[k];
  if (trail > 0) roots[nr++] = 0.0;
  {  // ascending, by rank counting
    double sorted[MAXDEG + 1];
    for (int j = 0; j < nr; ++j) {
      int rank = 0;
      for (int i = 0; i < nr; ++i) rank += (roots[i] < roots[j] || (roots[i] == roots[j] && i < j)) ? 1 : 0;
      sorted[rank] = roots[j];
    }
    for (int j = 0; j < nr; ++j) roots[j] = sorted[j];
  }
  return nr;
}

// Hypotheses of one sample in a basis- and sign-independent order (oracle: CanonicalOrder): the
// reference's order is an artefact of Eigen's eigenvalue deflation and SVD null-space basis.
// Rank-counting (no data-dependent shifting loops, see the toolchain note in DESIGN.md).
__device__ inline void canonical_order(double* models, int nm) {
  if (nm < 2) return;
  double key[10], tmp[90];
  for (int j = 0; j < nm; ++j) {
    double s = 0;
    for (int k = 0; k < 9; ++k) {
      tmp[9 * j + k] = models[9 * j + k];
      s += (double)(k + 1) * (models[9 * j + k] * models[9 * j + k]);
    }
    key[j] = s;
  }
  for (int j = 0; j < nm; ++j) {
    int rank = 0;
    for (int i = 0; i < nm; ++i) rank += (key[i] < key[j] || (key[i] == key[j] && i < j)) ? 1 : 0;
    for (int k = 0; k < 9; ++k) models[9 * rank + k] = tmp[9 * j + k];
  }
}

// ------------------------------------------------------------------ F 7-point
// p1, p2: 7 points each (x,y interleaved).  models: up to 3 x 9.  Returns count.
template <int STRIDE>
__device__ __noinline__ int solve_f7(View<STRIDE> ws, const double* p1, const double* p2, double* models) {
  for (int i = 0; i < 7; ++i) {
    const double x0 = p1[2 * i], y0 = p1[2 * i + 1], x1 = p2[2 * i], y1 = p2[2 * i + 1];
    const double a[9] = {x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, 1};
    for (int r = 0; r < 9; ++r) ws[r * 7 + i] = a[r];
  }
  double nv[18];
  null_space_qr<7>(ws, nv);
  double f1[9], f2[9];
  for (int i = 0; i < 9; ++i) {
    f2[i] = nv[9 + i];
    f1[i] = nv[i] - f2[i];
  }
  const double t0 = f1[4] * f1[8] - f1[5] * f1[7];
  const double t1 = f1[3] * f1[8] - f1[5] * f1[6];
  const double t2 = f1[3] * f1[7] - f1[4] * f1[6];
  const double t3 = f2[4] * f2[8] - f2[5] * f2[7];
  const double t4 = f2[3] * f2[8] - f2[5] * f2[6];
  const double t5 = f2[3] * f2[7] - f2[4] * f2[6];
  double c[4];
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  c[2] = f1[0] * t3 - f1[1] * t4 + f1[2] * t5 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  c[3] = f2[0] * t3 - f2[1] * t4 + f2[2] * t5;
  double roots[3];
  const int nr = real_roots<3>(c, 4, roots);
  int nm = 0;
  for (int i = 0; i < nr; ++i) {
    const double lambda = roots[i];
    double F[9];
    for (int k = 0; k < 9; ++k) F[k] = lambda * f1[k] + 1 * f2[k];
    if (fabs(F[8]) < 1e-10) continue;
    for (int k = 0; k < 9; ++k) models[9 * nm + k] = F[k] / F[8];
    ++nm;
  }
  canonical_order(models, nm);
  return nm;
}

// ---------------------------------------------------------- Hartley normalisation
// utils.cc:38-85 for n points (x,y interleaved), sequential sums in index order.
__device__ inline void center_and_normalize(const double* p, int n, double* np, double* T) {
  double cx = 0, cy = 0;
  for (int i = 0; i < n; ++i) {
    cx += p[2 * i];
    cy += p[2 * i + 1];
  }
  cx /= n;
  cy /= n;
  double rms = 0;
  for (int i = 0; i < n; ++i) {
    const double dx = p[2 * i] - cx, dy = p[2 * i + 1] - cy;
    rms += dx * dx + dy * dy;
  }
  rms = sqrt(rms / n);
  const double nf = sqrt(2.0) / rms;
  T[0] = nf; T[1] = 0; T[2] = -nf * cx; T[3] = 0; T[4] = nf; T[5] = -nf * cy; T[6] = 0; T[7] = 0; T[8] = 1;
  for (int i = 0; i < n; ++i) {
    const double p0 = p[2 * i], p1 = p[2 * i + 1];
    const double n0 = T[0] * p0 + T[1] * p1 + T[2];
    const double n1 = T[3] * p0 + T[4] * p1 + T[5];
    const double n2 = T[6] * p0 + T[7] * p1 + T[8];
    const double inv = 1.0 / n2;
    np[2 * i] = n0 * inv;
    np[2 * i + 1] = n1 * inv;
  }
}

// ------------------------------------------------------------------ H 4-point
template <int STRIDE>
__device__ __noinline__ int solve_h4(View<STRIDE> ws, const double* p1, const double* p2, double* model) {
  double n1[8], n2[8], T1[9], T2[9];
  center_and_normalize(p1, 4, n1, T1);
  center_and_normalize(p2, 4, n2, T2);
  for (int i = 0; i < 72; ++i) ws[i] = 0.0;
  for (int i = 0, j = 4; i < 4; ++i, ++j) {
    const double s_0 = n1[2 * i], s_1 = n1[2 * i + 1], d_0 = n2[2 * i], d_1 = n2[2 * i + 1];
    // equation i: (-s, -1, 0, 0, 0, s d_0, d_0); equation j: (0, 0, 0, -s, -1, s d_1, d_1); stored transposed
    ws[0 * 8 + i] = -s_0; ws[1 * 8 + i] = -s_1; ws[2 * 8 + i] = -1;
    ws[6 * 8 + i] = s_0 * d_0; ws[7 * 8 + i] = s_1 * d_0; ws[8 * 8 + i] = d_0;
    ws[3 * 8 + j] = -s_0; ws[4 * 8 + j] = -s_1; ws[5 * 8 + j] = -1;
    ws[6 * 8 + j] = s_0 * d_1; ws[7 * 8 + j] = s_1 * d_1; ws[8 * 8 + j] = d_1;
  }
  double Ht[9], T2i[9], tmp[9];
  null_space_qr<8>(ws, Ht);
  mat3_inverse(T2, T2i);
  mat3_mul(T2i, Ht, tmp);
  mat3_mul(tmp, T1, model);
  return 1;
}

// Finishes the 8-point / DLT local estimators from the null vector of the (normalised)
// constraint matrix: kind 0 = F (rank 2), 1 = H.
__device__ inline void finish_f8(const double* nullvec, const double* T1, const double* T2, double* model) {
  double F[9], T2t[9], tmp[9];
  svd3_rebuild(nullvec, 0, F);
  mat3_transpose(T2, T2t);
  mat3_mul(T2t, F, tmp);
  mat3_mul(tmp, T1, model);
}
__device__ inline void finish_h(const double* nullvec, const double* T1, const double* T2, double* model) {
  double T2i[9], tmp[9];
  mat3_inverse(T2, T2i);
  mat3_mul(T2i, nullvec, tmp);
  mat3_mul(tmp, T1, model);
}

// ------------------------------------------------------------------ E 5-point
// Monomial index tables (generated): linear [x,y,z,1]; quadratic
// [x2,y2,z2,xy,xz,yz,x,y,z,1]; cubic in the Nister / Stewenius column order
// x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1.
__device__ __constant__ unsigned char kQI[4][4] = {{0, 3, 4, 6}, {3, 1, 5, 7}, {4, 5, 2, 8}, {6, 7, 8, 9}};
__device__ __constant__ unsigned char kCI[10][4] = {{0, 2, 4, 5},   {3, 1, 6, 7},    {10, 13, 16, 17}, {2, 3, 8, 9},
                                                    {4, 8, 10, 11}, {8, 6, 13, 14},  {5, 9, 11, 12},   {9, 7, 14, 15},
                                                    {11, 14, 17, 18}, {12, 15, 18, 19}};

__device__ inline void lin_mul_acc(double* q, const double* a, const double* b, double sgn) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) q[kQI[i][j]] += sgn * (a[i] * b[j]);
}
__device__ inline void quad_mul_acc(double* c, const double* q, const double* l, double sgn) {
  for (int i = 0; i < 10; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[kCI[i][j]] += sgn * (q[i] * l[j]);
}

// Eb: 4 basis vectors X,Y,Z,W (each 9, row-major 3x3).  models: up to 10 x 9.
inline __device__ __noinline__ int solve_e5_from_basis(const double* Eb, double* models) {
  // E(r,c) as a linear polynomial [x,y,z,1]
  double L[9][4];
  for (int i = 0; i < 9; ++i) {
    L[i][0] = Eb[i];
    L[i][1] = Eb[9 + i];
    L[i][2] = Eb[18 + i];
    L[i][3] = Eb[27 + i];
  }
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int c = 0; c < 20; ++c) A[r][c] = 0.0;
  {  // row 0: det(E)
    const int cof[3][4] = {{4, 8, 5, 7}, {3, 8, 5, 6}, {3, 7, 4, 6}};
    for (int k = 0; k < 3; ++k) {
      double q[10];
      for (int i = 0; i < 10; ++i) q[i] = 0.0;
      lin_mul_acc(q, L[cof[k][0]], L[cof[k][1]], 1.0);
      lin_mul_acc(q, L[cof[k][2]], L[cof[k][3]], -1.0);
      quad_mul_acc(A[0], q, L[k], (k == 1) ? -1.0 : 1.0);
    }
  }
  {  // rows 1..9: E E^T E - 0.5 trace(E E^T) E
    double EEt[6][10];  // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    const int pr[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    for (int e = 0; e < 6; ++e) {
      for (int i = 0; i < 10; ++i) EEt[e][i] = 0.0;
      for (int k = 0; k < 3; ++k) lin_mul_acc(EEt[e], L[3 * pr[e][0] + k], L[3 * pr[e][1] + k], 1.0);
    }
    double trh[10];
    for (int i = 0; i < 10; ++i) trh[i] = 0.5 * ((EEt[0][i] + EEt[3][i]) + EEt[5][i]);
    const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        double* row = A[1 + 3 * r + c];
        for (int k = 0; k < 3; ++k) quad_mul_acc(row, EEt[sym[r][k]], L[3 * k + c], 1.0);
        quad_mul_acc(row, trh, L[3 * r + c], -1.0);
      }
  }
  // Gaussian elimination with partial pivoting, then back-substitute only the six rows
  // (4..9) of A1^{-1} A2 that the method uses -- all ten are needed for back-substitution.
  for (int col = 0; col < 10; ++col) {
    int piv = col;
    for (int r = col + 1; r < 10; ++r)
      if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
    if (A[piv][col] == 0.0) return 0;
    if (piv != col)
      for (int c = 0; c < 20; ++c) {
        const double t = A[piv][c];
        A[piv][c] = A[col][c];
        A[col][c] = t;
      }
    for (int r = col + 1; r < 10; ++r) {
      const double f = A[r][col] / A[col][col];
      if (f == 0.0) continue;
      for (int c = col; c < 20; ++c) A[r][c] -= f * A[col][c];
    }
  }
  double AA[10][10];
  for (int c = 0; c < 10; ++c)
    for (int r = 9; r >= 0; --r) {
      double s = A[r][10 + c];
      for (int k = r + 1; k < 10; ++k) s -= A[r][k] * AA[k][c];
      AA[r][c] = s / A[r][r];
    }
  double B[13][3];
  for (int i = 0; i < 3; ++i) {
    B[0][i] = 0; B[4][i] = 0; B[8][i] = 0;
    for (int k = 0; k < 3; ++k) B[1 + k][i] = AA[i * 2 + 4][k];
    for (int k = 0; k < 3; ++k) B[5 + k][i] = AA[i * 2 + 4][3 + k];
    for (int k = 0; k < 4; ++k) B[9 + k][i] = AA[i * 2 + 4][6 + k];
    for (int k = 0; k < 3; ++k) B[0 + k][i] -= AA[i * 2 + 5][k];
    for (int k = 0; k < 3; ++k) B[4 + k][i] -= AA[i * 2 + 5][3 + k];
    for (int k = 0; k < 4; ++k) B[8 + k][i] -= AA[i * 2 + 5][6 + k];
  }
  // det of the polynomial matrix: sum over rows (r,s,u) of sign * (b0_r b1_s - b0_s b1_r) * b2_u
  double det[11];
  for (int i = 0; i < 11; ++i) det[i] = 0.0;
  const int perm[3][3] = {{1, 2, 0}, {0, 2, 1}, {0, 1, 2}};
  for (int tix = 0; tix < 3; ++tix) {
    const int r = perm[tix][0], s = perm[tix][1], u = perm[tix][2];
    double mn[7];
    for (int i = 0; i < 7; ++i) mn[i] = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) mn[i + j] += B[i][r] * B[4 + j][s];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) mn[i + j] -= B[i][s] * B[4 + j][r];
    const double sgn = (tix == 1) ? -1.0 : 1.0;
    for (int i = 0; i < 7; ++i)
      for (int j = 0; j < 5; ++j) det[i + j] += sgn * (mn[i] * B[8 + j][u]);
  }
  double roots[10];
  const int nr = real_roots<10>(det, 11, roots);
  int nm = 0;
  for (int i = 0; i < nr; ++i) {
    const double z1 = roots[i], z2 = z1 * z1, z3 = z2 * z1, z4 = z3 * z1;
    double Bz[9], V3[9], s3[3];
    for (int j = 0; j < 3; ++j) {
      Bz[3 * j + 0] = B[0][j] * z3 + B[1][j] * z2 + B[2][j] * z1 + B[3][j];
      Bz[3 * j + 1] = B[4][j] * z3 + B[5][j] * z2 + B[6][j] * z1 + B[7][j];
      Bz[3 * j + 2] = B[8][j] * z4 + B[9][j] * z3 + B[10][j] * z2 + B[11][j] * z1 + B[12][j];
    }
    jacobi_svd<3>(View<1>{Bz}, 3, 3, View<1>{V3}, s3);
    const double X0 = V3[2], X1 = V3[5], X2 = V3[8];
    if (fabs(X2) < 1e-10) continue;
    double ev[9], nrm = 0;
    for (int k = 0; k < 9; ++k) {
      ev[k] = Eb[k] * (X0 / X2) + Eb[9 + k] * (X1 / X2) + Eb[18 + k] * z1 + Eb[27 + k];
      nrm += ev[k] * ev[k];
    }
    nrm = sqrt(nrm);
    for (int k = 0; k < 9; ++k) models[9 * nm + k] = ev[k] / nrm;
    ++nm;
  }
  canonical_order(models, nm);
  return nm;
}

// Minimal 5-point: 5 x 9 epipolar constraint matrix -> 4-dim null space -> models.
template <int STRIDE>
__device__ __noinline__ int solve_e5(View<STRIDE> ws, const double* p1, const double* p2, double* models) {
  for (int i = 0; i < 5; ++i) {
    const double x1_0 = p1[2 * i], x1_1 = p1[2 * i + 1], x2_0 = p2[2 * i], x2_1 = p2[2 * i + 1];
    const double q[9] = {x1_0 * x2_0, x1_1 * x2_0, x2_0, x1_0 * x2_1, x1_1 * x2_1, x2_1, x1_0, x1_1, 1};
    for (int r = 0; r < 9; ++r) ws[r * 5 + i] = q[r];
  }
  double Eb[36];
  null_space_qr<5>(ws, Eb);
  return solve_e5_from_basis(Eb, models);
}

}  // namespace vf
}  // namespace b2
