// Shared declarations of the bundle adjuster (device-side problem view).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

namespace b2 {

// Per-observation residual and Jacobi-scaled Jacobian blocks (224 bytes): the layout of the three camera models with
// at most four parameters (SIMPLE_PINHOLE, PINHOLE, SIMPLE_RADIAL) -- the reference's defaults and the fast path.
struct ObsJac {
  static constexpr int kKI = 4;    // intrinsics slots per camera
  static constexpr int kNC = 10;   // camera-side columns per observation
  double r[2];
  double Jc[20];  // 2 x 10: [rotation 3 | translation 3 | intrinsics 4]
  double Jp[6];   // 2 x 3
};
// The wide layout (352 bytes) for problems that contain any other model of camera_models.h (up to 12 parameters:
// FULL_OPENCV, THIN_PRISM_FISHEYE).  Every kernel that touches Jacobian blocks is a template over the layout; the
// 4-slot instantiations are the production kernels, unchanged.
struct ObsJacW {
  static constexpr int kKI = 12;
  static constexpr int kNC = 18;
  double r[2];
  double Jc[36];  // 2 x 18: [rotation 3 | translation 3 | intrinsics 12]
  double Jp[6];
};

struct BaDev {
  int32_t n_img, n_cam, n_pts;
  int64_t n_obs;
  int64_t D;  // reduced camera system dimension
  // problem (SoA over observations sorted by point)
  const int32_t* obs_img;
  const int32_t* obs_pt;
  const double2* obs_xy;
  const int64_t* pt_start;
  const int32_t* img_cam;
  const int32_t* cam_model;
  const int32_t* pose_col;  // [n_img*6]
  const int32_t* intr_col;  // [n_cam*4]
  const int32_t* pt_col;    // [n_pts] index of the variable point or -1
  // parameters and candidates
  double *qvec, *tvec, *cam_params, *xyz;
  double *qvec_new, *tvec_new, *cam_new, *xyz_new;
  // Jacobian / normal equations
  ObsJac* J;
  double *scale_c, *scale_p, *colnorm_c, *colnorm_p;
  double* S;       // [D*D]   | these four are contiguous: one all-reduce
  double* rhs;     // [D]     |
  double* g_c;     // [D]     |
  double* diag_c;  // [D]     |
  double *diag_p, *g_p, *Vinv;  // per variable point
  double *dc, *dp;              // step
  double* gmax;                 // gradient max norm (bit pattern of a non-negative double)
  // (appended: the offsets above are what the production kernels were compiled and validated with)
  ObsJacW* JW;                  // Jacobian blocks in the wide layout, when wide != 0 (then J == nullptr)
  int32_t wide;                 // 0: ObsJac, intrinsics stride 4; 1: ObsJacW, stride 12 (cam_params, cam_new, intr_col)
  int32_t reserved_;
  const int32_t* rot_img;       // [D] image whose rotation block STARTS at this column, -2 for its 2nd / 3rd column, else -1
};

// Column j's term of Ceres' gradient_max_norm = |x - Plus(x, -g)|_inf (trust_region_minimizer.cc, Ceres 1.14, external;
// g = the gradient in the tangent space of the UNSCALED problem).  Plain |g_j| for tvec / intrinsics columns (identity
// or subset parameterisations); for a rotation block the quaternion moves by QuaternionParameterization::Plus, which is
// not linear in g: the three columns are handled together by the block's first column.
__device__ __forceinline__ double grad_norm_term(const BaDev& P, int64_t j) {
  const int ri = P.rot_img[j];
  if (ri == -1) return fabs(P.g_c[j] / P.scale_c[j]);
  if (ri < 0) return 0.0;
  const double d0 = -P.g_c[j] / P.scale_c[j], d1 = -P.g_c[j + 1] / P.scale_c[j + 1], d2 = -P.g_c[j + 2] / P.scale_c[j + 2];
  const double nd = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
  if (!(nd > 0.0)) return 0.0;
  const double* x = P.qvec + 4 * (int64_t)ri;
  const double sn = sin(nd) / nd;
  const double q0 = cos(nd), q1 = sn * d0, q2 = sn * d1, q3 = sn * d2;
  const double y0 = q0 * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
  const double y1 = q0 * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
  const double y2 = q0 * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
  const double y3 = q0 * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
  return fmax(fmax(fabs(x[0] - y0), fabs(x[1] - y1)), fmax(fabs(x[2] - y2), fabs(x[3] - y3)));
}
template <class J> __host__ __device__ inline J* jac(const BaDev& P);
template <> __host__ __device__ inline ObsJac* jac<ObsJac>(const BaDev& P) { return P.J; }
template <> __host__ __device__ inline ObsJacW* jac<ObsJacW>(const BaDev& P) { return P.JW; }

cudaError_t ba_launch_jacobian(const BaDev& P, const double* q, const double* t, const double* k, const double* X,
                               int mode, double* cost_out, cudaStream_t s, int loss_type = 0, double loss_scale = 1.0);
cudaError_t ba_launch_camera_terms(const BaDev& P, cudaStream_t s);
cudaError_t ba_launch_schur(const BaDev& P, double radius, double min_diag, double max_diag, int n_sm, cudaStream_t s);
cudaError_t ba_launch_pm_enumerate(const BaDev& P, int n_img, bool fill, uint32_t* count_or_cursor, const uint32_t* start,
                                   void* tuples, int n_sm, cudaStream_t s);
cudaError_t ba_launch_schur_pm(const BaDev& P, double radius, double min_diag, double max_diag, int n_img,
                               const uint32_t* start, const void* tuples, double* Wg, double* Yg, int n_sm,
                               cudaStream_t s);
cudaError_t ba_launch_add_diag(const BaDev& P, double radius, double min_diag, double max_diag, cudaStream_t s);
cudaError_t ba_launch_backsub(const BaDev& P, cudaStream_t s);
cudaError_t ba_launch_model_cost(const BaDev& P, double* out, cudaStream_t s);
cudaError_t ba_launch_candidate(const BaDev& P, double* out, bool cameras, cudaStream_t s);
cudaError_t ba_launch_make_scale(const double* colnorm, double* scale, int64_t n, cudaStream_t s);
cudaError_t ba_launch_negate(double* v, int64_t n, cudaStream_t s);

// ---- result metrics (ba_metrics.cu)
cudaError_t bam_launch_reprojection_errors(int n_pts, const int64_t* pt_start, const int32_t* obs_img, const double* obs_xy,
                                           const int32_t* img_cam, const int32_t* cam_model, const double* cam_params,
                                           int cam_stride, const double* qvec, const double* tvec, const double* xyz,
                                           double* point_error, double* partial_sum, int* n_blocks, cudaStream_t s);

// ---- ITERATIVE_SCHUR (ba_iterative.cu): state of the conjugate-gradient solve of the reduced system
constexpr int kBaIterMaxPartials = 128;  // blocks of the vector kernels = partial sums per dot product
struct BaIter {
  const int64_t* img_start;  // [n_img + 1] observations grouped by image ...
  const int32_t* img_obs;    // [n_obs]     ... as indices into the point-major arrays
  const int32_t* blk_first;  // [D] first column of the column's parameter block (rotation / tvec / intrinsics)
  const int32_t* blk_size;   // [D] its number of columns (<= 4)
  double* tp;                // [NP * 3] V^-1 g_p
  double* zp;                // [NP * 3] V^-1 E'F x of the current product
  double* lm_c;              // [D] LM diagonal of the camera columns
  double* M;                 // [D * m_stride] block diagonal of S, then its inverse (SCHUR_JACOBI)
  int* flag;                 // raised by precond_invert_kernel on a block that is not positive definite
  int32_t m_stride;          // 4, or 12 in the wide layout (largest parameter block = a camera's intrinsics)
};
cudaError_t bai_launch_point_prepare(const BaDev& P, const BaIter& I, double radius, double min_diag, double max_diag, cudaStream_t s);
cudaError_t bai_launch_rhs(const BaDev& P, const BaIter& I, cudaStream_t s);
cudaError_t bai_launch_camera_terms_image(const BaDev& P, const BaIter& I, cudaStream_t s);
cudaError_t bai_launch_cam_diag(const BaDev& P, const BaIter& I, double radius, double min_diag, double max_diag, cudaStream_t s);
cudaError_t bai_launch_precond(const BaDev& P, const BaIter& I, cudaStream_t s);
cudaError_t bai_launch_precond_invert(const BaDev& P, const BaIter& I, cudaStream_t s);
cudaError_t bai_launch_matvec(const BaDev& P, const BaIter& I, const double* x, double* out, cudaStream_t s);
cudaError_t bai_launch_dot(int64_t D, const double* a, const double* b, double* partial, int* n_partial, cudaStream_t s);
cudaError_t bai_launch_cg_precond(int64_t D, const BaIter& I, const double* r, double* z, double* partial, int* n_partial, cudaStream_t s);
cudaError_t bai_launch_cg_update_p(int64_t D, const double* z, double* p, double beta, bool first, cudaStream_t s);
cudaError_t bai_launch_cg_finish_q(int64_t D, const double* lm_c, const double* p, double* q, double* partial, int* n_partial, cudaStream_t s);
cudaError_t bai_launch_cg_update_xr(int64_t D, double* x, const double* p, double* r, const double* q, const double* b,
                                    double alpha, int mode, double* partial, int* n_partial, cudaStream_t s);

}  // namespace b2

// =====================================================================================================================
// Fused exact path (ba_fused.cu + ba_chol.cu): no staged Jacobian blocks, the reduced camera system S in packed 64 x 64
// tiles (upper block triangle, only the tiles the camera graph + Cholesky fill make non-zero), a hand-written tiled
// Cholesky.  Selected by b2_ba_solve for the 4-slot camera models when every variable point's track fits a window
// (<= kWinMaxLoc images, no image twice); everything else keeps the staged path above.
namespace b2 {

constexpr int kST = 64;            // tile side of the packed reduced system
constexpr int kWinMaxLoc = 16;     // images per accumulation window (schur_window_kernel<12> / <16>)
constexpr int kWinBatch = 12;      // points per k-panel of the window product

struct BaTiles {
  int32_t nt;               // tiles per side = ceil(D / 64)
  int32_t n_tiles;          // structurally non-zero upper tiles (after symbolic fill)
  const int32_t* tile_id;   // [nt * nt] packed index of tile (ti, tj), ti <= tj, or -1
  const int32_t* row_ptr;   // [nt + 1] CSR of the upper tiles by tile row, diagonal tile first
  const int32_t* row_col;   // [n_tiles] tile column; a tile's packed index is its CSR position
  double* tiles;            // [n_tiles][64 * 64] row-major
  double* rinv;             // [nt][64 * 64] inverse of each factored diagonal tile (upper triangular)
  int* info;                // != 0: a pivot was not positive
};
__host__ __device__ inline double* tile_entry(const BaTiles& T, int r, int c) {  // r <= c, the tile must exist
  const int id = T.tile_id[(r >> 6) * T.nt + (c >> 6)];
  return T.tiles + (size_t)id * (kST * kST) + (r & 63) * kST + (c & 63);
}

struct BaWin {
  int32_t n_chunks;
  int32_t nloc;               // 12 or 16
  const int32_t* chunk_pt0;   // [n_chunks + 1] ranges of pt_order
  const int32_t* pt_order;    // variable points in processing order (sorted by their lowest image)
  const int32_t* chunk_img;   // [n_chunks * nloc] image of each window slot, -1 = unused
  const uint8_t* obs_slot;    // [n_obs] window slot of the observation's image
  double* Z;                  // [n_obs][30]  W_a M_p  (V_p^-1 = M_p M_p')
  double* U;                  // [n_pts][3]   M_p' g_p
};

cudaError_t baf_launch_camera_terms(const BaDev& P, const BaIter& I, const BaTiles& T, int loss_type, double loss_scale, cudaStream_t s);
cudaError_t baf_launch_schur_points(const BaDev& P, const BaWin& W, double radius, double min_diag, double max_diag,
                                    int loss_type, double loss_scale, cudaStream_t s);
cudaError_t baf_launch_schur_window(const BaDev& P, const BaWin& W, const BaTiles& T, int n_sm, cudaStream_t s);
cudaError_t baf_launch_finish(const BaDev& P, const BaTiles& T, double radius, double min_diag, double max_diag, cudaStream_t s);
cudaError_t baf_launch_backsub(const BaDev& P, double* scal, int loss_type, double loss_scale, cudaStream_t s);
// tiled Cholesky of the packed system + the two triangular solves as ONE persistent task-graph kernel (ba_chol.cu)
struct BaCholGraph {                 // host: built once per solve from the tile pattern
  std::vector<int32_t> task;         // [n_tasks][4] kind, i, j, tile
  std::vector<int32_t> dep_ptr, dep; // CSR of (tile, tile | vector block) pairs a task waits for and consumes
  int32_t n_tasks = 0;
};
struct BaCholDev {                   // its device copy + the per-launch state
  const int32_t *task, *dep_ptr, *dep;
  int32_t n_tasks;
  int* flags;                        // bac_flag_count(T) ints
  double* rdiag;                     // [nt * 64]
  unsigned long long* trace;         // nullptr, or [n_tasks * 8] time stamps (B2_BA_CHOL_TRACE, profiling aid)
};
void bac_build_graph(int nt, const int32_t* tile_id, const int32_t* row_ptr, const int32_t* row_col, BaCholGraph* G);
size_t bac_flag_count(const BaTiles& T);
// S x = b in place (x = b on entry, padded to whole tiles, out = solution); S is overwritten by its factor
cudaError_t bac_solve_system(const BaTiles& T, const BaCholDev& G, double* x, int n_sm, cudaStream_t s);

}  // namespace b2
