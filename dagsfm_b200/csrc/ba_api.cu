// C ABI of the bundle adjuster (include/dagsfm_b200.h, BA section): packs the problem
// into HBM, lays out the reduced camera system, and runs the Levenberg-Marquardt loop that
// ceres::Solve runs for the reference (BundleAdjuster::Solve, bundle_adjustment.cc:258-310)
// with the Ceres 1.14 trust-region defaults colmap leaves untouched:
//   initial radius 1e4, max 1e16, min 1e-32; min_relative_decrease 1e-3; LM diagonal
//   clamp [1e-6, 1e32]; Jacobi scaling 1/(1+|col|); radius /= max(1/3, 1-(2 rho-1)^3) on
//   success, /= decrease_factor (2, doubling) on failure.
// Up to 1000 images (the reference's rule, bundle_adjustment.cc:274-284) the reduced system is
// solved exactly (DENSE_SCHUR / SPARSE_SCHUR semantics) with a dense Cholesky -- cuSOLVER
// potrf/potrs, a library call for the factorisation only; Jacobian, Schur elimination and
// back-substitution are the hand-written kernels in ba_kernels.cu.  Above, ITERATIVE_SCHUR +
// SCHUR_JACOBI: Ceres' ConjugateGradientsSolver loop (restated in cg_solve below) over the
// matrix-free Schur product and block-Jacobi preconditioner of ba_iterative.cu.
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/dagsfm_b200.h"
#include "ba_common.cuh"
#include "common_host.h"

using namespace b2;

namespace b2 {  // exclusive scan kernel shared with the matcher (match_post.cu)
cudaError_t launch_scan_u32(const uint32_t* in, int64_t n, uint32_t* out, uint32_t* total, cudaStream_t s);
}

struct b2_ba {
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  cusolverDnHandle_t solver = nullptr;
  cusolverDnParams_t solver_params = nullptr;
  b2_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  void* nccl_comm = nullptr;   // b2_ba_init_nccl: the library's own communicator (takes precedence over the hook)
  std::vector<void*> allocs;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

namespace {

// camera_models.h:117-129 and every model's Initialize{FocalLength,PrincipalPoint,ExtraParams}Idxs
constexpr int kMaxParams = 12;
int num_params(int model) {
  static const int n[11] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
  return (model >= 0 && model <= 10) ? n[model] : 0;
}
bool two_focal(int model) { return model == 1 || (model >= 4 && model <= 7) || model == 10; }
void param_kinds(int model, int kind[kMaxParams]) {  // 0 focal, 1 principal point, 2 extra, -1 unused
  const int nf = two_focal(model) ? 2 : 1, n = num_params(model);
  for (int k = 0; k < kMaxParams; ++k) kind[k] = k >= n ? -1 : k < nf ? 0 : k < nf + 2 ? 1 : 2;
}

template <typename T>
int dev_alloc(b2_ba* h, T** p, size_t n) {
  *p = nullptr;
  B2_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  h->allocs.push_back(*p);
  return B2_OK;
}
template <typename T>
int dev_upload(b2_ba* h, T** p, const T* src, size_t n) {
  B2_TRY(dev_alloc(h, p, n));
  if (n) B2_CUDA(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return B2_OK;
}
void free_all(b2_ba* h) {
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
}

// ---- NCCL, bound at run time: libnccl.so.2 as the process already holds it (e.g. the one torch brought) or the path in
// B2_NCCL_LIBRARY.  Only the five entry points the bundle adjuster needs; enum values of nccl.h 2.x.
namespace nccl {
struct UniqueId { char internal[128]; };
typedef void* Comm;
struct Api {
  void* lib = nullptr;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
  int (*CommDestroy)(Comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kFloat64 = 8, kSum = 0, kMax = 2;
Api* api() {
  static Api a;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[3] = {getenv("B2_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      if ((a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (a.lib) {
      a.GetUniqueId = (int (*)(UniqueId*))dlsym(a.lib, "ncclGetUniqueId");
      a.CommInitRank = (int (*)(Comm*, int, UniqueId, int))dlsym(a.lib, "ncclCommInitRank");
      a.AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))dlsym(a.lib, "ncclAllReduce");
      a.CommDestroy = (int (*)(Comm))dlsym(a.lib, "ncclCommDestroy");
      a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
      if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) a.lib = nullptr;
    }
  }
  return a.lib ? &a : nullptr;
}
}  // namespace nccl

inline bool distributed(const b2_ba* h) { return h->nccl_comm != nullptr || h->allreduce != nullptr; }

// In-place reduction of a device buffer across the ranks.  With the library's own communicator this is one ncclAllReduce
// on the solver's stream -- no host synchronisation, the next kernel is simply ordered behind it; the caller-supplied
// hook (b2_ba_set_allreduce) is a host call and needs the stream drained first.
int sync_reduce(b2_ba* h, double* buf, int64_t n, int op) {
  if (n == 0) return B2_OK;
  if (h->nccl_comm) {
    nccl::Api* a = nccl::api();
    const int rc = a->AllReduce(buf, buf, (size_t)n, nccl::kFloat64, op == 1 ? nccl::kMax : nccl::kSum, h->nccl_comm, h->stream);
    if (rc != 0) return set_error(B2_ERR_CUDA, a->GetErrorString ? a->GetErrorString(rc) : "ncclAllReduce failed");
    return B2_OK;
  }
  if (!h->allreduce) return B2_OK;
  B2_CUDA(cudaStreamSynchronize(h->stream));
  h->allreduce(buf, n, op, h->allreduce_user);
  return B2_OK;
}

// Sum of n per-block partial sums (fixed order => the same value on every run and every rank).
int sum_partials(b2_ba* h, const double* d_partial, int n, double* out) {
  double hp[kBaIterMaxPartials];
  B2_CUDA(cudaMemcpyAsync(hp, d_partial, (size_t)n * 8, cudaMemcpyDeviceToHost, h->stream));
  B2_CUDA(cudaStreamSynchronize(h->stream));
  double s = 0;
  for (int k = 0; k < n; ++k) s += hp[k];
  *out = s;
  return B2_OK;
}

// ------------------------------------------------------------------ fused exact path: host-side plan (once per solve)
// The structure of a bundle-adjustment problem does not change across LM iterations, so everything that depends on
// it alone is laid out here: the processing order of the points, the windows (chunks of points whose images fit NLOC
// slots) of schur_window_kernel, and the tile pattern of the packed reduced system incl. the fill of its Cholesky factor.
struct FusedPlan {
  int nloc = 12;
  std::vector<int32_t> pt_order, chunk_pt0, chunk_img;
  std::vector<uint8_t> obs_slot;
  int nt = 0, n_tiles = 0;
  std::vector<uint8_t> tmap;                       // nt x nt, upper: structural non-zero
  std::vector<int32_t> tile_id, row_ptr, row_col;  // packed layout (after finish_tiles)
};
constexpr int kChunkMaxPoints = 192;

// false: some variable point does not fit a window (track longer than kWinMaxLoc, or one image twice) -> staged path
bool plan_windows(const b2_ba_problem* pr, const std::vector<int64_t>& pt_start, const std::vector<int32_t>& pt_col, FusedPlan* F) {
  const int n_pts = pr->n_points;
  int max_len = 0;
  std::vector<int64_t> key;  // (lowest image, highest image, point) packed for the sort
  key.reserve(n_pts);
  int32_t tmp[kWinMaxLoc];
  for (int p = 0; p < n_pts; ++p) {
    if (pt_col[p] < 0) continue;
    const int64_t o0 = pt_start[p];
    const int L = (int)(pt_start[p + 1] - o0);
    if (L > kWinMaxLoc) return false;
    if (L == 0) continue;
    for (int a = 0; a < L; ++a) tmp[a] = pr->obs_image[o0 + a];
    std::sort(tmp, tmp + L);
    for (int a = 1; a < L; ++a)
      if (tmp[a] == tmp[a - 1]) return false;
    max_len = std::max(max_len, L);
    if (pr->n_images >= (1 << 21) || n_pts >= (1 << 21)) return false;  // key packing below
    key.push_back(((int64_t)tmp[0] << 42) | ((int64_t)tmp[L - 1] << 21) | (int64_t)p);
  }
  std::sort(key.begin(), key.end());
  F->nloc = max_len <= 12 ? 12 : 16;
  const int nloc = F->nloc;
  F->obs_slot.assign((size_t)pr->n_obs, 0);
  F->pt_order.clear();
  F->chunk_pt0.assign(1, 0);
  F->chunk_img.clear();
  std::vector<int32_t> cur;  // images of the open chunk
  int cur_pts = 0;
  auto close_chunk = [&]() {
    if (cur_pts == 0) return;
    for (int k = 0; k < nloc; ++k) F->chunk_img.push_back(k < (int)cur.size() ? cur[k] : -1);
    F->chunk_pt0.push_back((int32_t)F->pt_order.size());
    cur.clear();
    cur_pts = 0;
  };
  for (int64_t kv : key) {
    const int p = (int)(kv & ((1 << 21) - 1));
    const int64_t o0 = pt_start[p];
    const int L = (int)(pt_start[p + 1] - o0);
    int fresh = 0;
    for (int a = 0; a < L; ++a)
      fresh += std::find(cur.begin(), cur.end(), pr->obs_image[o0 + a]) == cur.end() ? 1 : 0;
    if ((int)cur.size() + fresh > nloc || cur_pts >= kChunkMaxPoints) {
      close_chunk();
    }
    for (int a = 0; a < L; ++a) {
      const int32_t img = pr->obs_image[o0 + a];
      auto it = std::find(cur.begin(), cur.end(), img);
      if (it == cur.end()) { cur.push_back(img); it = cur.end() - 1; }
      F->obs_slot[(size_t)(o0 + a)] = (uint8_t)(it - cur.begin());
    }
    F->pt_order.push_back(p);
    ++cur_pts;
  }
  close_chunk();
  return true;
}

// tile pattern of S before fill: every pair of columns that some window or some image's own block can touch
void plan_tile_map(const b2_ba_problem* pr, const std::vector<int32_t>& pose_col, const std::vector<int32_t>& intr_col, int64_t D,
                   FusedPlan* F) {
  const int nt = (int)((D + kST - 1) / kST);
  F->nt = nt;
  F->tmap.assign((size_t)nt * nt, 0);
  for (int t = 0; t < nt; ++t) F->tmap[(size_t)t * nt + t] = 1;
  std::vector<int> tl;
  auto mark_images = [&](const int32_t* imgs, int n) {
    tl.clear();
    for (int k = 0; k < n; ++k) {
      const int i = imgs[k];
      if (i < 0) continue;
      for (int c = 0; c < 6; ++c)
        if (pose_col[6 * (size_t)i + c] >= 0) tl.push_back(pose_col[6 * (size_t)i + c] / kST);
      const int cm = pr->image_camera[i];
      for (int c = 0; c < 4; ++c)
        if (intr_col[4 * (size_t)cm + c] >= 0) tl.push_back(intr_col[4 * (size_t)cm + c] / kST);
    }
    std::sort(tl.begin(), tl.end());
    tl.erase(std::unique(tl.begin(), tl.end()), tl.end());
    for (size_t a = 0; a < tl.size(); ++a)
      for (size_t b = a; b < tl.size(); ++b) F->tmap[(size_t)tl[a] * nt + tl[b]] = 1;
  };
  const int n_chunks = (int)F->chunk_pt0.size() - 1;
  for (int c = 0; c < n_chunks; ++c) mark_images(&F->chunk_img[(size_t)c * F->nloc], F->nloc);
  for (int32_t i = 0; i < pr->n_images; ++i) mark_images(&i, 1);
}

// symbolic right-looking fill + packed layout (CSR by tile row, diagonal tile first)
void finish_tiles(FusedPlan* F) {
  const int nt = F->nt;
  std::vector<int> cols;
  for (int k = 0; k < nt; ++k) {
    cols.clear();
    for (int j = k + 1; j < nt; ++j)
      if (F->tmap[(size_t)k * nt + j]) cols.push_back(j);
    for (size_t a = 0; a < cols.size(); ++a)
      for (size_t b = a; b < cols.size(); ++b) F->tmap[(size_t)cols[a] * nt + cols[b]] = 1;
  }
  F->tile_id.assign((size_t)nt * nt, -1);
  F->row_ptr.assign(nt + 1, 0);
  F->row_col.clear();
  for (int k = 0; k < nt; ++k) {
    for (int j = k; j < nt; ++j)
      if (F->tmap[(size_t)k * nt + j]) {
        F->tile_id[(size_t)k * nt + j] = (int32_t)F->row_col.size();
        F->row_col.push_back(j);
      }
    F->row_ptr[k + 1] = (int32_t)F->row_col.size();
  }
  F->n_tiles = (int)F->row_col.size();
}

// task graph of the tiled Cholesky (ba_chol.cu) for the plan's tile pattern -> device
int upload_chol_graph(b2_ba* h, const FusedPlan& plan, const BaTiles& T, BaCholDev* out) {
  BaCholGraph G;
  bac_build_graph(plan.nt, plan.tile_id.data(), plan.row_ptr.data(), plan.row_col.data(), &G);
  int32_t *d_task, *d_ptr, *d_dep;
  if (G.dep.empty()) G.dep.assign(2, 0);
  B2_TRY(dev_upload(h, &d_task, (const int32_t*)G.task.data(), G.task.size()));
  B2_TRY(dev_upload(h, &d_ptr, (const int32_t*)G.dep_ptr.data(), G.dep_ptr.size()));
  B2_TRY(dev_upload(h, &d_dep, (const int32_t*)G.dep.data(), G.dep.size()));
  out->task = d_task; out->dep_ptr = d_ptr; out->dep = d_dep; out->n_tasks = G.n_tasks;
  B2_TRY(dev_alloc(h, &out->flags, bac_flag_count(T)));
  B2_TRY(dev_alloc(h, &out->rdiag, (size_t)plan.nt * kST));
  out->trace = nullptr;
  if (getenv("B2_BA_CHOL_TRACE")) {  // profiling aid: per-task time stamps, dumped after the first factorisation
    B2_TRY(dev_alloc(h, &out->trace, (size_t)G.n_tasks * 8));
    B2_CUDA(cudaMemsetAsync(out->trace, 0, (size_t)G.n_tasks * 64, h->stream));
  }
  return B2_OK;
}

struct CgVectors { double *x, *r, *p, *q, *tmp, *partial; };

// ceres::internal::ConjugateGradientsSolver::Solve (Ceres 1.14 conjugate_gradients_solver.cc, external)
// on S x = rhs as IterativeSchurComplementSolver drives it: x0 = 0, min_num_iterations 0,
// r_tolerance -1 (off), q_tolerance = eta = 0.1 (LevenbergMarquardtStrategy), residual reset every
// 10 iterations, preconditioner SCHUR_JACOBI.  *usable = false <=> LINEAR_SOLVER_FAILURE (the LM
// step is then invalid); hitting max_iter or p'Sp <= 0 keeps the current x (NO_CONVERGENCE).
// S p = D_c^2 p + [all-reduce over ranks of] F'(F p - E (E'E)^-1 E'F p): the only collective of the loop.
int cg_solve(b2_ba* h, const BaDev& P, const BaIter& I, const CgVectors& V, int max_iter, int64_t* n_iter, bool* usable) {
  const int64_t D = P.D;
  cudaStream_t s = h->stream;
  *usable = true;
  B2_CUDA(cudaMemsetAsync(V.x, 0, (size_t)D * 8, s));
  int np = 0;
  double v = 0;
  B2_CUDA(bai_launch_dot(D, P.rhs, P.rhs, V.partial, &np, s));
  B2_TRY(sum_partials(h, V.partial, np, &v));
  count_launches(1);
  const double norm_b = std::sqrt(v);
  if (norm_b == 0.0) return B2_OK;
  B2_CUDA(cudaMemcpyAsync(V.r, P.rhs, (size_t)D * 8, cudaMemcpyDeviceToDevice, s));  // r = b - S 0
  const double q_tolerance = 0.1;
  const int residual_reset_period = 10;
  double rho = 1.0, Q0 = -1.0 * 0.0;
  for (int it = 1;; ++it) {
    ++*n_iter;
    double* z = V.q;  // Ceres aliases z and q the same way
    B2_CUDA(bai_launch_cg_precond(D, I, V.r, z, V.partial, &np, s));
    const double last_rho = rho;
    B2_TRY(sum_partials(h, V.partial, np, &rho));
    if (rho == 0.0 || std::isinf(rho)) { *usable = false; break; }
    double beta = 0;
    if (it > 1) {
      beta = rho / last_rho;
      if (beta == 0.0 || std::isinf(beta)) { *usable = false; break; }
    }
    B2_CUDA(bai_launch_cg_update_p(D, z, V.p, beta, it == 1, s));
    B2_CUDA(bai_launch_matvec(P, I, V.p, V.q, s));
    B2_TRY(sync_reduce(h, V.q, D, 0));
    B2_CUDA(bai_launch_cg_finish_q(D, I.lm_c, V.p, V.q, V.partial, &np, s));
    double pq = 0;
    B2_TRY(sum_partials(h, V.partial, np, &pq));
    count_launches(5);
    if (pq <= 0 || std::isinf(pq)) break;
    const double alpha = rho / pq;
    if (std::isinf(alpha)) { *usable = false; break; }
    if (it % residual_reset_period == 0) {
      B2_CUDA(bai_launch_cg_update_xr(D, V.x, V.p, V.r, V.q, P.rhs, alpha, 1, V.partial, &np, s));
      B2_CUDA(bai_launch_matvec(P, I, V.x, V.tmp, s));
      B2_TRY(sync_reduce(h, V.tmp, D, 0));
      B2_CUDA(bai_launch_cg_finish_q(D, I.lm_c, V.x, V.tmp, V.partial, &np, s));
      B2_CUDA(bai_launch_cg_update_xr(D, V.x, V.p, V.r, V.tmp, P.rhs, alpha, 2, V.partial, &np, s));
      count_launches(5);
    } else {
      B2_CUDA(bai_launch_cg_update_xr(D, V.x, V.p, V.r, V.q, P.rhs, alpha, 0, V.partial, &np, s));
      count_launches(1);
    }
    double xq = 0, rr = 0;
    B2_TRY(sum_partials(h, V.partial, np, &xq));
    B2_TRY(sum_partials(h, V.partial + kBaIterMaxPartials, np, &rr));
    const double Q1 = -1.0 * xq;
    const double zeta = it * (Q1 - Q0) / Q1;
    if (zeta < q_tolerance) break;
    Q0 = Q1;
    (void)rr;  // residual-based termination is disabled (r_tolerance = -1 => tol_r < 0 <= |r|)
    if (it >= max_iter) break;
  }
  if (*usable) {  // IsArrayValid(step): a non-finite entry makes the step a linear-solver failure
    B2_CUDA(bai_launch_dot(D, V.x, V.x, V.partial, &np, s));
    B2_TRY(sum_partials(h, V.partial, np, &v));
    count_launches(1);
    if (!std::isfinite(v)) *usable = false;
  }
  return B2_OK;
}

int solve_impl(b2_ba* h, const b2_ba_problem* pr, const b2_ba_options* opt, b2_ba_summary* sum) {
  cudaStream_t s = h->stream;
  const int loss_type = opt->loss_function_type;
  const double loss_scale = opt->loss_function_scale;
  const int n_img = pr->n_images, n_cam = pr->n_cameras, n_pts = pr->n_points;
  const int64_t n_obs = pr->n_obs;
  // ---------------------------------------------------------------- validation
  for (int64_t o = 0; o < n_obs; ++o) {
    if (pr->obs_image[o] < 0 || pr->obs_image[o] >= n_img || pr->obs_point[o] < 0 || pr->obs_point[o] >= n_pts)
      return set_error(B2_ERR_INVALID, "observation index out of range");
    if (o > 0 && pr->obs_point[o] < pr->obs_point[o - 1])
      return set_error(B2_ERR_INVALID, "observations must be sorted by point");
  }
  for (int i = 0; i < n_img; ++i)
    if (pr->image_camera[i] < 0 || pr->image_camera[i] >= n_cam) return set_error(B2_ERR_INVALID, "bad image_camera");
  // SIMPLE_PINHOLE / PINHOLE / SIMPLE_RADIAL (<= 4 parameters) run on the 224-byte Jacobian layout; any other model of
  // camera_models.h switches the whole problem to the wide layout (12 intrinsics slots, dual-number derivatives)
  const int cs = pr->camera_params_stride > 0 ? pr->camera_params_stride : 4;
  bool wide = false;
  for (int c = 0; c < n_cam; ++c) {
    if (pr->camera_model[c] < 0 || pr->camera_model[c] > 10) return set_error(B2_ERR_INVALID, "unknown camera model id");
    if (num_params(pr->camera_model[c]) > cs) return set_error(B2_ERR_INVALID, "camera_params_stride smaller than the model's parameter count");
    wide = wide || pr->camera_model[c] > 2;
  }
  const int KI = wide ? kMaxParams : 4;
  // ---------------------------------------------------------------- layout
  // Column order: images (rotation 3, variable tvec components), then cameras (variable
  // intrinsics) -- only blocks that appear in a residual (as Ceres' reduced program).
  // In a multi-GPU run every rank sees all images / cameras of the job, so `used` must be
  // the union: a rank marks everything that is not explicitly constant as used when a
  // collective hook is installed.
  std::vector<char> img_used(n_img, distributed(h) ? 1 : 0), cam_used(n_cam, distributed(h) ? 1 : 0), pt_used(n_pts, 0);
  for (int64_t o = 0; o < n_obs; ++o) {
    img_used[pr->obs_image[o]] = 1;
    cam_used[pr->image_camera[pr->obs_image[o]]] = 1;
    pt_used[pr->obs_point[o]] = 1;
  }
  std::vector<int32_t> pose_col((size_t)n_img * 6, -1), intr_col((size_t)n_cam * KI, -1), pt_col(n_pts, -1);
  int64_t D = 0;
  for (int i = 0; i < n_img; ++i) {
    if (!img_used[i] || pr->const_pose[i]) continue;
    for (int k = 0; k < 3; ++k) pose_col[6 * i + k] = (int32_t)D++;
    for (int k = 0; k < 3; ++k)
      if (!(pr->const_tvec[i] & (1 << k))) pose_col[6 * i + 3 + k] = (int32_t)D++;
  }
  for (int c = 0; c < n_cam; ++c) {
    if (!cam_used[c] || pr->const_camera[c]) continue;
    int kind[kMaxParams];
    param_kinds(pr->camera_model[c], kind);
    for (int k = 0; k < num_params(pr->camera_model[c]); ++k) {
      const bool var = (kind[k] == 0 && opt->refine_focal_length) || (kind[k] == 1 && opt->refine_principal_point) ||
                       (kind[k] == 2 && opt->refine_extra_params);
      if (var) intr_col[(size_t)KI * c + k] = (int32_t)D++;
    }
  }
  int64_t NP = 0;
  for (int p = 0; p < n_pts; ++p)
    if (pt_used[p] && !pr->const_point[p]) pt_col[p] = (int32_t)NP++;
  // BundleAdjuster::Solve's "empirical choice" (bundle_adjustment.cc:274-284): direct solvers up to
  // kMaxNumImagesDirectSparseSolver = 1000 images, ITERATIVE_SCHUR + SCHUR_JACOBI above
  const bool iterative = opt->linear_solver_type == 2 || (opt->linear_solver_type == 0 && n_img > 1000);
  sum->linear_solver_type_used = iterative ? 2 : 1;
  sum->num_linear_solver_iterations = 0;
  std::vector<int64_t> pt_start(n_pts + 1, 0);
  for (int64_t o = 0; o < n_obs; ++o) pt_start[pr->obs_point[o] + 1]++;
  for (int p = 0; p < n_pts; ++p) pt_start[p + 1] += pt_start[p];
  // Exact step: the fused path (ba_fused.cu + ba_chol.cu: no staged Jacobian blocks, packed tiles, own Cholesky) whenever
  // the problem qualifies -- 4-slot camera models, every variable track inside a window; B2_BA_EXACT=staged forces the
  // first-generation path (staged ObsJac, dense S, cuSOLVER), which also serves everything that does not qualify.
  FusedPlan plan;
  bool fused = !iterative && !wide && n_obs > 0 && n_obs < 0x7fffffffLL && D > 0;
  if (const char* e = getenv("B2_BA_EXACT")) fused = fused && strcmp(e, "staged") != 0;
  if (fused) fused = plan_windows(pr, pt_start, pt_col, &plan);
  if (distributed(h)) {  // every rank must take the same path: MIN over the ranks of the local decision
    double* d_flag = nullptr;
    B2_TRY(dev_alloc(h, &d_flag, 1));
    const double mine = fused ? 0.0 : 1.0;
    B2_CUDA(cudaMemcpyAsync(d_flag, &mine, 8, cudaMemcpyHostToDevice, s));
    B2_TRY(sync_reduce(h, d_flag, 1, 1));
    double any_staged = 0;
    B2_CUDA(cudaMemcpyAsync(&any_staged, d_flag, 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    fused = fused && any_staged == 0.0;
  }
  if (fused) {
    plan_tile_map(pr, pose_col, intr_col, D, &plan);
    if (distributed(h)) {  // union of the ranks' tile patterns (each rank sees its own points only)
      std::vector<double> m(plan.tmap.begin(), plan.tmap.end());
      double* d_m = nullptr;
      B2_TRY(dev_upload(h, &d_m, (const double*)m.data(), m.size()));
      B2_TRY(sync_reduce(h, d_m, (int64_t)m.size(), 1));
      B2_CUDA(cudaMemcpyAsync(m.data(), d_m, m.size() * 8, cudaMemcpyDeviceToHost, s));
      B2_CUDA(cudaStreamSynchronize(s));
      for (size_t k = 0; k < m.size(); ++k) plan.tmap[k] = m[k] != 0.0 ? 1 : 0;
    }
    finish_tiles(&plan);
  }
  sum->exact_path_used = iterative ? 0 : (fused ? 2 : 1);
  sum->linear_solve_seconds = 0;
  sum->reduced_system_bytes = 0;
  {  // exact path: the reduced camera system must fit in this GPU's memory
    size_t free_b = 0, total_b = 0;
    B2_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double need = fused ? ((double)plan.n_tiles * kST * kST + (double)plan.nt * kST * kST + 8.0 * (double)D) * 8.0 + (double)n_obs * 260.0
                              : (iterative ? 24.0 * (double)D : (double)D * (double)D + 3.0 * (double)D) * 8.0 +
                        (double)n_obs * ((wide ? sizeof(ObsJacW) : sizeof(ObsJac)) + (iterative ? 4.0 : 0.0));
    if (need > 0.9 * (double)free_b)
      return set_error(B2_ERR_INVALID, iterative ? "problem too large for this GPU's memory"
                                                 : "reduced camera system too large for the dense Schur path on this GPU "
                                                   "(linear_solver_type = 2 selects ITERATIVE_SCHUR)");
  }
  if (iterative && n_obs >= 0x7fffffffLL) return set_error(B2_ERR_INVALID, "too many observations for ITERATIVE_SCHUR");

  // Ceres' reduced program drops residual blocks whose parameter blocks are all constant
  // (bundle_adjustment_test.cc:360-364 counts 402, not 404, for exactly that reason)
  int64_t n_obs_reduced = 0;
  for (int64_t o = 0; o < n_obs; ++o) {
    const int i = pr->obs_image[o], c = pr->image_camera[i];
    bool free_block = pt_col[pr->obs_point[o]] >= 0;
    for (int k = 0; k < 6 && !free_block; ++k) free_block = pose_col[6 * i + k] >= 0;
    for (int k = 0; k < KI && !free_block; ++k) free_block = intr_col[(size_t)KI * c + k] >= 0;
    n_obs_reduced += free_block ? 1 : 0;
  }
  sum->num_residuals_reduced = (int32_t)(2 * n_obs_reduced);
  sum->num_effective_parameters_reduced = (int32_t)(D + 3 * NP);
  sum->num_successful_steps = sum->num_unsuccessful_steps = sum->num_iterations = 0;
  sum->termination_type = 1;
  sum->initial_cost = sum->final_cost = 0;
  sum->solve_seconds = sum->schur_kernel_seconds = 0;
  sum->schur_kernel_launches = 0;

  // image.NormalizeQvec() (bundle_adjustment.cc:345)
  std::vector<double> qn(pr->qvec, pr->qvec + (size_t)n_img * 4);
  for (int i = 0; i < n_img; ++i) {
    double* q = &qn[4 * i];
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n > 0) for (int k = 0; k < 4; ++k) q[k] /= n;
  }

  // ---------------------------------------------------------------- upload
  BaDev P;
  memset(&P, 0, sizeof P);
  P.n_img = n_img; P.n_cam = n_cam; P.n_pts = n_pts; P.n_obs = n_obs; P.D = D;
  P.wide = wide ? 1 : 0;
  // camera parameters on the device: KI doubles per camera, zero-padded (the host array has its own stride)
  std::vector<double> cam_dev((size_t)n_cam * KI, 0.0);
  for (int c = 0; c < n_cam; ++c)
    for (int k = 0; k < num_params(pr->camera_model[c]); ++k) cam_dev[(size_t)KI * c + k] = pr->camera_params[(size_t)cs * c + k];
  int32_t *d_obs_img, *d_obs_pt, *d_img_cam, *d_cam_model, *d_pose_col, *d_intr_col, *d_pt_col;
  double* d_obs_xy;
  int64_t* d_pt_start;
  B2_TRY(dev_upload(h, &d_obs_img, pr->obs_image, (size_t)n_obs));
  B2_TRY(dev_upload(h, &d_obs_pt, pr->obs_point, (size_t)n_obs));
  B2_TRY(dev_upload(h, &d_obs_xy, pr->obs_xy, (size_t)n_obs * 2));
  B2_TRY(dev_upload(h, &d_pt_start, pt_start.data(), pt_start.size()));
  B2_TRY(dev_upload(h, &d_img_cam, pr->image_camera, (size_t)n_img));
  B2_TRY(dev_upload(h, &d_cam_model, pr->camera_model, (size_t)n_cam));
  B2_TRY(dev_upload(h, &d_pose_col, pose_col.data(), pose_col.size()));
  {  // rotation blocks by column: their term of the gradient max norm goes through QuaternionParameterization::Plus
    std::vector<int32_t> rot_img((size_t)std::max<int64_t>(D, 1), -1);
    for (int i = 0; i < n_img; ++i) {
      const int32_t c = pose_col[6 * (size_t)i];
      if (c < 0) continue;
      rot_img[c] = i;
      rot_img[c + 1] = rot_img[c + 2] = -2;
    }
    int32_t* d_rot_img;
    B2_TRY(dev_upload(h, &d_rot_img, rot_img.data(), rot_img.size()));
    P.rot_img = d_rot_img;
  }
  B2_TRY(dev_upload(h, &d_intr_col, intr_col.data(), intr_col.size()));
  B2_TRY(dev_upload(h, &d_pt_col, pt_col.data(), pt_col.size()));
  P.obs_img = d_obs_img; P.obs_pt = d_obs_pt; P.obs_xy = (const double2*)d_obs_xy; P.pt_start = d_pt_start;
  P.img_cam = d_img_cam; P.cam_model = d_cam_model; P.pose_col = d_pose_col; P.intr_col = d_intr_col; P.pt_col = d_pt_col;
  B2_TRY(dev_upload(h, &P.qvec, qn.data(), qn.size()));
  B2_TRY(dev_upload(h, &P.tvec, (const double*)pr->tvec, (size_t)n_img * 3));
  B2_TRY(dev_upload(h, &P.cam_params, (const double*)cam_dev.data(), cam_dev.size()));
  B2_TRY(dev_upload(h, &P.xyz, (const double*)pr->xyz, (size_t)n_pts * 3));
  B2_TRY(dev_alloc(h, &P.qvec_new, (size_t)n_img * 4));
  B2_TRY(dev_alloc(h, &P.tvec_new, (size_t)n_img * 3));
  B2_TRY(dev_alloc(h, &P.cam_new, (size_t)n_cam * KI));
  B2_TRY(dev_alloc(h, &P.xyz_new, (size_t)n_pts * 3));
  if (wide) B2_TRY(dev_alloc(h, &P.JW, (size_t)n_obs));
  else if (!fused) B2_TRY(dev_alloc(h, &P.J, (size_t)n_obs));  // the fused path never stages Jacobian blocks
  B2_TRY(dev_alloc(h, &P.scale_c, (size_t)D));
  B2_TRY(dev_alloc(h, &P.scale_p, (size_t)NP * 3));
  B2_TRY(dev_alloc(h, &P.colnorm_c, (size_t)D));
  B2_TRY(dev_alloc(h, &P.colnorm_p, (size_t)NP * 3));
  double* reduced;  // S | rhs | g_c | diag_c : one buffer, one all-reduce
  // S is never formed by ITERATIVE_SCHUR; the fused path keeps it in packed tiles and pads the vectors to whole tiles
  const size_t Dv = fused ? (size_t)plan.nt * kST : (size_t)D;
  const size_t n_reduced = (iterative ? 0 : fused ? (size_t)plan.n_tiles * kST * kST : (size_t)D * D) + 3 * Dv;
  B2_TRY(dev_alloc(h, &reduced, n_reduced));
  P.S = (iterative || fused) ? nullptr : reduced; P.rhs = reduced + (n_reduced - 3 * Dv); P.g_c = P.rhs + Dv; P.diag_c = P.g_c + Dv;
  sum->reduced_system_bytes = (double)n_reduced * 8.0;
  B2_TRY(dev_alloc(h, &P.diag_p, (size_t)NP * 3));
  B2_TRY(dev_alloc(h, &P.g_p, (size_t)NP * 3));
  B2_TRY(dev_alloc(h, &P.Vinv, (size_t)NP * 9));
  B2_TRY(dev_alloc(h, &P.dc, Dv));
  B2_TRY(dev_alloc(h, &P.dp, (size_t)NP * 3));
  BaTiles Tl;
  memset(&Tl, 0, sizeof Tl);
  BaCholDev Cg;
  memset(&Cg, 0, sizeof Cg);
  BaWin Wn;
  memset(&Wn, 0, sizeof Wn);
  if (fused) {
    int32_t *d_tile_id, *d_row_ptr, *d_row_col, *d_order, *d_cpt0, *d_cimg;
    uint8_t* d_slot;
    B2_TRY(dev_upload(h, &d_tile_id, (const int32_t*)plan.tile_id.data(), plan.tile_id.size()));
    B2_TRY(dev_upload(h, &d_row_ptr, (const int32_t*)plan.row_ptr.data(), plan.row_ptr.size()));
    B2_TRY(dev_upload(h, &d_row_col, (const int32_t*)plan.row_col.data(), plan.row_col.size()));
    B2_TRY(dev_upload(h, &d_order, (const int32_t*)plan.pt_order.data(), plan.pt_order.size()));
    B2_TRY(dev_upload(h, &d_cpt0, (const int32_t*)plan.chunk_pt0.data(), plan.chunk_pt0.size()));
    B2_TRY(dev_upload(h, &d_cimg, (const int32_t*)plan.chunk_img.data(), plan.chunk_img.size()));
    B2_TRY(dev_upload(h, &d_slot, (const uint8_t*)plan.obs_slot.data(), plan.obs_slot.size()));
    Tl.nt = plan.nt; Tl.n_tiles = plan.n_tiles; Tl.tile_id = d_tile_id; Tl.row_ptr = d_row_ptr; Tl.row_col = d_row_col;
    Tl.tiles = reduced;
    B2_TRY(dev_alloc(h, &Tl.rinv, (size_t)plan.nt * kST * kST));
    B2_TRY(dev_alloc(h, &Tl.info, 1));
    B2_TRY(upload_chol_graph(h, plan, Tl, &Cg));
    Wn.n_chunks = (int32_t)plan.chunk_pt0.size() - 1; Wn.nloc = plan.nloc; Wn.chunk_pt0 = d_cpt0; Wn.pt_order = d_order;
    Wn.chunk_img = d_cimg; Wn.obs_slot = d_slot;
    B2_TRY(dev_alloc(h, &Wn.Z, (size_t)n_obs * 30));
    B2_TRY(dev_alloc(h, &Wn.U, (size_t)n_pts * 3));
  }
  double* scal;  // [0] step^2 cams [1] x^2 cams [2] step^2 pts [3] x^2 pts [4] model change [5] new cost [6] cost [7] gmax
  B2_TRY(dev_alloc(h, &scal, 8));
  P.gmax = scal + 7;
  int* d_info;
  B2_TRY(dev_alloc(h, &d_info, 1));
  // 64-bit cuSOLVER entry points: D * D exceeds 2^31 elements from D = 46 341 on
  uint8_t* work = nullptr;
  size_t work_dev = 0, work_host = 0;
  std::vector<uint8_t> work_h;
  if (D > 0 && !iterative && !fused) {
    if (cusolverDnXpotrf_bufferSize(h->solver, h->solver_params, CUBLAS_FILL_MODE_LOWER, D, CUDA_R_64F, P.S, D,
                                    CUDA_R_64F, &work_dev, &work_host) != CUSOLVER_STATUS_SUCCESS)
      return set_error(B2_ERR_CUDA, "cusolverDnXpotrf_bufferSize failed");
    B2_TRY(dev_alloc(h, &work, std::max<size_t>(work_dev, 8)));
    work_h.resize(std::max<size_t>(work_host, 8));
  }
  // ---------------------------------------------------------------- ITERATIVE_SCHUR state
  BaIter I;
  memset(&I, 0, sizeof I);
  CgVectors V;
  memset(&V, 0, sizeof V);
  // experimental (B2_BA_CAMTERMS=image): the exact path's camera terms summed per image before they touch S
  bool camterms_image = false;
  if (const char* e = getenv("B2_BA_CAMTERMS")) camterms_image = !iterative && strcmp(e, "image") == 0 && n_obs > 0 && n_obs < 0x7fffffffLL;
  if (iterative || camterms_image || fused) {
    // observations grouped by image (counting sort of the point-major order) for the image-major pass
    std::vector<int64_t> img_start(n_img + 1, 0);
    for (int64_t o = 0; o < n_obs; ++o) img_start[pr->obs_image[o] + 1]++;
    for (int i = 0; i < n_img; ++i) img_start[i + 1] += img_start[i];
    std::vector<int32_t> img_obs((size_t)n_obs);
    {
      std::vector<int64_t> cur(img_start.begin(), img_start.end() - 1);
      for (int64_t o = 0; o < n_obs; ++o) img_obs[(size_t)cur[pr->obs_image[o]]++] = (int32_t)o;
    }
    // Ceres parameter blocks of the camera side: qvec (3 local columns), tvec (its variable components),
    // camera parameters (the variable ones) -- bundle_adjustment.cc:383-418 adds them as separate blocks
    std::vector<int32_t> blk_first((size_t)std::max<int64_t>(D, 1), 0), blk_size((size_t)std::max<int64_t>(D, 1), 0);
    auto mark = [&](const int32_t* cols, int n) {
      int f = -1, cnt = 0;
      for (int k = 0; k < n; ++k) if (cols[k] >= 0) { if (f < 0) f = cols[k]; ++cnt; }
      for (int k = 0; k < n; ++k) if (cols[k] >= 0) { blk_first[cols[k]] = f; blk_size[cols[k]] = cnt; }
    };
    for (int i = 0; i < n_img; ++i) { mark(&pose_col[6 * (size_t)i], 3); mark(&pose_col[6 * (size_t)i + 3], 3); }
    for (int c = 0; c < n_cam; ++c) mark(&intr_col[(size_t)KI * c], KI);
    int64_t* d_img_start; int32_t *d_img_obs, *d_blk_first, *d_blk_size;
    B2_TRY(dev_upload(h, &d_img_start, img_start.data(), img_start.size()));
    B2_TRY(dev_upload(h, &d_img_obs, img_obs.data(), img_obs.size()));
    B2_TRY(dev_upload(h, &d_blk_first, blk_first.data(), (size_t)D));
    B2_TRY(dev_upload(h, &d_blk_size, blk_size.data(), (size_t)D));
    I.img_start = d_img_start; I.img_obs = d_img_obs; I.blk_first = d_blk_first; I.blk_size = d_blk_size;
    B2_CUDA(cudaStreamSynchronize(s));  // the host vectors above go out of scope
  }
  if (iterative) {
    B2_TRY(dev_alloc(h, &I.tp, (size_t)NP * 3));
    B2_TRY(dev_alloc(h, &I.zp, (size_t)NP * 3));
    B2_TRY(dev_alloc(h, &I.lm_c, (size_t)D));
    B2_TRY(dev_alloc(h, &I.M, (size_t)D * KI));
    I.m_stride = KI;
    B2_TRY(dev_alloc(h, &I.flag, 1));
    V.x = P.dc;
    B2_TRY(dev_alloc(h, &V.r, (size_t)D));
    B2_TRY(dev_alloc(h, &V.p, (size_t)D));
    B2_TRY(dev_alloc(h, &V.q, (size_t)D));
    B2_TRY(dev_alloc(h, &V.tmp, (size_t)D));
    B2_TRY(dev_alloc(h, &V.partial, (size_t)2 * kBaIterMaxPartials));
  }
  if (n_obs == 0 && !distributed(h)) return B2_OK;  // BundleAdjuster::Solve returns false: nothing to do

  // ---------------------------------------------------------------- Jacobi scaling + initial cost
  B2_CUDA(cudaMemsetAsync(P.colnorm_c, 0, std::max<size_t>(D, 1) * 8, s));
  B2_CUDA(cudaMemsetAsync(P.colnorm_p, 0, std::max<size_t>(NP * 3, 1) * 8, s));
  B2_CUDA(cudaMemsetAsync(scal, 0, 8 * 8, s));
  B2_CUDA(ba_launch_jacobian(P, P.qvec, P.tvec, P.cam_params, P.xyz, 2, scal + 6, s, loss_type, loss_scale));
  B2_TRY(sync_reduce(h, P.colnorm_c, D, 0));
  B2_TRY(sync_reduce(h, scal + 6, 1, 0));
  B2_CUDA(ba_launch_make_scale(P.colnorm_c, P.scale_c, D, s));
  B2_CUDA(ba_launch_make_scale(P.colnorm_p, P.scale_p, NP * 3, s));
  double cost = 0;
  B2_CUDA(cudaMemcpyAsync(&cost, scal + 6, 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  count_launches(3);
  sum->initial_cost = cost;

  B2_CUDA(cudaEventRecord(h->ev[0], s));
  if (!fused) {
    B2_CUDA(cudaMemsetAsync(scal + 6, 0, 8, s));
    B2_CUDA(ba_launch_jacobian(P, P.qvec, P.tvec, P.cam_params, P.xyz, 0, scal + 6, s, loss_type, loss_scale));
    count_launches(1);
  }

  // Experimental pair-major Schur accumulation (B2_BA_SCHUR=blocks): the (lo, hi) observation tuples of
  // every (image, image) block are gathered once -- the structure is fixed across LM iterations --
  // by a counting sort on the device; per iteration the point-major kernel then only stores W / Y
  // per observation and pm_blocks_kernel sums each block before touching S.  Default: the
  // production kernel with per-tuple atomics.
  bool pair_major = false;
  uint32_t *pm_count = nullptr, *pm_start = nullptr;
  uint64_t* pm_tuples = nullptr;
  double *pm_W = nullptr, *pm_Y = nullptr;
  if (const char* e = getenv("B2_BA_SCHUR")) pair_major = !iterative && !fused && !wide && strcmp(e, "blocks") == 0 && n_img > 0 && n_img <= 4096 && n_obs > 0;
  if (pair_major) {
    uint64_t n_tuples = 0;
    for (int p = 0; p < n_pts; ++p) {
      const uint64_t L = (uint64_t)(pt_start[p + 1] - pt_start[p]);
      n_tuples += (pt_col[p] >= 0) ? L * (L + 1) / 2 : L;  // constant points: diagonal tuples only (camera terms)
    }
    if (n_tuples == 0 || n_tuples >= 0xFFFFFFF0ull) pair_major = false;
    if (pair_major) {
      const int64_t n_keys = (int64_t)n_img * n_img;
      B2_TRY(dev_alloc(h, &pm_count, (size_t)n_keys));
      B2_TRY(dev_alloc(h, &pm_start, (size_t)n_keys + 1));
      B2_TRY(dev_alloc(h, &pm_tuples, (size_t)n_tuples));
      B2_TRY(dev_alloc(h, &pm_W, (size_t)n_obs * 30));
      B2_TRY(dev_alloc(h, &pm_Y, (size_t)n_obs * 30));
      B2_CUDA(cudaMemsetAsync(pm_W, 0, (size_t)n_obs * 30 * 8, s));  // observations of constant points keep W = Y = 0
      B2_CUDA(cudaMemsetAsync(pm_Y, 0, (size_t)n_obs * 30 * 8, s));
      B2_CUDA(cudaMemsetAsync(pm_count, 0, (size_t)n_keys * sizeof(uint32_t), s));
      B2_CUDA(ba_launch_pm_enumerate(P, n_img, false, pm_count, nullptr, nullptr, h->n_sm, s));
      B2_CUDA(launch_scan_u32(pm_count, n_keys, pm_start, pm_start + n_keys, s));
      B2_CUDA(cudaMemsetAsync(pm_count, 0, (size_t)n_keys * sizeof(uint32_t), s));
      B2_CUDA(ba_launch_pm_enumerate(P, n_img, true, pm_count, pm_start, pm_tuples, h->n_sm, s));
      count_launches(3);
    }
  }

  double radius = 1e4, decrease_factor = 2.0;
  const double min_radius = 1e-32, max_radius = 1e16, min_rel_dec = 1e-3, min_diag = 1e-6, max_diag = 1e32;
  const char* dump_path = getenv("B2_BA_DUMP");  // debugging aid: the reduced system of the first iteration, raw doubles
  const bool trace = getenv("B2_BA_TRACE") != nullptr;  // per-iteration line on stderr (debugging aid)
  double schur_ms = 0, solve_ms = 0;
  int64_t schur_launches = 0;
  for (int iter = 0; iter < opt->max_num_iterations; ++iter) {
    // ---- normal equations of the camera block + Schur complement
    B2_CUDA(cudaMemsetAsync(reduced, 0, n_reduced * 8, s));
    B2_CUDA(cudaMemsetAsync(scal, 0, 6 * 8, s));
    B2_CUDA(cudaMemsetAsync(scal + 7, 0, 8, s));
    int info = 0;
    if (iterative) {
      // ---- ITERATIVE_SCHUR: point blocks, rhs / g_c / diag_c, SCHUR_JACOBI blocks, then CG on S x = rhs
      B2_CUDA(cudaEventRecord(h->ev[2], s));
      B2_CUDA(bai_launch_point_prepare(P, I, radius, min_diag, max_diag, s));
      B2_CUDA(bai_launch_rhs(P, I, s));
      B2_TRY(sync_reduce(h, reduced, (int64_t)n_reduced, 0));  // (rhs, g_c, diag_c)
      B2_CUDA(bai_launch_cam_diag(P, I, radius, min_diag, max_diag, s));
      B2_TRY(sync_reduce(h, P.gmax, 1, 1));
      {  // gradient_max_norm <= gradient_tolerance: converged before another inner solve is spent
        double gmax_now = 0;
        B2_CUDA(cudaMemcpyAsync(&gmax_now, P.gmax, 8, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        count_launches(3);
        if (gmax_now <= opt->gradient_tolerance) {
          sum->termination_type = 0;
          break;
        }
      }
      B2_CUDA(cudaMemsetAsync(I.M, 0, std::max<size_t>((size_t)D * KI, 1) * 8, s));
      B2_CUDA(cudaMemsetAsync(I.flag, 0, sizeof(int), s));
      B2_CUDA(bai_launch_precond(P, I, s));
      B2_TRY(sync_reduce(h, I.M, D * KI, 0));
      B2_CUDA(bai_launch_precond_invert(P, I, s));
      count_launches(2);
      if (D > 0) {
        B2_CUDA(cudaMemcpyAsync(&info, I.flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        bool usable = info == 0;  // a preconditioner block that is not PD = LINEAR_SOLVER_FAILURE
        if (usable) B2_TRY(cg_solve(h, P, I, V, opt->max_linear_solver_iterations, &sum->num_linear_solver_iterations, &usable));
        if (!usable) {
          info = 1;
          B2_CUDA(cudaMemsetAsync(P.dc, 0, (size_t)D * 8, s));
        }
        B2_CUDA(ba_launch_negate(P.dc, D, s));  // dc = -x
      }
      B2_CUDA(cudaEventRecord(h->ev[3], s));
      schur_launches += 2;
    } else if (fused) {
      // ---- fused exact step: camera terms + per-point Z + window products -> packed tiles; own Cholesky
      B2_CUDA(cudaEventRecord(h->ev[2], s));
      B2_CUDA(baf_launch_camera_terms(P, I, Tl, loss_type, loss_scale, s));
      B2_CUDA(baf_launch_schur_points(P, Wn, radius, min_diag, max_diag, loss_type, loss_scale, s));
      B2_CUDA(baf_launch_schur_window(P, Wn, Tl, h->n_sm, s));
      B2_CUDA(cudaEventRecord(h->ev[3], s));
      schur_launches += 3;
      count_launches(3);
      B2_TRY(sync_reduce(h, reduced, (int64_t)n_reduced, 0));  // packed tiles + rhs + g_c + diag_c: one all-reduce
      B2_CUDA(baf_launch_finish(P, Tl, radius, min_diag, max_diag, s));
      B2_TRY(sync_reduce(h, P.gmax, 1, 1));
      if (dump_path && iter == 0) {  // debugging aid: the assembled system as a dense upper triangle + rhs
        std::vector<double> tiles((size_t)plan.n_tiles * kST * kST), dense((size_t)D * D, 0.0), rhs_h((size_t)D);
        B2_CUDA(cudaMemcpyAsync(tiles.data(), Tl.tiles, tiles.size() * 8, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaMemcpyAsync(rhs_h.data(), P.rhs, (size_t)D * 8, cudaMemcpyDeviceToHost, s));
        B2_CUDA(cudaStreamSynchronize(s));
        for (int64_t r = 0; r < D; ++r)
          for (int64_t c = r; c < D; ++c) {
            const int id = plan.tile_id[(size_t)(r / kST) * plan.nt + (c / kST)];
            if (id >= 0) dense[(size_t)r * D + c] = tiles[(size_t)id * kST * kST + (r % kST) * kST + (c % kST)];
          }
        if (FILE* f = fopen(dump_path, "wb")) { fwrite(dense.data(), 8, dense.size(), f); fwrite(rhs_h.data(), 8, rhs_h.size(), f); fclose(f); }
      }
      B2_CUDA(cudaEventRecord(h->ev[4], s));
      B2_CUDA(cudaMemsetAsync(Tl.info, 0, sizeof(int), s));
      B2_CUDA(cudaMemcpyAsync(P.dc, P.rhs, Dv * 8, cudaMemcpyDeviceToDevice, s));
      B2_CUDA(bac_solve_system(Tl, Cg, P.dc, h->n_sm, s));  // one persistent task-graph kernel: factor + both solves
      if (Cg.trace && iter == 1) {
        if (const char* path = getenv("B2_BA_CHOL_TRACE")) {
          std::vector<unsigned long long> tr((size_t)Cg.n_tasks * 8);
          std::vector<int32_t> tk((size_t)Cg.n_tasks * 4);
          B2_CUDA(cudaMemcpyAsync(tr.data(), Cg.trace, tr.size() * 8, cudaMemcpyDeviceToHost, s));
          B2_CUDA(cudaMemcpyAsync(tk.data(), Cg.task, tk.size() * 4, cudaMemcpyDeviceToHost, s));
          B2_CUDA(cudaStreamSynchronize(s));
          if (FILE* f = fopen(path, "w")) {
            for (int t = 0; t < Cg.n_tasks; ++t)
              fprintf(f, "%d %d %d %d %llu %llu %llu %llu %llu %llu %llu %llu\n", tk[4 * t], tk[4 * t + 1], tk[4 * t + 2], tk[4 * t + 3], tr[8 * t],
                      tr[8 * t + 1], tr[8 * t + 2], tr[8 * t + 3], tr[8 * t + 4], tr[8 * t + 5], tr[8 * t + 6], tr[8 * t + 7]);
            fclose(f);
          }
        }
      }
      const int nl = 1;
      B2_CUDA(ba_launch_negate(P.dc, D, s));
      B2_CUDA(cudaEventRecord(h->ev[5], s));
      B2_CUDA(cudaMemcpyAsync(&info, Tl.info, sizeof(int), cudaMemcpyDeviceToHost, s));
      count_launches(nl + 3);
    } else {
    B2_CUDA(cudaEventRecord(h->ev[2], s));
    if (!pair_major && camterms_image) B2_CUDA(bai_launch_camera_terms_image(P, I, s));
    else if (!pair_major) B2_CUDA(ba_launch_camera_terms(P, s));  // folded into the (i, i) blocks in pair-major mode
    if (pair_major)
      B2_CUDA(ba_launch_schur_pm(P, radius, min_diag, max_diag, n_img, pm_start, pm_tuples, pm_W, pm_Y, h->n_sm, s));
    else
      B2_CUDA(ba_launch_schur(P, radius, min_diag, max_diag, h->n_sm, s));
    B2_CUDA(cudaEventRecord(h->ev[3], s));
    schur_launches += 2;
    B2_TRY(sync_reduce(h, reduced, (int64_t)n_reduced, 0));  // the one NVLink all-reduce of (S, rhs, g_c, diag)
    B2_CUDA(ba_launch_add_diag(P, radius, min_diag, max_diag, s));
    B2_TRY(sync_reduce(h, P.gmax, 1, 1));
    if (dump_path && iter == 0) {
      std::vector<double> dense((size_t)D * D), rhs_h((size_t)D);
      B2_CUDA(cudaMemcpyAsync(dense.data(), P.S, dense.size() * 8, cudaMemcpyDeviceToHost, s));
      B2_CUDA(cudaMemcpyAsync(rhs_h.data(), P.rhs, (size_t)D * 8, cudaMemcpyDeviceToHost, s));
      B2_CUDA(cudaStreamSynchronize(s));
      if (FILE* f = fopen(dump_path, "wb")) { fwrite(dense.data(), 8, dense.size(), f); fwrite(rhs_h.data(), 8, rhs_h.size(), f); fclose(f); }
    }
    // ---- reduced solve: S y = rhs, dc = -y
    if (D > 0) {
      B2_CUDA(cudaMemcpyAsync(P.dc, P.rhs, D * 8, cudaMemcpyDeviceToDevice, s));
      if (cusolverDnXpotrf(h->solver, h->solver_params, CUBLAS_FILL_MODE_LOWER, D, CUDA_R_64F, P.S, D, CUDA_R_64F,
                           work, work_dev, work_h.data(), work_host, d_info) != CUSOLVER_STATUS_SUCCESS)
        return set_error(B2_ERR_CUDA, "cusolverDnXpotrf failed");
      B2_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, s));
      if (cusolverDnXpotrs(h->solver, h->solver_params, CUBLAS_FILL_MODE_LOWER, D, 1, CUDA_R_64F, P.S, D, CUDA_R_64F,
                           P.dc, D, d_info) != CUSOLVER_STATUS_SUCCESS)
        return set_error(B2_ERR_CUDA, "cusolverDnXpotrs failed");
      B2_CUDA(ba_launch_negate(P.dc, D, s));
    }
    }
    if (fused) {
      B2_CUDA(baf_launch_backsub(P, scal, loss_type, loss_scale, s));  // dp, model cost change, candidate points
    } else {
      B2_CUDA(ba_launch_backsub(P, s));
      B2_CUDA(ba_launch_model_cost(P, scal + 4, s));
      B2_CUDA(ba_launch_candidate(P, scal, false, s));
    }
    B2_CUDA(ba_launch_candidate(P, scal, true, s));
    B2_CUDA(ba_launch_jacobian(P, P.qvec_new, P.tvec_new, P.cam_new, P.xyz_new, 1, scal + 5, s, loss_type, loss_scale));
    count_launches(10);
    B2_TRY(sync_reduce(h, scal + 2, 4, 0));
    double hs[8];
    B2_CUDA(cudaMemcpyAsync(hs, scal, 8 * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    float ms = 0;
    B2_CUDA(cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]));
    schur_ms += ms;
    if (fused) {
      B2_CUDA(cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]));
      solve_ms += ms;
    }
    const double gmax = hs[7];
    if (gmax <= opt->gradient_tolerance) {  // gradient_max_norm <= gradient_tolerance at the current iterate
      sum->termination_type = 0;
      break;
    }
    const double model_cost_change = hs[4];
    if (trace) fprintf(stderr, "[b2_ba] iter %d cost %.12e candidate %.12e model_change %.6e gmax %.3e radius %.3e info %d step^2 %.3e\n",
                       iter, cost, hs[5], model_cost_change, gmax, radius, info, hs[0] + hs[2]);
    const bool ok = (info == 0) && (model_cost_change > 0) && std::isfinite(model_cost_change);
    bool accepted = false;
    if (ok) {
      const double step = std::sqrt(hs[0] + hs[2]), xn = std::sqrt(hs[1] + hs[3]);
      if (step <= opt->parameter_tolerance * (xn + opt->parameter_tolerance)) {
        sum->termination_type = 0;
        break;
      }
      const double new_cost = hs[5];
      const double cost_change = cost - new_cost;
      // Ceres' TrustRegionMinimizer tests the function tolerance on every valid step before judging it; the candidate is
      // then not applied (function_tolerance = 0, the reference's setting, fires on an exactly unchanged cost)
      if (std::abs(cost_change) <= opt->function_tolerance * cost) {
        sum->termination_type = 0;
        break;
      }
      const double rho = cost_change / model_cost_change;
      if (rho > min_rel_dec) {
        accepted = true;
        std::swap(P.qvec, P.qvec_new);
        std::swap(P.tvec, P.tvec_new);
        std::swap(P.cam_params, P.cam_new);
        std::swap(P.xyz, P.xyz_new);
        sum->num_successful_steps++;
        const double t = 2.0 * rho - 1.0;
        radius = radius / std::max(1.0 / 3.0, 1.0 - t * t * t);
        radius = std::min(max_radius, radius);
        decrease_factor = 2.0;
        cost = new_cost;
        if (!fused) {  // the fused path re-evaluates the blocks inside its kernels
          B2_CUDA(cudaMemsetAsync(scal + 6, 0, 8, s));
          B2_CUDA(ba_launch_jacobian(P, P.qvec, P.tvec, P.cam_params, P.xyz, 0, scal + 6, s, loss_type, loss_scale));
          count_launches(1);
        }
      }
    }
    if (!accepted) {
      sum->num_unsuccessful_steps++;
      radius /= decrease_factor;
      decrease_factor *= 2.0;
      // TrustRegionMinimizer::MinTrustRegionRadiusReached (Ceres 1.14): "Minimum trust region radius reached" is
      // reported as CONVERGENCE, and is tested after the iteration limit
      if (radius <= min_radius) {
        sum->termination_type = (iter + 1 >= opt->max_num_iterations) ? 1 : 0;
        break;
      }
    }
  }
  B2_CUDA(cudaEventRecord(h->ev[1], s));
  // ---------------------------------------------------------------- download (in place)
  B2_CUDA(cudaMemcpyAsync(pr->qvec, P.qvec, (size_t)n_img * 4 * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(pr->tvec, P.tvec, (size_t)n_img * 3 * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(cam_dev.data(), P.cam_params, cam_dev.size() * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(pr->xyz, P.xyz, (size_t)n_pts * 3 * 8, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  for (int c = 0; c < n_cam; ++c)
    for (int k = 0; k < num_params(pr->camera_model[c]); ++k) pr->camera_params[(size_t)cs * c + k] = cam_dev[(size_t)KI * c + k];
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
  sum->solve_seconds = ms * 1e-3;
  sum->schur_kernel_seconds = schur_ms * 1e-3;
  sum->linear_solve_seconds = solve_ms * 1e-3;
  sum->schur_kernel_launches = schur_launches;
  sum->final_cost = cost;
  sum->num_iterations = sum->num_successful_steps + sum->num_unsuccessful_steps;
  return B2_OK;
}

}  // namespace

extern "C" {

void b2_ba_default_options(b2_ba_options* o) {
  if (!o) return;
  // DistributedMapperController::GlobalBundleAdjustment (distributed_mapper_controller.cpp:522-542)
  o->max_num_iterations = 50;
  o->refine_focal_length = 1;
  o->refine_principal_point = 0;
  o->refine_extra_params = 1;
  o->function_tolerance = 0.0;
  o->gradient_tolerance = 1.0;
  o->parameter_tolerance = 0.0;
  o->loss_function_type = 0;  // TRIVIAL
  o->linear_solver_type = 0;  // by the number of images, as BundleAdjuster::Solve chooses
  o->loss_function_scale = 1.0;
  o->max_linear_solver_iterations = 100;  // distributed_mapper_controller.cpp:529
  o->reserved = 0;
}

int b2_ba_create(int device, b2_ba** out) {
  if (!out) return set_error(B2_ERR_INVALID, "out == NULL");
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return set_error(B2_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
  }
  if (device < 0 || device >= n_dev) return set_error(B2_ERR_INVALID, "bad device ordinal");
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return set_error(B2_ERR_NO_DEVICE, "device is not sm_100");
  B2_CUDA(cudaSetDevice(device));
  b2_ba* h = new b2_ba();
  h->device = device;
  h->n_sm = prop.multiProcessorCount;
  const int rc = [&]() -> int {
    B2_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    if (cusolverDnCreate(&h->solver) != CUSOLVER_STATUS_SUCCESS) return set_error(B2_ERR_CUDA, "cusolverDnCreate failed");
    cusolverDnSetStream(h->solver, h->stream);
    if (cusolverDnCreateParams(&h->solver_params) != CUSOLVER_STATUS_SUCCESS) return set_error(B2_ERR_CUDA, "cusolverDnCreateParams failed");
    for (auto& e : h->ev) B2_CUDA(cudaEventCreate(&e));
    return B2_OK;
  }();
  if (rc != B2_OK) {  // a half-built handle is released here, never handed out
    b2_ba_destroy(h);
    return rc;
  }
  *out = h;
  return B2_OK;
}

int b2_nccl_unique_id(uint8_t* out128) {
  if (!out128) return set_error(B2_ERR_INVALID, "NULL argument");
  nccl::Api* a = nccl::api();
  if (!a) return set_error(B2_ERR_INVALID, "libnccl.so.2 not found (set B2_NCCL_LIBRARY)");
  nccl::UniqueId id;
  const int rc = a->GetUniqueId(&id);
  if (rc != 0) return set_error(B2_ERR_CUDA, a->GetErrorString ? a->GetErrorString(rc) : "ncclGetUniqueId failed");
  memcpy(out128, id.internal, 128);
  return B2_OK;
}

int b2_ba_init_nccl(b2_ba* h, int32_t n_ranks, int32_t rank, const uint8_t* id128) {
  if (!h || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return set_error(B2_ERR_INVALID, "bad argument");
  nccl::Api* a = nccl::api();
  if (!a) return set_error(B2_ERR_INVALID, "libnccl.so.2 not found (set B2_NCCL_LIBRARY)");
  B2_CUDA(cudaSetDevice(h->device));
  if (h->nccl_comm) { a->CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  nccl::UniqueId id;
  memcpy(id.internal, id128, 128);
  nccl::Comm c = nullptr;
  const int rc = a->CommInitRank(&c, n_ranks, id, rank);
  if (rc != 0) return set_error(B2_ERR_CUDA, a->GetErrorString ? a->GetErrorString(rc) : "ncclCommInitRank failed");
  h->nccl_comm = c;
  return B2_OK;
}

int b2_ba_destroy(b2_ba* h) {
  if (!h) return B2_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->nccl_comm) { if (nccl::Api* a = nccl::api()) a->CommDestroy(h->nccl_comm); h->nccl_comm = nullptr; }
  free_all(h);
  if (h->solver_params) cusolverDnDestroyParams(h->solver_params);
  if (h->solver) cusolverDnDestroy(h->solver);
  for (auto e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return B2_OK;
}

int b2_ba_set_allreduce(b2_ba* h, b2_allreduce_fn fn, void* user) {
  if (!h) return set_error(B2_ERR_INVALID, "NULL handle");
  h->allreduce = fn;
  h->allreduce_user = user;
  return B2_OK;
}

int b2_ba_reprojection_errors(b2_ba* h, const b2_ba_problem* pr, double* point_errors, double* mean_reprojection_error) {
  if (!h || !pr || !mean_reprojection_error) return set_error(B2_ERR_INVALID, "NULL argument");
  if (pr->n_images < 0 || pr->n_cameras < 0 || pr->n_points < 0 || pr->n_obs < 0) return set_error(B2_ERR_INVALID, "negative size");
  const int n_img = pr->n_images, n_cam = pr->n_cameras, n_pts = pr->n_points;
  const int64_t n_obs = pr->n_obs;
  const int cs = pr->camera_params_stride > 0 ? pr->camera_params_stride : 4;
  for (int64_t o = 0; o < n_obs; ++o) {
    if (pr->obs_image[o] < 0 || pr->obs_image[o] >= n_img || pr->obs_point[o] < 0 || pr->obs_point[o] >= n_pts)
      return set_error(B2_ERR_INVALID, "observation index out of range");
    if (o > 0 && pr->obs_point[o] < pr->obs_point[o - 1]) return set_error(B2_ERR_INVALID, "observations must be sorted by point");
  }
  for (int i = 0; i < n_img; ++i)
    if (pr->image_camera[i] < 0 || pr->image_camera[i] >= n_cam) return set_error(B2_ERR_INVALID, "bad image_camera");
  for (int c = 0; c < n_cam; ++c)
    if (pr->camera_model[c] < 0 || pr->camera_model[c] > 10 || num_params(pr->camera_model[c]) > cs)
      return set_error(B2_ERR_INVALID, "unknown camera model id or camera_params_stride too small");
  *mean_reprojection_error = 0.0;
  if (n_pts == 0 || n_obs == 0) return B2_OK;  // the reference divides 0 / 0 here; callers only log the value
  B2_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  std::vector<int64_t> pt_start(n_pts + 1, 0);
  for (int64_t o = 0; o < n_obs; ++o) pt_start[pr->obs_point[o] + 1]++;
  for (int p = 0; p < n_pts; ++p) pt_start[p + 1] += pt_start[p];
  auto run = [&]() -> int {
    int32_t *d_obs_img, *d_img_cam, *d_cam_model;
    int64_t* d_pt_start;
    double *d_obs_xy, *d_cam, *d_q, *d_t, *d_xyz, *d_err, *d_partial;
    const int n_blocks_max = (int)(((int64_t)n_pts * 32 + 255) / 256);
    B2_TRY(dev_upload(h, &d_obs_img, pr->obs_image, (size_t)n_obs));
    B2_TRY(dev_upload(h, &d_obs_xy, pr->obs_xy, (size_t)n_obs * 2));
    B2_TRY(dev_upload(h, &d_pt_start, pt_start.data(), pt_start.size()));
    B2_TRY(dev_upload(h, &d_img_cam, pr->image_camera, (size_t)n_img));
    B2_TRY(dev_upload(h, &d_cam_model, pr->camera_model, (size_t)n_cam));
    B2_TRY(dev_upload(h, &d_cam, (const double*)pr->camera_params, (size_t)n_cam * cs));
    B2_TRY(dev_upload(h, &d_q, (const double*)pr->qvec, (size_t)n_img * 4));
    B2_TRY(dev_upload(h, &d_t, (const double*)pr->tvec, (size_t)n_img * 3));
    B2_TRY(dev_upload(h, &d_xyz, (const double*)pr->xyz, (size_t)n_pts * 3));
    B2_TRY(dev_alloc(h, &d_err, (size_t)n_pts));
    B2_TRY(dev_alloc(h, &d_partial, (size_t)n_blocks_max));
    int n_blocks = 0;
    B2_CUDA(bam_launch_reprojection_errors(n_pts, d_pt_start, d_obs_img, d_obs_xy, d_img_cam, d_cam_model, d_cam, cs, d_q, d_t,
                                           d_xyz, d_err, d_partial, &n_blocks, s));
    count_launches(1);
    std::vector<double> partial((size_t)n_blocks);
    B2_CUDA(cudaMemcpyAsync(partial.data(), d_partial, (size_t)n_blocks * 8, cudaMemcpyDeviceToHost, s));
    if (point_errors) B2_CUDA(cudaMemcpyAsync(point_errors, d_err, (size_t)n_pts * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    double sum = 0;
    for (double v : partial) sum += v;
    *mean_reprojection_error = sum / (double)n_obs;  // total_reprojected_points = sum of track lengths
    return B2_OK;
  };
  const int rc = run();
  cudaStreamSynchronize(s);
  free_all(h);
  return rc;
}

// Test seam of the tiled Cholesky (ba_chol.cu): solves A x = b for a dense symmetric positive definite A (row-major,
// upper triangle read) through exactly the code path b2_ba_solve uses -- packed 64 x 64 tiles (here: every tile whose
// block of A holds a non-zero, plus the symbolic fill), panel / update kernels, the two triangular solves.
int b2_ba_debug_cholesky_solve(b2_ba* h, int64_t D, const double* A, const double* b, double* x, int32_t* info_out, int32_t* n_tiles_out) {
  if (!h || !A || !b || !x || D <= 0) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int rc = [&]() -> int {
    FusedPlan plan;
    const int nt = (int)((D + kST - 1) / kST);
    plan.nt = nt;
    plan.tmap.assign((size_t)nt * nt, 0);
    for (int64_t r = 0; r < D; ++r)
      for (int64_t c = r; c < D; ++c)
        if (A[r * D + c] != 0.0 || r == c) plan.tmap[(size_t)(r / kST) * nt + (c / kST)] = 1;
    finish_tiles(&plan);
    std::vector<double> tiles((size_t)plan.n_tiles * kST * kST, 0.0), rhs((size_t)nt * kST, 0.0);
    for (int64_t r = 0; r < (int64_t)nt * kST; ++r)
      for (int64_t c = r; c < (int64_t)nt * kST; ++c) {
        const int id = plan.tile_id[(size_t)(r / kST) * nt + (c / kST)];
        if (id < 0) continue;
        const double v = (r < D && c < D) ? A[r * D + c] : (r == c ? 1.0 : 0.0);
        tiles[(size_t)id * kST * kST + (r % kST) * kST + (c % kST)] = v;
      }
    for (int64_t r = 0; r < D; ++r) rhs[(size_t)r] = b[r];
    BaTiles T;
    memset(&T, 0, sizeof T);
    int32_t *d_tile_id, *d_row_ptr, *d_row_col;
    double* d_x;
    B2_TRY(dev_upload(h, &d_tile_id, (const int32_t*)plan.tile_id.data(), plan.tile_id.size()));
    B2_TRY(dev_upload(h, &d_row_ptr, (const int32_t*)plan.row_ptr.data(), plan.row_ptr.size()));
    B2_TRY(dev_upload(h, &d_row_col, (const int32_t*)plan.row_col.data(), plan.row_col.size()));
    B2_TRY(dev_upload(h, &T.tiles, (const double*)tiles.data(), tiles.size()));
    B2_TRY(dev_upload(h, &d_x, (const double*)rhs.data(), rhs.size()));
    B2_TRY(dev_alloc(h, &T.rinv, (size_t)nt * kST * kST));
    B2_TRY(dev_alloc(h, &T.info, 1));
    B2_CUDA(cudaMemsetAsync(T.info, 0, sizeof(int), s));
    T.nt = nt; T.n_tiles = plan.n_tiles; T.tile_id = d_tile_id; T.row_ptr = d_row_ptr; T.row_col = d_row_col;
    BaCholDev Cg;
    B2_TRY(upload_chol_graph(h, plan, T, &Cg));
    B2_CUDA(bac_solve_system(T, Cg, d_x, h->n_sm, s));
    count_launches(1);
    int info = 0;
    B2_CUDA(cudaMemcpyAsync(&info, T.info, sizeof(int), cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaMemcpyAsync(x, d_x, (size_t)D * 8, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    if (info_out) *info_out = info;
    if (n_tiles_out) *n_tiles_out = plan.n_tiles;
    return B2_OK;
  }();
  cudaStreamSynchronize(s);
  free_all(h);
  return rc;
}

int b2_ba_solve(b2_ba* h, const b2_ba_problem* pr, const b2_ba_options* opt, b2_ba_summary* sum) {
  if (!h || !pr || !opt || !sum) return set_error(B2_ERR_INVALID, "NULL argument");
  if (opt->loss_function_type < 0 || opt->loss_function_type > 2 || !(opt->loss_function_scale > 0))
    return set_error(B2_ERR_INVALID, "unknown loss_function_type or non-positive loss_function_scale");
  if (opt->linear_solver_type < 0 || opt->linear_solver_type > 2 || opt->max_linear_solver_iterations < 0)
    return set_error(B2_ERR_INVALID, "bad linear_solver_type / max_linear_solver_iterations");
  if (pr->n_images < 0 || pr->n_cameras < 0 || pr->n_points < 0 || pr->n_obs < 0 || opt->max_num_iterations < 0)
    return set_error(B2_ERR_INVALID, "negative size");
  B2_CUDA(cudaSetDevice(h->device));
  const int rc = solve_impl(h, pr, opt, sum);
  cudaStreamSynchronize(h->stream);
  free_all(h);
  return rc;
}

}  // extern "C"
