// Host build of the device solvers (verify_solvers.cuh is plain C++ once the CUDA qualifiers are
// defined away): lets the CPU test-suite compare the exact solver source the kernel runs with the
// oracle, without a GPU.  Test infrastructure only.
#include <cmath>
#include <cstdint>
#include <cstdio>
#define __device__
#define __constant__ static
#define __forceinline__ inline
#define __noinline__
using std::fabs;
using std::sqrt;
#include "../../dagsfm_b200/csrc/verify_solvers.cuh"

static double ws[256];
extern "C" int host_f7(const double* p1, const double* p2, double* m) {
  return b2::vf::solve_f7(b2::vf::View<1>{ws}, p1, p2, m);
}
extern "C" int host_e5(const double* p1, const double* p2, double* m) {
  return b2::vf::solve_e5(b2::vf::View<1>{ws}, p1, p2, m);
}
extern "C" int host_h4(const double* p1, const double* p2, double* m) {
  return b2::vf::solve_h4(b2::vf::View<1>{ws}, p1, p2, m);
}
extern "C" int host_roots(const double* c, int nc, double* r) { return b2::vf::real_roots<10>(c, nc, r); }
