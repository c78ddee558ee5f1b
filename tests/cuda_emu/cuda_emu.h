// cuda_emu.h -- a small CUDA-on-CPU execution model for the CPU test-suite.  TEST INFRASTRUCTURE ONLY:
// nothing under dagsfm_b200/ includes it, the product has no CPU path.
//
// tests/cuda_emu/build_emu.py rewrites the `kernel<<<grid, block, smem, stream>>>(args)` statements of the
// library's .cu files into cuda_emu::launch(...) calls and compiles the SAME sources with g++ against
// this header, so that the kernels (minus the tcgen05 matcher, which has no CPU meaning) and the host
// code around them can be exercised against the oracle without a GPU.
//
// Execution model: blocks run one after the other; the threads of a block are fibers (own stacks, a hand-written context
// switch) on the calling OS thread, scheduled round-robin and switched only at synchronisation points
// (__syncthreads, __syncwarp, warp shuffles / votes), i.e. deterministically.  A fiber that returns
// simply leaves its barriers (like an exited CUDA thread).  Device memory is host memory.
#pragma once

#include <execinfo.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

// ------------------------------------------------------------------ qualifiers / built-in types
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ /* must stay empty: libstdc++ spells __attribute__((__noinline__)) */
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static const
#define __shared__ static
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

struct uint3_emu { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3() {}
  dim3(unsigned X, unsigned Y = 1, unsigned Z = 1) : x(X), y(Y), z(Z) {}
};
// alignment as in vector_types.h: a misaligned vector access faults on the device; here B2_EMU_UBSAN=1 (build_emu.py)
// turns it into a trap
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
constexpr int warpSize = 32;
// CUDA's global-namespace min / max (mixed integer types allowed)
template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) {
  typedef typename std::common_type<A, B>::type T;
  return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) {
  typedef typename std::common_type<A, B>::type T;
  return (T)a > (T)b ? (T)a : (T)b;
}

#if !defined(__x86_64__)
#error "cuda_emu.h switches fibers with a few lines of x86-64 assembly"
#endif
// cuda_emu_switch(&save_sp, load_sp): pushes the callee-saved registers, stores the stack pointer, adopts
// load_sp, pops that context's registers and returns into it.  (glibc's swapcontext costs a system call
// per switch and longjmp across stacks trips its cleanup-handler bookkeeping.)
extern "C" void cuda_emu_switch(void** save_sp, void* load_sp);
asm(R"(
.pushsection .text
.weak cuda_emu_switch
.type cuda_emu_switch,@function
cuda_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cuda_emu_switch,.-cuda_emu_switch
.popsection
)");

namespace cuda_emu {

struct Fiber {
  void* sp = nullptr;  // saved stack pointer while the fiber is switched out
  std::vector<char> stack;
  bool done = false;
  int wait_kind = 0;  // 0 runnable, 1 waiting at the block barrier, 2 waiting at its warp barrier
  unsigned tid = 0;
  int site = 0;        // which warp-synchronous call of the source the fiber waits in (see Site)
  const char* site_file = "";
  int site_line = 0;
};

// The source location of a warp-synchronous intrinsic.  Lanes released together must wait in the SAME call: lanes
// that pair a shuffle of one statement with a ballot (or a shuffle) of another exchange garbage on the device too, and
// a host build would hide it.
struct Site { const char* file; int line; int id; };

struct BlockState {
  std::vector<Fiber> fibers;
  void* sched_sp = nullptr;
  int current = -1;
  unsigned n_threads = 0;
  // warp exchange buffers (one 64-bit slot per lane) and arrival counters
  std::vector<uint64_t> slot;
  std::vector<int> warp_arrived;
  int block_arrived = 0;
  std::function<void()> body;
};

inline BlockState*& cur() {
  static thread_local BlockState* b = nullptr;
  return b;
}
struct Idx { uint3_emu threadIdx, blockIdx; dim3 blockDim, gridDim; };
inline Idx& idx() {
  static thread_local Idx i;
  return i;
}
inline void*& dyn_smem() {
  static thread_local void* p = nullptr;
  return p;
}

inline void yield_to_scheduler() {
  BlockState* b = cur();
  Fiber& f = b->fibers[b->current];
  cuda_emu_switch(&f.sp, b->sched_sp);
}
inline int alive_in_warp(BlockState* b, unsigned warp) {
  int n = 0;
  for (unsigned t = warp * 32; t < std::min(b->n_threads, warp * 32 + 32); ++t) n += b->fibers[t].done ? 0 : 1;
  return n;
}
inline int alive_in_block(BlockState* b) {
  int n = 0;
  for (auto& f : b->fibers) n += f.done ? 0 : 1;
  return n;
}
// Blocks the calling fiber until every live fiber of its warp has arrived.
inline void warp_barrier(Site at = Site{"", 0, 0}, int phase = 0) {
  BlockState* b = cur();
  Fiber& f = b->fibers[b->current];
  f.site = at.id * 2 + phase;
  f.site_file = at.file;
  f.site_line = at.line;
  f.wait_kind = 2;
  b->warp_arrived[f.tid / 32] += 1;
  yield_to_scheduler();
}
inline void block_barrier() {
  BlockState* b = cur();
  Fiber& f = b->fibers[b->current];
  f.wait_kind = 1;
  b->block_arrived += 1;
  yield_to_scheduler();
}

inline void fiber_entry() {
  BlockState* b = cur();
  b->body();
  Fiber& f = b->fibers[b->current];
  f.done = true;
  cuda_emu_switch(&f.sp, b->sched_sp);  // never resumed
  std::abort();
}

inline void run_block(BlockState& b, unsigned n_threads, const std::function<void()>& body) {
  constexpr size_t kStack = 256u << 10;  // the verification kernel keeps several 10 KB of locals
  b.n_threads = n_threads;
  b.body = body;
  if (b.fibers.size() < n_threads) b.fibers.resize(n_threads);
  b.slot.assign(n_threads, 0);
  b.warp_arrived.assign((n_threads + 31) / 32, 0);
  b.block_arrived = 0;
  cur() = &b;
  for (unsigned t = 0; t < n_threads; ++t) {
    Fiber& f = b.fibers[t];
    if (f.stack.size() != kStack) f.stack.resize(kStack);
    f.done = false;
    f.wait_kind = 0;
    f.tid = t;
    // initial frame: six zeroed callee-saved registers, then the entry point as return address; after the
    // `ret` the stack pointer is 8 modulo 16, as at any function entry
    uintptr_t top = ((uintptr_t)f.stack.data() + f.stack.size()) & ~(uintptr_t)15;
    void** sp = (void**)(top - 8);
    *--sp = (void*)+[] { fiber_entry(); };
    for (int r = 0; r < 6; ++r) *--sp = nullptr;
    f.sp = sp;
  }
  // B2_EMU_SCHED=reverse runs the threads of a block (and the blocks of a grid) last to first: results that depend on
  // the order in which warps or blocks happen to run -- a missing barrier, a read of another block's output -- change
  static const bool reverse = [] { const char* e = std::getenv("B2_EMU_SCHED"); return e && std::strcmp(e, "reverse") == 0; }();
  for (;;) {
    bool progressed = false, any_alive = false;
    for (unsigned k = 0; k < n_threads; ++k) {
      const unsigned t = reverse ? n_threads - 1 - k : k;
      Fiber& f = b.fibers[t];
      if (f.done) continue;
      any_alive = true;
      if (f.wait_kind != 0) continue;
      b.current = (int)t;
      idx().threadIdx.x = t;
      cuda_emu_switch(&b.sched_sp, f.sp);
      progressed = true;
    }
    if (!any_alive) break;
    // release barriers whose live participants have all arrived
    for (unsigned w = 0; w < b.warp_arrived.size(); ++w) {
      const int alive = alive_in_warp(&b, w);
      if (b.warp_arrived[w] > 0 && b.warp_arrived[w] >= alive) {
        b.warp_arrived[w] = 0;
        const Fiber* first = nullptr;
        for (unsigned t = w * 32; t < std::min(n_threads, w * 32 + 32); ++t)
          if (!b.fibers[t].done && b.fibers[t].wait_kind == 2) {
            const Fiber& g = b.fibers[t];
            if (!first) first = &g;
            if (g.site != first->site) {
              std::fprintf(stderr, "cuda_emu: lanes %u and %u of warp %u meet in different warp-synchronous calls (%s:%d and %s:%d)\n",
                           first->tid % 32, g.tid % 32, w, first->site_file, first->site_line, g.site_file, g.site_line);
              std::abort();
            }
            b.fibers[t].wait_kind = 0;
            progressed = true;
          }
      }
    }
    if (b.block_arrived > 0 && b.block_arrived >= alive_in_block(&b)) {
      b.block_arrived = 0;
      for (auto& f : b.fibers)
        if (!f.done && f.wait_kind == 1) { f.wait_kind = 0; progressed = true; }
    }
    if (!progressed) {
      std::fprintf(stderr, "cuda_emu: deadlock (threads wait at different barriers)\n");
      std::abort();
    }
  }
  cur() = nullptr;
}

// kernel launch: blocks sequentially, threads as fibers.  `smem_bytes` backs `extern __shared__`.
// B2_EMU_DEVMEM=1 (with guarded allocations): "device" memory is readable and writable only while a kernel, a cudaMemcpy /
// cudaMemset or a cuSOLVER call runs; host code that dereferences a device pointer faults as it would on the device
inline void device_access(bool on);
struct DeviceWindow {
  DeviceWindow() { device_access(true); }
  ~DeviceWindow() { device_access(false); }
};

// cudaGetLastError(): a launch whose configuration the driver would refuse (an empty grid or block, more than 1024
// threads, grid.y / grid.z beyond 65535, more dynamic shared memory than an sm_100 block can opt into) does not run
// and leaves cudaErrorInvalidConfiguration (9) behind, as on the device
inline int& last_error() { static thread_local int e = 0; return e; }

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  static thread_local BlockState state;
  {
    const unsigned long long threads = (unsigned long long)block.x * block.y * block.z;
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || threads == 0 || threads > 1024 || block.z > 64 ||
        grid.x > 2147483647u || grid.y > 65535u || grid.z > 65535u || smem_bytes > 232448) {
      last_error() = 9;
      return;
    }
  }
  std::vector<uint64_t> smem((smem_bytes + 7) / 8 + 1, ~0ull);  // shared memory is not cleared at launch either
  dyn_smem() = smem.data();
  idx().gridDim = grid;
  idx().blockDim = block;
  static const bool reverse = [] { const char* e = std::getenv("B2_EMU_SCHED"); return e && std::strcmp(e, "reverse") == 0; }();
  DeviceWindow window;
  for (unsigned kz = 0; kz < grid.z; ++kz)
    for (unsigned ky = 0; ky < grid.y; ++ky)
      for (unsigned kx = 0; kx < grid.x; ++kx) {
        const unsigned bx = reverse ? grid.x - 1 - kx : kx, by = reverse ? grid.y - 1 - ky : ky, bz = reverse ? grid.z - 1 - kz : kz;
        idx().blockIdx.x = bx; idx().blockIdx.y = by; idx().blockIdx.z = bz;
        run_block(state, block.x * block.y * block.z, body);
      }
  dyn_smem() = nullptr;
}

}  // namespace cuda_emu

#define threadIdx (cuda_emu::idx().threadIdx)
#define blockIdx (cuda_emu::idx().blockIdx)
#define blockDim (cuda_emu::idx().blockDim)
#define gridDim (cuda_emu::idx().gridDim)

// ------------------------------------------------------------------ synchronisation / warp intrinsics
inline void __syncthreads() { cuda_emu::block_barrier(); }
namespace cuda_emu {
struct SyncWarp {  // __syncwarp() and __syncwarp(mask): a functor, so the macro below needs no comma tricks
  Site at;
  void operator()(unsigned = 0xffffffffu) const { warp_barrier(at); }
};
template <class T>
inline T warp_exchange(Site at, T v, int src_lane_of_me) {  // every live lane publishes v, then reads the lane it asks for
  static_assert(sizeof(T) <= 8, "warp_exchange moves at most 64 bits");
  BlockState* b = cur();
  const unsigned tid = b->fibers[b->current].tid;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  b->slot[tid] = bits;
  warp_barrier(at, 0);
  const unsigned base = tid / 32 * 32;
  unsigned src = base + (unsigned)(src_lane_of_me & 31);
  if (src >= b->n_threads) src = tid;
  const uint64_t got = b->slot[src];
  warp_barrier(at, 1);  // nobody overwrites its slot before everyone has read
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T> inline T shfl_at(Site at, unsigned, T v, int src, int = 32) { return warp_exchange(at, v, src); }
template <class T> inline T shfl_xor_at(Site at, unsigned, T v, int mask, int = 32) {
  return warp_exchange(at, v, (int)(threadIdx.x & 31) ^ mask);
}
template <class T> inline T shfl_up_at(Site at, unsigned, T v, unsigned d, int = 32) {
  const int lane = (int)(threadIdx.x & 31);
  return warp_exchange(at, v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> inline T shfl_down_at(Site at, unsigned, T v, unsigned d, int = 32) {
  const int lane = (int)(threadIdx.x & 31);
  return warp_exchange(at, v, lane + (int)d < 32 ? lane + (int)d : lane);
}
inline unsigned ballot_at(Site at, unsigned, int pred) {
  BlockState* b = cur();
  const unsigned tid = b->fibers[b->current].tid;
  b->slot[tid] = pred ? 1u : 0u;
  warp_barrier(at, 0);
  unsigned m = 0;
  const unsigned base = tid / 32 * 32;
  for (unsigned l = 0; l < 32 && base + l < b->n_threads; ++l)
    if (!b->fibers[base + l].done && b->slot[base + l]) m |= 1u << l;
  warp_barrier(at, 1);
  return m;
}
inline int any_at(Site at, unsigned m, int p) { return ballot_at(at, m, p) != 0; }
}  // namespace cuda_emu

// every textual call gets its own Site (__COUNTER__ is unique within a translation unit; 0 is "untagged")
#define CUDA_EMU_SITE (cuda_emu::Site{__FILE__, __LINE__, __COUNTER__ + 1})
#define __syncwarp(...) (cuda_emu::SyncWarp{CUDA_EMU_SITE})(__VA_ARGS__)
#define __shfl_sync(...) cuda_emu::shfl_at(CUDA_EMU_SITE, __VA_ARGS__)
#define __shfl_xor_sync(...) cuda_emu::shfl_xor_at(CUDA_EMU_SITE, __VA_ARGS__)
#define __shfl_up_sync(...) cuda_emu::shfl_up_at(CUDA_EMU_SITE, __VA_ARGS__)
#define __shfl_down_sync(...) cuda_emu::shfl_down_at(CUDA_EMU_SITE, __VA_ARGS__)
#define __ballot_sync(...) cuda_emu::ballot_at(CUDA_EMU_SITE, __VA_ARGS__)
#define __any_sync(...) cuda_emu::any_at(CUDA_EMU_SITE, __VA_ARGS__)
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline float __frcp_rn(float x) { return 1.0f / x; }   // correctly rounded on the device too
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline long long clock64() { return 0; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() {}   // one OS thread runs the whole grid: program order is memory order
inline void __nanosleep(unsigned) {}
inline long long __double_as_longlong(double d) { long long r; std::memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; std::memcpy(&r, &v, 8); return r; }
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int q = 0; q < 4; ++q) c += ((a >> (8 * q)) & 255u) * ((b >> (8 * q)) & 255u);
  return c;
}
inline int __dp4a(int a, int b, int c) {
  for (int q = 0; q < 4; ++q) c += (int)(signed char)((a >> (8 * q)) & 255) * (int)(signed char)((b >> (8 * q)) & 255);
  return c;
}

// ------------------------------------------------------------------ atomics (fibers of one OS thread: plain RMW;
// std::atomic_ref keeps them correct should blocks ever run on several OS threads)
template <class T> inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned* p, int v) { auto o = *p; *p = o + (unsigned)v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <class T> inline T atomicMin(T* p, T v) { T old = *p; if (v < old) *p = v; return old; }
template <class T> inline T atomicExch(T* p, T v) { T old = *p; *p = v; return old; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { T old = *p; if (old == cmp) *p = v; return old; }

// ------------------------------------------------------------------ runtime API (device memory = host memory)
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef struct emu_stream_* cudaStream_t;
struct emu_event_ { std::chrono::steady_clock::time_point t; };
typedef emu_event_* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
constexpr unsigned cudaStreamNonBlocking = 1;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 8;
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 2; char name[64] = "cuda_emu"; size_t totalGlobalMem = (size_t)64 << 30; };

inline cudaError_t cudaGetLastError() { const int e = cuda_emu::last_error(); cuda_emu::last_error() = 0; return e; }
inline cudaError_t cudaPeekAtLastError() { return cuda_emu::last_error(); }
inline const char* cudaGetErrorString(cudaError_t e) {
  return e == 9 ? "invalid configuration argument (cuda_emu)" : e == 2 ? "out of memory (cuda_emu)" : "cuda_emu: no error";
}
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
// Asynchronous device-to-host copies are DELIVERED at the next synchronisation point (the data is taken at enqueue time,
// in stream order): host code that reads the destination before it synchronised sees the old contents, as it may with
// pinned memory on the device.
namespace cuda_emu {
struct PendingCopy { void* dst; std::vector<char> data; };
inline std::vector<PendingCopy>& pending_copies() { static std::vector<PendingCopy> v; return v; }
inline void deliver_pending_copies() {
  for (auto& c : pending_copies()) std::memmove(c.dst, c.data.data(), c.data.size());
  pending_copies().clear();
}
}  // namespace cuda_emu
inline cudaError_t cudaDeviceSynchronize() { cuda_emu::deliver_pending_copies(); return cudaSuccess; }
// B2_EMU_GUARD=1: every "device" allocation ends right before an inaccessible page, so an overrun of the kind
// compute-sanitizer reports on the GPU faults at the offending access (debugging aid for the emulator).
namespace cuda_emu {
inline void guard_segv(int, siginfo_t* si, void*) {
  void* bt[48];
  const int n = backtrace(bt, 48);
  std::fprintf(stderr, "cuda_emu: access violation at %p (B2_EMU_GUARD)\n", si->si_addr);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
inline bool guard_mode() {
  static const bool g = [] {
    const bool on = std::getenv("B2_EMU_GUARD") != nullptr;
    if (on) {
      struct sigaction sa;
      std::memset(&sa, 0, sizeof sa);
      sa.sa_sigaction = guard_segv;
      sa.sa_flags = SA_SIGINFO;
      sigaction(SIGSEGV, &sa, nullptr);
    }
    return on;
  }();
  return g;
}
inline std::unordered_map<void*, std::pair<void*, size_t>>& guard_registry() {
  static std::unordered_map<void*, std::pair<void*, size_t>> r;
  return r;
}
inline bool devmem_mode() {
  static const bool on = std::getenv("B2_EMU_DEVMEM") != nullptr && std::getenv("B2_EMU_GUARD") != nullptr;
  return on;
}
inline int& device_depth() { static int d = 0; return d; }
inline void device_access(bool on) {
  if (!devmem_mode()) return;
  int& d = device_depth();
  if (on ? d++ == 0 : --d == 0)
    for (auto& kv : guard_registry())
      mprotect(kv.second.first, kv.second.second - 4096, on ? PROT_READ | PROT_WRITE : PROT_NONE);
}
inline void* guarded_alloc(size_t n) {
  const size_t page = 4096, body = (n + 15) / 16 * 16, span = (body + page - 1) / page * page + page;
  char* base = (char*)mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (base == (char*)MAP_FAILED) return nullptr;
  mprotect(base + span - page, page, PROT_NONE);
  void* p = base + span - page - body;   // the allocation ends where the inaccessible page begins
  guard_registry()[p] = {base, span};
  // cudaMalloc does not clear: doubles start as NaN, indices as -1, not as zero (the tail of a very large allocation, to
  // bound the cost of touching pages nobody may use: the allocation's last 32 MiB and its first 1 MiB)
  {
    const size_t len = span - page, cap = (size_t)32 << 20;
    if (len <= cap + (1u << 20)) std::memset(base, 0xFF, len);
    else {
      std::memset(base, 0xFF, 1u << 20);
      std::memset(base + len - cap, 0xFF, cap);
    }
  }
  if (devmem_mode() && device_depth() == 0) mprotect(base, span - page, PROT_NONE);
  return p;
}
inline void guarded_free(void* p) {
  auto it = guard_registry().find(p);
  if (it == guard_registry().end()) return;
  munmap(it->second.first, it->second.second);
  guard_registry().erase(it);
}
}  // namespace cuda_emu
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) {
  n = std::max<size_t>(n, 1);
  *p = (T*)(cuda_emu::guard_mode() ? cuda_emu::guarded_alloc(n) : std::malloc(n));
  if (*p && !cuda_emu::guard_mode()) std::memset(*p, 0xFF, n);
  return *p ? cudaSuccess : 2;
}
inline cudaError_t cudaFree(void* p) {
  if (cuda_emu::guard_mode()) cuda_emu::guarded_free(p); else std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  cuda_emu::DeviceWindow window;
  cuda_emu::deliver_pending_copies();
  if (n) std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) {
  if (k != cudaMemcpyDeviceToHost) {
    cuda_emu::DeviceWindow window;
    if (n) std::memmove(d, s, n);
    return cudaSuccess;
  }
  cuda_emu::DeviceWindow window;
  cuda_emu::PendingCopy c{d, std::vector<char>((const char*)s, (const char*)s + n)};
  cuda_emu::pending_copies().push_back(std::move(c));
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  cuda_emu::DeviceWindow window;
  if (n) std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { cuda_emu::deliver_pending_copies(); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event_(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { cuda_emu::deliver_pending_copies(); return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::max(1e-3f, std::chrono::duration<float, std::milli>(b->t - a->t).count());
  return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)64 << 30; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ------------------------------------------------------------------ cuSOLVER dense Cholesky (column-major, 64-bit API)
typedef struct emu_solver_* cusolverDnHandle_t;
typedef struct emu_params_* cusolverDnParams_t;
typedef int cusolverStatus_t;
constexpr cusolverStatus_t CUSOLVER_STATUS_SUCCESS = 0;
enum cublasFillMode_t { CUBLAS_FILL_MODE_LOWER = 0, CUBLAS_FILL_MODE_UPPER = 1 };
enum cudaDataType { CUDA_R_64F = 1 };
inline cusolverStatus_t cusolverDnCreate(cusolverDnHandle_t* h) { *h = nullptr; return 0; }
inline cusolverStatus_t cusolverDnDestroy(cusolverDnHandle_t) { return 0; }
inline cusolverStatus_t cusolverDnSetStream(cusolverDnHandle_t, cudaStream_t) { return 0; }
inline cusolverStatus_t cusolverDnCreateParams(cusolverDnParams_t* p) { *p = nullptr; return 0; }
inline cusolverStatus_t cusolverDnDestroyParams(cusolverDnParams_t) { return 0; }
inline cusolverStatus_t cusolverDnXpotrf_bufferSize(cusolverDnHandle_t, cusolverDnParams_t, cublasFillMode_t, int64_t, cudaDataType,
                                                    const void*, int64_t, cudaDataType, size_t* dev, size_t* host) {
  *dev = 8; *host = 8; return 0;
}
// A = L L^T, lower triangle of a column-major n x n matrix, in place; info = 0 or the failing column (1-based).
inline cusolverStatus_t cusolverDnXpotrf(cusolverDnHandle_t, cusolverDnParams_t, cublasFillMode_t uplo, int64_t n, cudaDataType,
                                         void* Av, int64_t lda, cudaDataType, void*, size_t, void*, size_t, int* info) {
  cuda_emu::DeviceWindow window;
  double* A = (double*)Av;
  *info = 0;
  if (uplo != CUBLAS_FILL_MODE_LOWER) return 1;
  for (int64_t j = 0; j < n; ++j) {
    double d = A[j + j * lda];
    for (int64_t k = 0; k < j; ++k) d -= A[j + k * lda] * A[j + k * lda];
    if (!(d > 0)) { *info = (int)(j + 1); return 0; }
    d = std::sqrt(d);
    A[j + j * lda] = d;
    for (int64_t i = j + 1; i < n; ++i) {
      double s = A[i + j * lda];
      for (int64_t k = 0; k < j; ++k) s -= A[i + k * lda] * A[j + k * lda];
      A[i + j * lda] = s / d;
    }
  }
  return 0;
}
inline cusolverStatus_t cusolverDnXpotrs(cusolverDnHandle_t, cusolverDnParams_t, cublasFillMode_t, int64_t n, int64_t nrhs, cudaDataType,
                                         const void* Av, int64_t lda, cudaDataType, void* Bv, int64_t ldb, int* info) {
  cuda_emu::DeviceWindow window;
  const double* A = (const double*)Av;
  double* B = (double*)Bv;
  *info = 0;
  for (int64_t r = 0; r < nrhs; ++r) {
    double* b = B + r * ldb;
    for (int64_t i = 0; i < n; ++i) {
      double s = b[i];
      for (int64_t k = 0; k < i; ++k) s -= A[i + k * lda] * b[k];
      b[i] = s / A[i + i * lda];
    }
    for (int64_t i = n - 1; i >= 0; --i) {
      double s = b[i];
      for (int64_t k = i + 1; k < n; ++k) s -= A[k + i * lda] * b[k];
      b[i] = s / A[i + i * lda];
    }
  }
  return 0;
}
