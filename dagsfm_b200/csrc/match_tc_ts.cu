// Descriptor matching, stage 1: dense u8 x u8 -> s32 contraction on tcgen05 + fused chunked top-2
// epilogue, with the stationary operand held in TENSOR MEMORY.
//
// Why TS mode: cycle counters in the first-generation kernel (both operands in shared memory, round 1) showed the
// MMA-issuing thread waiting on barriers only ~15 % of the time while the epilogue warps starved ~70 %: the SS-mode MMA
// (both operands in shared memory) reads 8 KB of smem per 64-cycle M128 N128 K32 step, which
// is exactly the 128 B/clk shared-memory port -- plus the TMA writes.  The query supertile X is
// reused by every block of an item, so it is written once per item into TMEM (tcgen05.st by the
// epilogue warps, double buffered) and the MMAs run in TS mode (A from TMEM): the smem port then
// carries only the streamed Y operand (64 B/clk) and the TMA fill (32 B/clk).
//
// Measured on B200 (micro-benchmark, 148 SMs, junk operands): one tcgen05.mma M128 x N x K32B
// takes 64 cycles for any N <= 128 (72.5 in chains of 4 K-steps, 128 for N = 256), identically for
// kind::i8 and kind::f16 and for SS and TS operands -- so N = 128 is the efficient shape and the
// smem port, not the tensor pipe, is what TS mode relieves.
//
// TMEM budget (512 columns): accumulators 2 tiles x N=128 columns = 256, SINGLE buffered per tile --
// the two tiles ping-pong instead: while the MMAs of tile 1 run, the epilogue warps of tile 0 drain
// their accumulator (one LDTM.x128 ~ 120 cycles < 4 MMAs ~ 290 cycles) and hand it back;
// A operand 2 buffers x 2 tiles x 32 columns (128 K-bytes / 4) = 128.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "match_common.cuh"
#include "ptx.cuh"

namespace b2 {
namespace ts {

constexpr int kN = 128;                     // UMMA N = Y rows per block
constexpr int kStagesY = 8;
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kNumEpiWarps;     // 320
constexpr uint32_t kYBytes = kN * kDescBytes;        // 16 KiB per stage
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol = 0;                      // accumulator of tile t at t * 128
constexpr uint32_t kACol = 256;                      // A operand (abuf, tile) at 256 + (2*abuf + tile) * 32
constexpr uint32_t kSmemY = 0;
constexpr uint32_t kSmemBar = kStagesY * kYBytes;
constexpr uint32_t kNumBars = 2 * kStagesY + 8;      // y_full/empty, a_full[2], a_empty[2], t_full[2], t_empty[2]
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;

__device__ __forceinline__ int max32(const uint32_t* v) {
  int m[11];
#pragma unroll
  for (int i = 0; i < 10; ++i)
    m[i] = __vimax3_s32((int)v[3 * i], (int)v[3 * i + 1], (int)v[3 * i + 2]);
  m[10] = max((int)v[30], (int)v[31]);
  int a = __vimax3_s32(m[0], m[1], m[2]);
  int b = __vimax3_s32(m[3], m[4], m[5]);
  int c = __vimax3_s32(m[6], m[7], m[8]);
  int d = max(m[9], m[10]);
  return max(__vimax3_s32(a, b, c), d);
}

// EXP != 0 are timing experiments (results invalid): 1 = epilogue without LDTM and ALU, 2 = LDTM without ALU.  They
// exist only in builds with -DB2_MATCH_EXPERIMENTS (never in the shipped library); production is EXP == 0.
template <int EXP>
__global__ void __launch_bounds__(kThreads, 1)
match_top2_ts_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ pool,
                     const MatchItem* __restrict__ items, const uint32_t* __restrict__ n_items_ptr,
                     int thr_dist, const int* __restrict__ ratio_lim, int* __restrict__ midx,
                     uint4* __restrict__ cands, unsigned int* __restrict__ cand_count,
                     unsigned int cand_capacity) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kSmemBar;
  auto y_full = [&](uint32_t s) { return bar_base + 8 * s; };
  auto y_empty = [&](uint32_t s) { return bar_base + 8 * (kStagesY + s); };
  auto a_full = [&](uint32_t a) { return bar_base + 8 * (2 * kStagesY + a); };
  auto a_empty = [&](uint32_t a) { return bar_base + 8 * (2 * kStagesY + 2 + a); };
  auto t_full = [&](uint32_t t) { return bar_base + 8 * (2 * kStagesY + 4 + t); };
  auto t_empty = [&](uint32_t t) { return bar_base + 8 * (2 * kStagesY + 6 + t); };
  const uint32_t tmem_slot = bar_base + 8 * kNumBars;

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t n_items = *n_items_ptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (uint32_t s = 0; s < kStagesY; ++s) {
      mbar_init(y_full(s), 1);
      mbar_init(y_empty(s), 1);
    }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(a_full(s), kNumEpiWarps);
      mbar_init(a_empty(s), 1);
      mbar_init(t_full(s), 1);
      mbar_init(t_empty(s), kNumEpiWarps / 2);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===================================================== TMA producer (Y blocks only)
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const MatchItem w = items[item];
        for (uint32_t b = 0; b < w.y_nblk; ++b) {
          const uint32_t s = it % kStagesY;
          mbar_wait(y_empty(s), ((it / kStagesY) & 1) ^ 1);
          mbar_arrive_expect_tx(y_full(s), kYBytes);
          tma_load_2d(smem_base + kSmemY + s * kYBytes, &tmap, 0, (int)(w.y_row + b * kN), y_full(s));
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer (TS mode, whole warp converged)
    {
      constexpr uint32_t idesc = make_idesc_u8_s32(kTileRows, kN);
      uint32_t it = 0, xi = 0, tb = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t y_nblk = items[item].y_nblk;
        const uint32_t ab = xi & 1;
        mbar_wait(a_full(ab), (xi >> 1) & 1);
        tc_fence_after();
        for (uint32_t b = 0; b < y_nblk; ++b) {
          const uint32_t s = it % kStagesY;
          mbar_wait(y_full(s), (it / kStagesY) & 1);
          const uint32_t ya = smem_base + kSmemY + s * kYBytes;
#pragma unroll
          for (uint32_t t = 0; t < 2; ++t) {
            mbar_wait(t_empty(t), (tb & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + kAccCol + t * kN;
            const uint32_t a = tmem_base + kACol + (2 * ab + t) * 32;
            if (elect_one()) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) {
                const uint64_t bd = make_kmajor_sw128_desc(ya + k * 32);
                mma_i8_ts(d, a + k * 8, bd, idesc, k);  // K = 32 bytes = 8 TMEM columns of A
              }
              tc_commit(t_full(t));
              if (t == 1) tc_commit(y_empty(s));
            }
            __syncwarp();
          }
          ++it;
          ++tb;
        }
        if (elect_one()) tc_commit(a_empty(ab));  // this item's MMAs have consumed the A buffer
        __syncwarp();
        ++xi;
      }
    }
  } else {
    // ================================================= epilogue (+ A-operand loader)
    const int ew = warp - 2;
    const uint32_t quad = warp & 3;
    const uint32_t tile = ew >> 2;
    const uint32_t lane_addr = (quad * 32u) << 16;
    const uint32_t row_in_item = tile * 128 + quad * 32 + lane;
    // writes this thread's query row (128 bytes = 32 columns) of `item` into A buffer `ab`
    auto load_a = [&](uint32_t item, uint32_t xi) {
      const uint32_t ab = xi & 1;
      mbar_wait(a_empty(ab), ((xi >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint4* src = reinterpret_cast<const uint4*>(pool + (size_t)(items[item].x_row + row_in_item) * kDescBytes);
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(src + q);
        r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
      }
      tmem_st_32x32b_x32_wait(tmem_base + lane_addr + kACol + (2 * ab + tile) * 32, r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(ab));
    };
    uint32_t tb = 0, xi = 0;
    if (blockIdx.x < n_items) load_a(blockIdx.x, 0);
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t y_nblk = items[item].y_nblk;
      // prefetch the next item's query rows into the other A buffer
      if (item + gridDim.x < n_items) load_a(item + gridDim.x, xi + 1);
      int best = 0, second = 0, bchunk = 0;
      for (uint32_t b = 0; b < y_nblk; ++b) {
        mbar_wait(t_full(tile), tb & 1);
        tc_fence_after();
        uint32_t v[128];
        if (EXP != 1) tmem_ld_32x32b_x128_wait(tmem_base + lane_addr + kAccCol + tile * kN, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(tile));
        if (EXP == 1) { best += (int)tb; }
        else if (EXP == 2) { best = max(best, (int)v[0]); }
        else
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int m = max32(v + 32 * c);
          const int chunk = (int)(b * 4 + c);
          second = max(second, min(best, m));
          bchunk = (m > best) ? chunk : bchunk;
          best = max(best, m);
        }
        ++tb;
      }
      const uint32_t out = item * kSuperRows + row_in_item;
      bool cand = false;
      if (best >= thr_dist) {
        const int lim = __ldg(ratio_lim + min(best, kDotClamp));
        cand = (second <= lim);
      }
      midx[out] = -1;
      const unsigned mask = __ballot_sync(0xffffffffu, cand);
      if (mask) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(cand_count, (unsigned)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) {
          const unsigned pos = base + __popc(mask & ((1u << lane) - 1));
          if (pos < cand_capacity)
            cands[pos] = make_uint4(out, (uint32_t)bchunk, (uint32_t)best, (uint32_t)second);
        }
      }
      ++xi;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace ts

cudaError_t launch_match_top2_ts(const CUtensorMap& tmap, const uint8_t* pool, const MatchItem* items,
                                 const uint32_t* n_items_ptr, int thr_dist, const int* ratio_lim, int* midx,
                                 uint4* cands, unsigned int* cand_count, unsigned int cand_capacity, int grid,
                                 cudaStream_t stream) {
  const size_t smem = ts::kSmemTotal + 1024;
#ifdef B2_MATCH_EXPERIMENTS
  const char* ex = getenv("B2_MATCH_EXP");
  const int exp_mode = ex ? atoi(ex) : 0;
#endif
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, ts::kThreads, smem, stream>>>(tmap, pool, items, n_items_ptr, thr_dist, ratio_lim, midx, cands,
                                               cand_count, cand_capacity);
    return cudaSuccess;
  };
#ifdef B2_MATCH_EXPERIMENTS
  cudaError_t e = exp_mode == 1 ? go(ts::match_top2_ts_kernel<1>) : exp_mode == 2 ? go(ts::match_top2_ts_kernel<2>)
                                                                                    : go(ts::match_top2_ts_kernel<0>);
#else
  cudaError_t e = go(ts::match_top2_ts_kernel<0>);
#endif
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

}  // namespace b2
