// Descriptor matching, stage 1: the dense u8 x u8 -> s32 contraction on tcgen05
// tensor cores with a fused chunked top-2 epilogue.
//
// Replaces ComputeSiftDistanceMatrix + the row scan of FindBestMatchesOneWay
// (reference src/feature/sift.cc:76-162) and SiftGPU's MultiplyDescriptor_Kernel /
// RowMatch_Kernel (lib/SiftGPU/ProgramCU.cu:1408-1492,1692-1749).  The N1 x N2 int
// matrix is never written to memory: accumulators live in TMEM and are reduced on
// the fly.
//
// Work item = one 256-row supertile X of the "query" image against the whole
// "target" image Y (the reverse direction of the pair is a second set of items
// with the roles swapped -- "transpose by recomputation": the i8 tensor pipe has
// the headroom, the epilogue ALUs do not).
//
// Per item, per 128-row block of Y (K = 128 bytes is ONE smem stage):
//   TMA   : Y block  -> smem (128 rows x 128 B, SWIZZLE_128B)           warp 0
//   MMA   : 2 tiles x 4 x tcgen05.mma.kind::i8 (M128,N128,K32) -> TMEM  warp 1
//   EPI   : tcgen05.ld 128 lanes x 128 cols, per row: max of each 32-column
//           chunk (VIMNMX3 trees), running top-2 OVER CHUNK MAXIMA + best chunk id
//                                                                       warps 2-9
// At the end of the item each row knows  B  = its exact best dot,  C = the first
// 32-column chunk that attains it and  S' = the best maximum of any OTHER chunk.
// The exact second-best is max(S', second-best inside chunk C) and the exact best
// index lies in chunk C; both need only a 32-column rescan, which is done by the
// fix-up kernel and only for rows that can still pass the reference's distance and
// ratio tests evaluated in the integer domain (tables built on the host with the
// host's acosf, see match_api.cu).  Everything is integer-exact.
#include <cuda.h>
#include <cuda_runtime.h>

#include "match_common.cuh"
#include "ptx.cuh"

namespace b2 {

constexpr int kStagesY = 6;
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kNumEpiWarps;  // 320
constexpr uint32_t kTileBytes = kTileRows * kDescBytes;  // 16 KiB
constexpr uint32_t kXBytes = 2 * kTileBytes;             // 32 KiB
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSmemX = 0;
constexpr uint32_t kSmemY = 2 * kXBytes;
constexpr uint32_t kSmemBar = kSmemY + kStagesY * kTileBytes;
constexpr uint32_t kNumBars = 2 * kStagesY + 12;
constexpr uint32_t kSmemTotal = kSmemBar + kNumBars * 8 + 16;

__host__ size_t match_tc_smem_bytes() { return kSmemTotal + 1024; }

// max over 32 registers with 3-input integer max (VIMNMX3): 16 ALU ops.
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

// PROF: per-role cycle counters (B2_MATCH_PROFILE=1); the counters cost ~15 %, so the
// production instantiation has none.
template <bool PROF>
__global__ void __launch_bounds__(kThreads, 1)
match_top2_kernel(const __grid_constant__ CUtensorMap tmap, const MatchItem* __restrict__ items,
                  const uint32_t* __restrict__ n_items_ptr, int thr_dist,
                  const int* __restrict__ ratio_lim, int* __restrict__ midx,
                  uint4* __restrict__ cands, unsigned int* __restrict__ cand_count,
                  unsigned int cand_capacity, unsigned long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kSmemBar;
  auto y_full = [&](uint32_t s) { return bar_base + 8 * s; };
  auto y_empty = [&](uint32_t s) { return bar_base + 8 * (kStagesY + s); };
  auto x_full = [&](uint32_t s) { return bar_base + 8 * (2 * kStagesY + s); };
  auto x_empty = [&](uint32_t s) { return bar_base + 8 * (2 * kStagesY + 2 + s); };
  // accumulator barriers are per (TMEM buffer, X tile): the two epilogue warps that share an
  // SMSP (tile 0 / tile 1 of the same lane quadrant) run half a block out of phase, so one
  // drains TMEM while the other occupies the ALU pipe
  auto t_full = [&](uint32_t buf, uint32_t t) { return bar_base + 8 * (2 * kStagesY + 4 + 2 * buf + t); };
  auto t_empty = [&](uint32_t buf, uint32_t t) { return bar_base + 8 * (2 * kStagesY + 8 + 2 * buf + t); };
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
      mbar_init(x_full(s), 1);
      mbar_init(x_empty(s), 1);
      for (uint32_t t = 0; t < 2; ++t) {
        mbar_init(t_full(s, t), 1);
        mbar_init(t_empty(s, t), kNumEpiWarps / 2);
      }
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
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t it = 0, xi = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const MatchItem w = items[item];
        const uint32_t xs = xi & 1;
        mbar_wait(x_empty(xs), ((xi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(x_full(xs), kXBytes);
        tma_load_2d(smem_base + kSmemX + xs * kXBytes, &tmap, 0, (int)w.x_row, x_full(xs));
        tma_load_2d(smem_base + kSmemX + xs * kXBytes + kTileBytes, &tmap, 0,
                    (int)(w.x_row + kTileRows), x_full(xs));
        for (uint32_t b = 0; b < w.y_nblk; ++b) {
          const uint32_t s = it % kStagesY;
          mbar_wait(y_empty(s), ((it / kStagesY) & 1) ^ 1);
          mbar_arrive_expect_tx(y_full(s), kTileBytes);
          tma_load_2d(smem_base + kSmemY + s * kTileBytes, &tmap, 0,
                      (int)(w.y_row + b * kTileRows), y_full(s));
          ++it;
        }
        ++xi;
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer (whole warp converged)
    {
      constexpr uint32_t idesc = make_idesc_u8_s32(kTileRows, kTileRows);
      uint32_t it = 0, xi = 0, tb = 0;
      long long w_y = 0, w_t = 0, t_begin = (PROF ? clock64() : 0ll);
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t y_nblk = items[item].y_nblk;
        const uint32_t xs = xi & 1;
        mbar_wait(x_full(xs), (xi >> 1) & 1);
        const uint32_t xa = smem_base + kSmemX + xs * kXBytes;
        // The single issuing thread must never sit in a barrier round trip (~100 cycles) with the
        // tensor pipe's short queue drained: the barriers of block b+1 are PROBED (test_wait) before
        // the MMAs of block b are issued and only waited on if the probe failed.  This needs the
        // double-buffered accumulators: t_empty of block b+1 belongs to block b-1's buffer.
        uint32_t ok_y = 0, ok_t0 = 0, ok_t1 = 0;  // probes of the current block (0 = unknown)
        for (uint32_t b = 0; b < y_nblk; ++b) {
          const uint32_t s = it % kStagesY;
          const uint32_t buf = tb & 1;
          long long c0 = (PROF ? clock64() : 0ll);
          if (!ok_y) mbar_wait(y_full(s), (it / kStagesY) & 1);
          if (PROF) w_y += clock64() - c0;
          const uint32_t ya = smem_base + kSmemY + s * kTileBytes;
          // probes for the next block of this item
          uint32_t nk_y = 0, nk_t0 = 0, nk_t1 = 0;
          if (b + 1 < y_nblk) {
            const uint32_t itn = it + 1, tbn = tb + 1;
            nk_y = mbar_test(y_full(itn % kStagesY), (itn / kStagesY) & 1);
            nk_t0 = mbar_test(t_empty(tbn & 1, 0), ((tbn >> 1) & 1) ^ 1);
            nk_t1 = mbar_test(t_empty(tbn & 1, 1), ((tbn >> 1) & 1) ^ 1);
          }
#pragma unroll
          for (uint32_t t = 0; t < 2; ++t) {
            c0 = (PROF ? clock64() : 0ll);
            if (!(t == 0 ? ok_t0 : ok_t1)) mbar_wait(t_empty(buf, t), ((tb >> 1) & 1) ^ 1);
            if (PROF) w_t += clock64() - c0;
            tc_fence_after();
            const uint32_t d = tmem_base + buf * 256 + t * 128;
            if (elect_one()) {
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k) {
                const uint64_t ad = make_kmajor_sw128_desc(xa + t * kTileBytes + k * 32);
                const uint64_t bd = make_kmajor_sw128_desc(ya + k * 32);
                mma_i8_ss(d, ad, bd, idesc, k);
              }
              tc_commit(t_full(buf, t));  // this tile's accumulator is ready for its epilogue warps
              if (t == 1) tc_commit(y_empty(s));  // smem stage reusable once these MMAs retire
            }
            __syncwarp();
          }
          ok_y = nk_y;
          ok_t0 = nk_t0;
          ok_t1 = nk_t1;
          ++it;
          ++tb;
        }
        if (elect_one()) tc_commit(x_empty(xs));
        __syncwarp();
        ++xi;
      }
      if (PROF && prof && lane == 0) {
        atomicAdd(prof + 0, (unsigned long long)w_y);
        atomicAdd(prof + 1, (unsigned long long)w_t);
        atomicAdd(prof + 2, (unsigned long long)((PROF ? clock64() : 0ll) - t_begin));
      }
    }
  } else {
    // ========================================================= epilogue
    const int ew = warp - 2;           // 0..7
    const uint32_t quad = warp & 3;    // TMEM lane quadrant this warp may read
    const uint32_t tile = ew >> 2;     // which of the two X tiles (M=128 accumulators)
    const uint32_t lane_addr = (quad * 32u) << 16;
    const uint32_t row_in_item = tile * 128 + quad * 32 + lane;
    uint32_t tb = 0;
    long long e_wait = 0, e_ld = 0, e_alu = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t y_nblk = items[item].y_nblk;
      int best = 0, second = 0, bchunk = 0;
      for (uint32_t b = 0; b < y_nblk; ++b) {
        const uint32_t buf = tb & 1;
        const long long c0 = (PROF ? clock64() : 0ll);
        mbar_wait(t_full(buf, tile), (tb >> 1) & 1);
        const long long c1 = (PROF ? clock64() : 0ll);
        tc_fence_after();
        uint32_t v[128];
        const uint32_t ta = tmem_base + lane_addr + buf * 256 + tile * 128;
        tmem_ld_32x32b_x128_wait(ta, v);
        // values are in registers: hand the TMEM buffer back before the ALU work
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(buf, tile));
        const long long c2 = (PROF ? clock64() : 0ll);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int m = max32(v + 32 * c);
          const int chunk = (int)(b * 4 + c);
          // sequential top-2 over chunk maxima; strict '>' keeps the FIRST chunk on ties
          second = max(second, min(best, m));
          bchunk = (m > best) ? chunk : bchunk;
          best = max(best, m);
        }
        e_wait += c1 - c0;
        e_ld += c2 - c1;
        e_alu += (PROF ? clock64() : 0ll) - c2;
        ++tb;
      }
      // distance + ratio tests in the integer domain (monotone tables, see match_api.cu)
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
    }
    if (PROF && prof && lane == 0) {
      atomicAdd(prof + 3, (unsigned long long)e_wait);
      atomicAdd(prof + 4, (unsigned long long)e_ld);
      atomicAdd(prof + 5, (unsigned long long)e_alu);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

cudaError_t launch_match_top2(const CUtensorMap& tmap, const MatchItem* items,
                              const uint32_t* n_items_ptr, int thr_dist, const int* ratio_lim,
                              int* midx, uint4* cands, unsigned int* cand_count,
                              unsigned int cand_capacity, int grid, cudaStream_t stream,
                              unsigned long long* prof) {
  const size_t smem = match_tc_smem_bytes();
  if (prof) {
    cudaError_t e = cudaFuncSetAttribute(match_top2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    match_top2_kernel<true><<<grid, kThreads, smem, stream>>>(tmap, items, n_items_ptr, thr_dist, ratio_lim, midx, cands,
                                                               cand_count, cand_capacity, prof);
  } else {
    cudaError_t e = cudaFuncSetAttribute(match_top2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    match_top2_kernel<false><<<grid, kThreads, smem, stream>>>(tmap, items, n_items_ptr, thr_dist, ratio_lim, midx, cands,
                                                                cand_count, cand_capacity, nullptr);
  }
  return cudaGetLastError();
}

}  // namespace b2
