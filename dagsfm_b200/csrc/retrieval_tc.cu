// Vocabulary-tree retrieval, stage 1: the EXACT k nearest visual words of every descriptor as a dense u8 x u8 -> s32
// contraction on tcgen05 with a fused top-k epilogue (VisualIndex::FindWordIds, src/retrieval/visual_index.h:701-744; the
// reference asks FLANN's approximate kd-forest, one descriptor at a time on the CPU).
//
//   |d - w|^2 = |d|^2 + |w|^2 - 2 d.w : the k nearest words of d are the k largest  s(w) = 2 d.w - |w|^2  (integer, exact:
//   d.w < 2^23, |w|^2 < 2^23), ties to the lower word id.
//
// Same machine as the descriptor matcher (match_tc_ts.cu): a work item is 256 descriptors (two M = 128 tiles), written
// once into TENSOR MEMORY as the A operand (TS mode: the stationary operand stays off the shared-memory port); the words
// stream through an 8-stage TMA ring as 128-row K-major SWIZZLE_128B blocks (the whole vocabulary -- 4 MB at 32 k words --
// is L2 resident and is read by every CTA); four K = 32-byte tcgen05.mma.kind::i8 per (tile, block) accumulate 128 x 128
// s32 in TMEM; eight epilogue warps (one TMEM lane = one descriptor per thread) pull a block's 128 dot products with one
// LDTM.x128, hand the accumulator back, and fold the block into the thread's running top-k: per four columns a few
// integer instructions (scores, their maximum, one vote against a register copy of the list's tail) decide warp-uniformly
// whether anything can enter a list (about one group in six), and only then the out-of-line insertion code runs; the
// block's |w|^2 reach the lanes as shared-memory broadcasts (staged per warp from one coalesced load).  Words are visited in ascending id and an insertion needs a strictly larger
// score, so equal scores keep the lower id first -- the oracle's order.
//
// TMEM (512 columns): accumulators 2 tiles x 128 columns (ping-pong between the tiles), A operand 2 buffers x 2 tiles x 32.
//
// Measured alternative (round 2, session 14): 16 epilogue warps, each thread folding one 64-column half of every block and
// the halves merged per item, were SLOWER (89 vs 52 ms per 2 M descriptors x 32 k words): two half-streams warm up two
// lists, so the insertion events -- the expensive part -- nearly double (14.5 G vs 10.9 G instructions).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "ptx.cuh"

namespace b2 {
namespace rtc {

constexpr int kDescBytes = 128;
constexpr int kSuperRows = 256;
constexpr int kTileRows = 128;
constexpr int kN = 128;                     // UMMA N = words per block
constexpr int kStagesY = 8;
constexpr int kNumEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kNumEpiWarps;     // 320
constexpr uint32_t kYBytes = kN * kDescBytes;        // 16 KiB per stage
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCol = 0;
constexpr uint32_t kACol = 256;
constexpr uint32_t kSmemY = 0;
constexpr uint32_t kSmemBar = kStagesY * kYBytes;
constexpr uint32_t kNumBars = 2 * kStagesY + 8;
constexpr uint32_t kSmemWsq = kSmemBar + kNumBars * 8 + 16;   // per epilogue warp: the |w|^2 of the current block (128 ints)
constexpr uint32_t kSmemTotal = kSmemWsq + kNumEpiWarps * kN * 4;
constexpr int kInvalidWord = 0x7fffffff;
constexpr int kNever = -0x7fffffff;         // score of a padding word (|w|^2 = INT_MAX, d.w = 0): never strictly above the initial lists

// The running k best of a descriptor (score descending; bw = word ids) live in LOCAL memory: only the rare insertion path
// touches them, the per-column path compares against a register copy of the list's tail.  One out-of-line copy of the
// insertion code serves all 32 column groups of a block (inlined per group it was 50-80 KB of SASS and the epilogue waited
// for instructions).  Scores are offered in ascending column order and must be strictly larger to enter, so equal scores
// keep the lower word id first.  Returns the new tail.
template <int K>
__device__ __noinline__ int offer4(int* bd, int* bw, int s0, int s1, int s2, int s3, int w) {
  int tail = bd[K - 1];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int sc = e == 0 ? s0 : e == 1 ? s1 : e == 2 ? s2 : s3;
    if (sc > tail) {
      int cd = sc, cw = w + e;
      bool placed = false;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int td = bd[k], tw = bw[k];
        if (placed || cd > td) {
          bd[k] = cd; bw[k] = cw;
          cd = td; cw = tw;
          placed = true;
        }
      }
      tail = bd[K - 1];
    }
  }
  return tail;
}

template <int K>
__global__ void __launch_bounds__(kThreads, 1)
word_knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_words, const uint8_t* __restrict__ desc, long long n_desc,
                   const int* __restrict__ word_sq, uint32_t n_blk, int32_t* __restrict__ out) {
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
  const uint32_t n_items = (uint32_t)((n_desc + kSuperRows - 1) / kSuperRows);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_words);
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
    // ===================================================== TMA producer: the word blocks, once per item
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        for (uint32_t b = 0; b < n_blk; ++b) {
          const uint32_t s = it % kStagesY;
          mbar_wait(y_empty(s), ((it / kStagesY) & 1) ^ 1);
          mbar_arrive_expect_tx(y_full(s), kYBytes);
          tma_load_2d(smem_base + kSmemY + s * kYBytes, &tmap_words, 0, (int)(b * kN), y_full(s));
          ++it;
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================= MMA issuer (TS mode, whole warp converged)
    constexpr uint32_t idesc = make_idesc_u8_s32(kTileRows, kN);
    uint32_t it = 0, xi = 0, tb = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t ab = xi & 1;
      mbar_wait(a_full(ab), (xi >> 1) & 1);
      tc_fence_after();
      for (uint32_t b = 0; b < n_blk; ++b) {
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
  } else {
    // ================================================= epilogue (+ A-operand loader)
    const int ew = warp - 2;
    const uint32_t quad = warp & 3;
    const uint32_t tile = ew >> 2;
    const uint32_t lane_addr = (quad * 32u) << 16;
    const uint32_t row_in_item = tile * 128 + quad * 32 + lane;
    int4* const wsq_sm = reinterpret_cast<int4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kSmemWsq) + ew * (kN / 4);
    auto load_a = [&](uint32_t item, uint32_t xi) {
      const uint32_t ab = xi & 1;
      mbar_wait(a_empty(ab), ((xi >> 1) & 1) ^ 1);
      tc_fence_after();
      const long long row = (long long)item * kSuperRows + row_in_item;
      uint32_t r[32];
      if (row < n_desc) {
        const uint4* src = reinterpret_cast<const uint4*>(desc + (size_t)row * kDescBytes);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v = __ldg(src + q);
          r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = 0u;
      }
      tmem_st_32x32b_x32_wait(tmem_base + lane_addr + kACol + (2 * ab + tile) * 32, r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(ab));
    };
    uint32_t tb = 0, xi = 0;
    if (blockIdx.x < n_items) load_a(blockIdx.x, 0);
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      if (item + gridDim.x < n_items) load_a(item + gridDim.x, xi + 1);
      int bd[K], bw[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { bd[k] = kNever; bw[k] = kInvalidWord; }
      int tail = kNever;
      for (uint32_t b = 0; b < n_blk; ++b) {
        // this block's |w|^2: one coalesced load per warp, issued before the wait so that its latency hides behind the MMAs
        const int4 wq_mine = __ldg(reinterpret_cast<const int4*>(word_sq + (size_t)b * kN) + lane);
        mbar_wait(t_full(tile), tb & 1);
        tc_fence_after();
        uint32_t v[128];
        tmem_ld_32x32b_x128_wait(tmem_base + lane_addr + kAccCol + tile * kN, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty(tile));
        wsq_sm[lane] = wq_mine;   // the previous block's reads of this buffer are complete (same warp, program order)
        __syncwarp();
        const int w0 = (int)(b * kN);
        int4 qn = wsq_sm[0];          // the same address in every lane: a shared-memory broadcast, fetched one group ahead
#pragma unroll
        for (int g = 0; g < 32; ++g) {
          const int4 q = qn;
          qn = wsq_sm[(g + 1) & 31];
          const int s0 = 2 * (int)v[4 * g] - q.x, s1 = 2 * (int)v[4 * g + 1] - q.y;
          const int s2 = 2 * (int)v[4 * g + 2] - q.z, s3 = 2 * (int)v[4 * g + 3] - q.w;
          const int m = max(__vimax3_s32(s0, s1, s2), s3);
          if (__builtin_expect(__any_sync(0xffffffffu, m > tail), 0)) {
            if (m > tail) tail = offer4<K>(bd, bw, s0, s1, s2, s3, w0 + 4 * g);
          }
        }
        __syncwarp();
        ++tb;
      }
      const long long row = (long long)item * kSuperRows + row_in_item;
      if (row < n_desc) {
#pragma unroll
        for (int k = 0; k < K; ++k) out[row * K + k] = bw[k];
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

}  // namespace rtc

// words: tensor map over the zero-padded word pool [n_blk * 128][128] (box 128 x 128, SWIZZLE_128B); word_sq [n_blk * 128]
// with INT_MAX at the padding rows; desc [n_desc][128] (16-byte aligned); out [n_desc * k].
cudaError_t launch_word_knn_tc(const CUtensorMap& tmap_words, const uint8_t* desc, long long n_desc, const int* word_sq,
                               uint32_t n_blk, int k, int32_t* out, int n_sm, cudaStream_t stream) {
  if (n_desc <= 0) return cudaSuccess;
  const size_t smem = rtc::kSmemTotal + 1024;
  const long long n_items = (n_desc + rtc::kSuperRows - 1) / rtc::kSuperRows;
  const int grid = (int)(n_items < n_sm ? n_items : n_sm);
  auto go = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, rtc::kThreads, smem, stream>>>(tmap_words, desc, n_desc, word_sq, n_blk, out);
    return cudaGetLastError();
  };
  switch (k) {
    case 1: return go(rtc::word_knn_tc_kernel<1>);
    case 2: return go(rtc::word_knn_tc_kernel<2>);
    case 3: return go(rtc::word_knn_tc_kernel<3>);
    case 4: return go(rtc::word_knn_tc_kernel<4>);
    case 5: return go(rtc::word_knn_tc_kernel<5>);
    case 6: return go(rtc::word_knn_tc_kernel<6>);
    case 7: return go(rtc::word_knn_tc_kernel<7>);
    default: return go(rtc::word_knn_tc_kernel<8>);
  }
}

}  // namespace b2
