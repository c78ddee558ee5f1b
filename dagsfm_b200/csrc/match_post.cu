// Descriptor matching, stages 0 and 2-4: work-item construction, the exact
// in-chunk rescan ("fix-up"), the cross check and the ordered compaction into
// FeatureMatches.  All integer work, HBM/L2-bound, tiny next to the contraction.
//
// Reference semantics restated here:
//   * FindBestMatchesOneWay (src/feature/sift.cc:111-162): best = first maximal
//     column, second = largest value among all OTHER columns, tests in float on
//     acos() of the normalised dots -- evaluated here through host-built integer
//     tables (thr_dist, ratio_lim[]) that are exactly equivalent.
//   * FindBestMatches (sift.cc:164-198): cross check m21[m12[i]] == i, output in
//     ascending idx1.
#include <cuda_runtime.h>

#include "match_common.cuh"

namespace b2 {

// ---------------------------------------------------------------- stage 0
// One thread per pair: how many work items does it need?
__global__ void pair_items_kernel(const uint32_t* __restrict__ pairs, int64_t n_pairs,
                                  const int32_t* __restrict__ img_n, int32_t n_images,
                                  uint32_t* __restrict__ n_items_of_pair, int* __restrict__ err) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
  if (i1 >= (uint32_t)n_images || i2 >= (uint32_t)n_images) {
    atomicExch(err, 1);
    n_items_of_pair[p] = 0;
    return;
  }
  const uint32_t n1 = img_n[i1], n2 = img_n[i2];
  uint32_t n = 0;
  if (n1 > 0 && n2 > 0) n = (n1 + kSuperRows - 1) / kSuperRows + (n2 + kSuperRows - 1) / kSuperRows;
  n_items_of_pair[p] = n;
}

// Single-block exclusive scan (n <= a few 100k): out[i] = carry_in + sum_{j<i} in[j];
// *total_out = carry_in + sum of all.  T = uint32_t or int64_t accumulators.
template <typename TIn, typename TOut>
__global__ void block_scan_kernel(const TIn* __restrict__ in, int64_t n, TOut* __restrict__ out,
                                  const TOut* carry_in, TOut* total_out,
                                  bool write_last) {
  __shared__ TOut warp_sums[32];
  __shared__ TOut running;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) running = carry_in ? *carry_in : (TOut)0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + tid;
    const TOut v = (i < n) ? (TOut)in[i] : (TOut)0;
    TOut x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const TOut y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
      TOut s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : (TOut)0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const TOut y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    const TOut warp_off = (w == 0) ? (TOut)0 : warp_sums[w - 1];
    const TOut r = running;
    if (i < n) out[i] = r + warp_off + x - v;
    __syncthreads();
    if (tid == 0) running = r + warp_sums[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  if (tid == 0) {
    if (write_last) out[n] = running;
    if (total_out) *total_out = running;
  }
}

// One thread per pair: write its items and its PairMeta.
__global__ void fill_items_kernel(const uint32_t* __restrict__ pairs, int64_t n_pairs,
                                  const int32_t* __restrict__ img_n, int32_t n_images,
                                  const uint32_t* __restrict__ img_row,
                                  const uint32_t* __restrict__ item_start,
                                  MatchItem* __restrict__ items, PairMeta* __restrict__ meta,
                                  uint32_t y_block_rows) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const uint32_t i1 = pairs[2 * p], i2 = pairs[2 * p + 1];
  // a pair that names an image outside the store was given no items (pair_items_kernel raised the error
  // flag): it must not index img_n / img_row either
  const bool valid = i1 < (uint32_t)n_images && i2 < (uint32_t)n_images;
  const uint32_t n1 = valid ? img_n[i1] : 0u, n2 = valid ? img_n[i2] : 0u;
  PairMeta pm;
  pm.item_start = item_start[p];
  pm.n1 = n1;
  pm.n2 = n2;
  pm.nt1 = 0;
  if (n1 > 0 && n2 > 0) {
    const uint32_t nt1 = (n1 + kSuperRows - 1) / kSuperRows, nt2 = (n2 + kSuperRows - 1) / kSuperRows;
    pm.nt1 = nt1;
    const uint32_t r1 = img_row[i1], r2 = img_row[i2];
    // Y is streamed in blocks of y_block_rows (128 for the SS kernel, 96 for the TS kernel); the
    // pool pads every image so that the last block never reaches the next image.
    const uint32_t yb2 = (n2 + y_block_rows - 1) / y_block_rows, yb1 = (n1 + y_block_rows - 1) / y_block_rows;
    MatchItem* it = items + pm.item_start;
    for (uint32_t t = 0; t < nt1; ++t) it[t] = MatchItem{r1 + t * kSuperRows, r2, yb2, 0u};
    for (uint32_t t = 0; t < nt2; ++t) it[nt1 + t] = MatchItem{r2 + t * kSuperRows, r1, yb1, 0u};
  }
  meta[p] = pm;
}

// ---------------------------------------------------------------- stage 2
// One warp per candidate row: recompute the 32 dots of the row against its best
// chunk (lane l <-> Y row 32*C + l) with dp4a, then
//   idx    = first lane attaining `best`      (strict '>' tie rule, sift.cc:126)
//   second = max(S', max over the other 31 lanes)
//   accept iff second <= ratio_lim[best]
__global__ void __launch_bounds__(256)
match_fixup_kernel(const uint8_t* __restrict__ pool, const MatchItem* __restrict__ items,
                   const uint4* __restrict__ cands, const unsigned int* __restrict__ cand_count,
                   unsigned int cand_capacity, const int* __restrict__ ratio_lim,
                   int* __restrict__ midx, int* __restrict__ err) {
  const unsigned n = min(*cand_count, cand_capacity);
  const int lane = threadIdx.x & 31;
  const unsigned warps_total = (gridDim.x * blockDim.x) >> 5;
  for (unsigned c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n; c += warps_total) {
    const uint4 cd = cands[c];
    const uint32_t out = cd.x, chunk = cd.y;
    const int best = (int)cd.z, s_outer = (int)cd.w;
    const MatchItem w = items[out / kSuperRows];
    // The chunk's 32 rows are read COALESCED: in step t the warp loads rows 4t .. 4t+3 as one 512-byte span (lane l: row
    // 4t + l/8, 16-byte piece l%8), each lane multiplies its piece with the matching piece of the query row, and three
    // butterfly steps add the eight pieces of a row.  (One row per lane, 8 loads of 16 bytes at a 128-byte stride, touched
    // 32 cache lines per load instruction and fetched every 32-byte sector twice.)  Integer sums: the order does not matter.
    const int piece = lane & 7;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(pool + (size_t)(w.x_row + out % kSuperRows) * kDescBytes) + piece);
    const uint4* yc = reinterpret_cast<const uint4*>(pool + (size_t)(w.y_row + chunk * kChunk) * kDescBytes) + lane;
    int v = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const uint4 b = __ldg(yc + t * 32);   // rows 4t .. 4t+3: 32 lanes x 16 bytes, contiguous
      unsigned acc = __dp4a(a.x, b.x, 0u);
      acc = __dp4a(a.y, b.y, acc);
      acc = __dp4a(a.z, b.z, acc);
      acc = __dp4a(a.w, b.w, acc);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);          // the eight lanes of group g = lane / 8 hold the dot of row 4t + g
      const unsigned mine = __shfl_sync(0xffffffffu, acc, 8 * (lane & 3));   // lane l wants row l = 4 (l/4) + l%4
      if ((lane >> 2) == t) v = (int)mine;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, v == best);
    if (hit == 0) {  // cannot happen: the tensor-core pass found `best` in this chunk
      if (lane == 0) atomicExch(err, 2);
      continue;
    }
    const int first = __ffs(hit) - 1;
    int other = (lane == first) ? 0 : v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) other = max(other, __shfl_xor_sync(0xffffffffu, other, o));
    const int second = max(s_outer, other);
    if (lane == 0) {
      const int lim = ratio_lim[min(best, kDotClamp)];
      if (second <= lim) midx[out] = (int)(chunk * kChunk + first);
    }
  }
}

// ------------------------------------------------------------- stages 3-4
// One warp per pair.  WRITE == false: count the surviving matches;
// WRITE == true: write them at offsets[p] in ascending idx1.
template <bool WRITE>
__global__ void __launch_bounds__(256)
match_cross_kernel(const PairMeta* __restrict__ meta, int64_t n_pairs, const int* __restrict__ midx,
                   int cross_check, uint32_t* __restrict__ counts,
                   const int64_t* __restrict__ offsets, uint32_t* __restrict__ out_matches,
                   int64_t capacity) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (p >= n_pairs) return;
  const PairMeta pm = meta[p];
  uint32_t total = 0;
  if (pm.n1 > 0 && pm.n2 > 0) {
    const int* m12 = midx + (size_t)pm.item_start * kSuperRows;
    const int* m21 = midx + (size_t)(pm.item_start + pm.nt1) * kSuperRows;
    const int64_t off = WRITE ? offsets[p] : 0;
    for (uint32_t i0 = 0; i0 < pm.n1; i0 += 32) {
      const uint32_t i = i0 + lane;
      int j = -1;
      if (i < pm.n1) j = m12[i];
      bool ok = j >= 0;
      if (ok && cross_check) ok = (m21[j] == (int)i);
      const unsigned mask = __ballot_sync(0xffffffffu, ok);
      if (WRITE && ok) {
        const int64_t pos = off + total + __popc(mask & ((1u << lane) - 1));
        if (pos < capacity) {
          out_matches[2 * pos] = i;
          out_matches[2 * pos + 1] = (uint32_t)j;
        }
      }
      total += __popc(mask);
    }
  }
  if (!WRITE && lane == 0) counts[p] = total;
}

// ---------------------------------------------------------------- launchers
cudaError_t launch_pair_items(const uint32_t* pairs, int64_t n_pairs, const int32_t* img_n,
                              int32_t n_images, uint32_t* n_items_of_pair, int* err,
                              cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  pair_items_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(pairs, n_pairs, img_n,
                                                                      n_images, n_items_of_pair, err);
  return cudaGetLastError();
}
cudaError_t launch_scan_u32(const uint32_t* in, int64_t n, uint32_t* out, uint32_t* total,
                            cudaStream_t s) {
  block_scan_kernel<uint32_t, uint32_t><<<1, 1024, 0, s>>>(in, n, out, nullptr, total, false);
  return cudaGetLastError();
}
cudaError_t launch_scan_counts(const uint32_t* counts, int64_t n, int64_t* offsets,
                               int64_t* carry_inout, bool write_last, cudaStream_t s) {
  block_scan_kernel<uint32_t, int64_t><<<1, 1024, 0, s>>>(counts, n, offsets, carry_inout,
                                                          carry_inout, write_last);
  return cudaGetLastError();
}
cudaError_t launch_fill_items(const uint32_t* pairs, int64_t n_pairs, const int32_t* img_n, int32_t n_images,
                              const uint32_t* img_row, const uint32_t* item_start, MatchItem* items,
                              PairMeta* meta, uint32_t y_block_rows, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  fill_items_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, s>>>(pairs, n_pairs, img_n, n_images, img_row,
                                                                      item_start, items, meta, y_block_rows);
  return cudaGetLastError();
}
cudaError_t launch_fixup(const uint8_t* pool, const MatchItem* items, const uint4* cands,
                         const unsigned int* cand_count, unsigned int cand_capacity,
                         const int* ratio_lim, int* midx, int* err, int n_sm, cudaStream_t s) {
  match_fixup_kernel<<<n_sm * 8, 256, 0, s>>>(pool, items, cands, cand_count, cand_capacity,
                                              ratio_lim, midx, err);
  return cudaGetLastError();
}
cudaError_t launch_cross_count(const PairMeta* meta, int64_t n_pairs, const int* midx,
                               int cross_check, uint32_t* counts, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  match_cross_kernel<false><<<(unsigned)((n_pairs * 32 + 255) / 256), 256, 0, s>>>(
      meta, n_pairs, midx, cross_check, counts, nullptr, nullptr, 0);
  return cudaGetLastError();
}
cudaError_t launch_cross_write(const PairMeta* meta, int64_t n_pairs, const int* midx,
                               int cross_check, const int64_t* offsets, uint32_t* out_matches,
                               int64_t capacity, cudaStream_t s) {
  if (n_pairs == 0) return cudaSuccess;
  match_cross_kernel<true><<<(unsigned)((n_pairs * 32 + 255) / 256), 256, 0, s>>>(
      meta, n_pairs, midx, cross_check, nullptr, offsets, out_matches, capacity);
  return cudaGetLastError();
}

}  // namespace b2
