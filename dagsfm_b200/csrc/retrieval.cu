// Vocabulary-tree image retrieval on the GPU: the stage that turns a set of images into the candidate pair list the
// matcher and the verifier then work through (the input stage of BASELINE configs[2]).
//
// Reference: VocabSimilarityGraph::Run (src/graph/similarity_graph.cpp:101-200) over retrieval::VisualIndex
//   VisualIndex::Add / Prepare / Query / QueryAndFindWordIds   src/retrieval/visual_index.h:207-262, 505-540, 668-698
//   InvertedIndex::AddEntry / Finalize / Query                  src/retrieval/inverted_index.h:229-283, 162-171
//   InvertedFile::ScoreFeature (burstiness + idf^2)             src/retrieval/inverted_file.h:305-366
//   HammingDistWeightFunctor<64, 16>                            src/retrieval/utils.h:52-82
// The reference runs FLANN's approximate kd-forest per descriptor on the CPU and scores one query image per thread.
// Here:
//   word_knn_tc_kernel  (retrieval_tc.cu) EXACT k nearest visual words of every descriptor: descriptor x word contraction on
//                       tcgen05 (u8 x u8 -> s32, descriptors in TMEM, words streamed by TMA) with a register top-k epilogue on
//                       2 d.w - |w|^2; ties -> lower word id.  word_knn_kernel (dp4a, below) is the independent SIMT
//                       statement of the same result, kept as the test seam b2_retrieval_debug_word_ids_simt.
//   project_kernel      the 64 Hamming-embedding projections of every descriptor (once per descriptor, shared by its words)
//   index_sig_kernel    signature of every indexed descriptor for its nearest word (thresholds, bits assembled with two ballots).
//   word histogram -> scan -> scatter -> per-word rank sort by (image, feature): the inverted files as one CSR array,
//                       deterministic whatever the atomics' order; idf weights, per-image normalisation constants.
//   query_kernel        one CTA per query image: every (feature, neighbouring word) walks that word's inverted file,
//                       votes with the Hamming weight, applies the per-(feature, image) burstiness normalisation and
//                       accumulates into a per-image score array in shared memory; normalisation and top-k in place.
// Differences from the reference (both stated in oracle/retrieval_oracle.cc as well): the word search is exact where
// FLANN is approximate; the projection is summed left to right in float where Eigen picks its own order.  Scores are
// float sums whose order differs from the reference's (atomics): parity is to 1e-5 relative, not bitwise.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dagsfm_b200.h"
#include "common_host.h"

namespace b2 {
cudaError_t launch_scan_u32(const uint32_t* in, int64_t n, uint32_t* out, uint32_t* total, cudaStream_t s);
cudaError_t launch_word_knn_tc(const CUtensorMap& tmap_words, const uint8_t* desc, long long n_desc, const int* word_sq,
                               uint32_t n_blk, int k, int32_t* out, int n_sm, cudaStream_t stream);   // retrieval_tc.cu

namespace rt {

constexpr int kDim = 128, kEmb = 64, kMaxHamming = 24, kMaxK = 8;
constexpr int kInvalidWord = 0x7fffffff;

struct __align__(16) Entry { int32_t image, feature; unsigned long long bits; };

// ------------------------------------------------------------------ exact k nearest words
constexpr int kKnnThreads = 128, kWordTile = 64;
template <int K>
__global__ void __launch_bounds__(kKnnThreads) word_knn_kernel(const uint8_t* __restrict__ desc, int64_t n, const uint8_t* __restrict__ words,
                                                               const int32_t* __restrict__ word_sq, int n_words, int32_t* __restrict__ out) {
  __shared__ uint4 tile[kWordTile * (kDim / 16)];
  __shared__ int tile_sq[kWordTile];
  const int64_t q = blockIdx.x * (int64_t)kKnnThreads + threadIdx.x;
  uint32_t d[kDim / 4];
  int dsq = 0;
  if (q < n) {
    const uint4* p = reinterpret_cast<const uint4*>(desc + q * kDim);
#pragma unroll
    for (int j = 0; j < kDim / 16; ++j) {
      const uint4 v = p[j];
      d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < kDim / 4; ++j) dsq = (int)__dp4a(d[j], d[j], (unsigned)dsq);
  } else {
#pragma unroll
    for (int j = 0; j < kDim / 4; ++j) d[j] = 0;
  }
  int bd[K], bw[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { bd[k] = 0x7fffffff; bw[k] = kInvalidWord; }
  for (int w0 = 0; w0 < n_words; w0 += kWordTile) {
    const int nw = min(kWordTile, n_words - w0);
    __syncthreads();
    for (int e = threadIdx.x; e < nw * (kDim / 16); e += kKnnThreads) tile[e] = reinterpret_cast<const uint4*>(words + (size_t)w0 * kDim)[e];
    for (int e = threadIdx.x; e < nw; e += kKnnThreads) tile_sq[e] = word_sq[w0 + e];
    __syncthreads();
    for (int w = 0; w < nw; ++w) {
      unsigned dot = 0;
#pragma unroll
      for (int j = 0; j < kDim / 16; ++j) {
        const uint4 v = tile[w * (kDim / 16) + j];
        dot = __dp4a(d[4 * j], v.x, dot);
        dot = __dp4a(d[4 * j + 1], v.y, dot);
        dot = __dp4a(d[4 * j + 2], v.z, dot);
        dot = __dp4a(d[4 * j + 3], v.w, dot);
      }
      const int dist = dsq + tile_sq[w] - 2 * (int)dot;  // exact: every term < 2^24
      if (dist < bd[K - 1]) {  // strict: an equal distance keeps the earlier (lower) word id
        int cd = dist, cw = w0 + w;
        bool placed = false;  // once the new word has its slot, everything behind it moves down one (equal distances keep their order)
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if (placed || cd < bd[k]) {
            const int td = bd[k], tw = bw[k];
            bd[k] = cd; bw[k] = cw;
            cd = td; cw = tw;
            placed = true;
          }
        }
      }
    }
  }
  if (q < n) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[q * K + k] = bw[k];
  }
}

__global__ void word_norms_kernel(const uint8_t* __restrict__ words, int n_words, int n_pad, int32_t* __restrict__ sq) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_pad) return;
  if (w >= n_words) { sq[w] = 0x7fffffff; return; }  // padding rows of the tensor-core blocks: never a nearest word
  int s = 0;
  for (int j = 0; j < kDim; ++j) { const int v = words[(size_t)w * kDim + j]; s += v * v; }
  sq[w] = s;
}

// ------------------------------------------------------------------ Hamming embedding
// project_kernel: P[f][i] = sum_j proj[i][j] * d_f[j] for the descriptors of a batch of images, one warp per descriptor,
// lane l computes rows l and l + 32, summing left to right in float (mul, then add: this file is compiled with
// --fmad=false) -- the projection of a descriptor does not depend on the visual word, only the thresholds do, so it is
// computed once per descriptor and shared by its num_neighbors signatures.  projT: [128][64] (transposed) in shared memory.
constexpr int kProjWarps = 8;
__global__ void __launch_bounds__(32 * kProjWarps) project_kernel(const uint8_t* __restrict__ desc, int64_t f0, int64_t f1, const float* __restrict__ proj,
                                                                  float* __restrict__ out) {
  __shared__ float projT[kDim * kEmb];
  for (int e = threadIdx.x; e < kDim * kEmb; e += blockDim.x) projT[(e % kDim) * kEmb + e / kDim] = proj[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int64_t f = f0 + blockIdx.x * (int64_t)kProjWarps + (threadIdx.x >> 5); f < f1; f += (int64_t)gridDim.x * kProjWarps) {
    const uint32_t* d4 = reinterpret_cast<const uint32_t*>(desc + f * kDim);
    float s0 = 0.0f, s1 = 0.0f;
    for (int j4 = 0; j4 < kDim / 4; ++j4) {
      const uint32_t v = d4[j4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float x = (float)((v >> (8 * b)) & 0xffu);
        const int j = 4 * j4 + b;
        s0 += projT[j * kEmb + lane] * x;
        s1 += projT[j * kEmb + lane + 32] * x;
      }
    }
    out[(f - f0) * kEmb + lane] = s0;
    out[(f - f0) * kEmb + lane + 32] = s1;
  }
}
__device__ __forceinline__ unsigned long long signature_bits(float lo, float hi, const float* __restrict__ thr, int lane) {
  const unsigned b0 = __ballot_sync(0xffffffffu, lo > thr[lane]);
  const unsigned b1 = __ballot_sync(0xffffffffu, hi > thr[lane + 32]);
  return (unsigned long long)b0 | ((unsigned long long)b1 << 32);
}

// Signature of every indexed descriptor of the batch for its nearest word + the per-word entry counts (one warp per descriptor).
constexpr int kSigWarps = 8;
__global__ void __launch_bounds__(32 * kSigWarps) index_sig_kernel(const float* __restrict__ P, int64_t f0, int64_t f1, const int32_t* __restrict__ nn, int nn_stride,
                                                                   const float* __restrict__ thr, const int32_t* __restrict__ feat_image,
                                                                   const int32_t* __restrict__ feat_index, uint32_t* __restrict__ word_count,
                                                                   Entry* __restrict__ staged) {
  const int lane = threadIdx.x & 31;
  for (int64_t f = f0 + blockIdx.x * (int64_t)kSigWarps + (threadIdx.x >> 5); f < f1; f += (int64_t)gridDim.x * kSigWarps) {
    const int w = nn[f * nn_stride];
    const float lo = P[(f - f0) * kEmb + lane], hi = P[(f - f0) * kEmb + lane + 32];
    const unsigned long long bits = (w == kInvalidWord) ? 0ull : signature_bits(lo, hi, thr + (size_t)w * kEmb, lane);
    if (lane == 0) {
      Entry e;
      e.image = (w == kInvalidWord) ? -1 : feat_image[f];
      e.feature = feat_index[f];
      e.bits = bits;
      staged[f] = e;
      if (w != kInvalidWord) atomicAdd(word_count + w, 1u);
    }
  }
}

__global__ void scatter_kernel(const Entry* __restrict__ staged, const int32_t* __restrict__ nn, int nn_stride, int64_t n,
                               const uint32_t* __restrict__ word_start, uint32_t* __restrict__ cursor, Entry* __restrict__ out) {
  const int64_t f = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int w = nn[f * nn_stride];
  if (w == kInvalidWord) return;
  out[word_start[w] + atomicAdd(cursor + w, 1u)] = staged[f];
}

// One warp per word: rank sort of its entries by (image, feature) -- keys are unique -- into `sorted`; distinct images -> idf.
__global__ void __launch_bounds__(256) sort_words_kernel(const Entry* __restrict__ in, Entry* __restrict__ sorted, const uint32_t* __restrict__ word_start,
                                                         int n_words, int n_images_total, float* __restrict__ idf) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_words) return;
  const uint32_t s = word_start[w], e = word_start[w + 1];
  int distinct = 0;
  for (uint32_t a = s + lane; a < e; a += 32) {
    const Entry x = in[a];
    const long long key = ((long long)x.image << 32) | (unsigned)x.feature;
    uint32_t rank = 0;
    bool first_of_image = true;
    for (uint32_t b = s; b < e; ++b) {
      const Entry y = in[b];
      const long long kb = ((long long)y.image << 32) | (unsigned)y.feature;
      rank += kb < key ? 1u : 0u;
      if (y.image == x.image && kb < key) first_of_image = false;
    }
    sorted[s + rank] = x;
    distinct += first_of_image ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) distinct += __shfl_xor_sync(0xffffffffu, distinct, o);
  if (lane == 0) idf[w] = (e > s) ? (float)log((double)n_images_total / (double)distinct) : 0.0f;
}

// self-similarity of every indexed image: sum over its entries of idf(word)^2, in double; norm = 1 / sqrt
__global__ void __launch_bounds__(256) image_norm_kernel(const int32_t* __restrict__ nn, int nn_stride, const int64_t* __restrict__ img_off, int n_images,
                                                         const float* __restrict__ idf, float* __restrict__ norm) {
  __shared__ double part[256];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int64_t f = img_off[i] + threadIdx.x; f < img_off[i + 1]; f += blockDim.x) {
    const int w = nn[f * nn_stride];
    if (w != kInvalidWord) s += (double)(idf[w] * idf[w]);
  }
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[i] = part[0] > 0.0 ? (float)(1.0 / sqrt(part[0])) : 0.0f;
}

// ------------------------------------------------------------------ query
struct QueryArgs {
  const float* P;             // projections of the batch's descriptors [(f - f_base) * 64]
  int64_t f_base;
  const int64_t* q_off;       // [n_images + 1] descriptor offsets of all images
  int q0, q1;                 // query images of this launch
  const int32_t* nn;          // [n_desc * K] nearest words of the query descriptors
  int K;
  const float* thr;
  const uint8_t* has_emb;
  const float* idf;
  const uint32_t* word_start;
  const Entry* entries;
  const float* norm;          // [n_index_images]
  const float* lut;           // [65]
  int n_index_images;
  int max_num_images;
  int32_t* out_ids;           // [n_query * max_num_images]
  float* out_scores;
  int32_t* out_count;
  float* score_scratch;       // [gridDim.x * n_index_images] when the score array does not fit shared memory, else null
};

// One CTA per query image.  A warp takes a feature and, for each of its K words, streams the word's inverted file 32
// entries at a time (one coalesced 512-byte load); lane l weighs entry l.  Entries are sorted by image, so the entries of
// one image form a run: the lane at the head of a run sums it IN ENTRY ORDER (shuffles inside the chunk, plain loads if it
// runs past the chunk's end -- the lanes of the next chunk that continue it are skipped), applies the burstiness
// normalisation and adds the vote to the image's score in shared memory.
constexpr int kQueryThreads = 256;
__global__ void __launch_bounds__(kQueryThreads) query_kernel(QueryArgs A, int scores_in_smem) {
  extern __shared__ float q_smem[];  // scores [n_index_images] (when they fit) | hit flags
  __shared__ double s_self[kQueryThreads / 32];
  __shared__ float s_best[kQueryThreads];
  __shared__ int s_besti[kQueryThreads];
  __shared__ float s_lut[kEmb + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = A.n_index_images;
  float* scores = scores_in_smem ? q_smem : A.score_scratch + (size_t)blockIdx.x * N;
  uint8_t* hit = reinterpret_cast<uint8_t*>(q_smem + (scores_in_smem ? N : 0));
  if (tid <= kEmb) s_lut[tid] = A.lut[tid];
  for (int q = A.q0 + blockIdx.x; q < A.q1; q += gridDim.x) {
    __syncthreads();
    for (int e = tid; e < N; e += kQueryThreads) { scores[e] = 0.0f; hit[e] = 0; }
    __syncthreads();
    const int64_t f0 = A.q_off[q], f1 = A.q_off[q + 1];
    double self = 0.0;
    for (int64_t f = f0 + warp; f < f1; f += kQueryThreads / 32) {
      const float lo = A.P[(f - A.f_base) * kEmb + lane], hi = A.P[(f - A.f_base) * kEmb + lane + 32];
      for (int k = 0; k < A.K; ++k) {
        const int w = A.nn[f * A.K + k];
        if (w == kInvalidWord) continue;
        const float idf = A.idf[w], sq = idf * idf;
        if (lane == 0) self += (double)sq;
        const uint32_t s = A.word_start[w], e = A.word_start[w + 1];
        if (!A.has_emb[w] || e == s) continue;  // InvertedFile::IsUsable
        const unsigned long long bits = signature_bits(lo, hi, A.thr + (size_t)w * kEmb, lane);
        int carry_img = -1;  // image of the previous chunk's last entry
        for (uint32_t base = s; base < e; base += 32) {
          const uint32_t a = base + lane;
          const bool valid = a < e;
          int img = -2 - lane;  // padding lanes: distinct, equal to no image
          float wgt = 0.0f;
          int vote = 0;
          if (valid) {
            const uint4 raw = *reinterpret_cast<const uint4*>(A.entries + a);
            img = (int)raw.x;
            const unsigned long long eb = (unsigned long long)raw.z | ((unsigned long long)raw.w << 32);
            const int hd = __popcll(bits ^ eb);
            if (hd <= kMaxHamming) { wgt = s_lut[hd]; vote = 1; }
          }
          int prev = __shfl_up_sync(0xffffffffu, img, 1);
          if (lane == 0) prev = carry_img;
          const bool head = valid && img != prev;
          float sc = wgt;
          int votes = vote;
          for (int step = 1; step < 32; ++step) {
            const int oi = __shfl_down_sync(0xffffffffu, img, step);
            const float ow = __shfl_down_sync(0xffffffffu, wgt, step);
            const int ov = __shfl_down_sync(0xffffffffu, vote, step);
            const bool ext = head && lane + step < 32 && oi == img;
            if (ext) { sc += ow; votes += ov; }
            if (!__any_sync(0xffffffffu, ext)) break;
          }
          const int last_img = __shfl_sync(0xffffffffu, img, 31);
          if (head && img == last_img) {  // the run may continue behind the chunk: it stays this lane's
            for (uint32_t b = base + 32; b < e; ++b) {
              const uint4 raw = *reinterpret_cast<const uint4*>(A.entries + b);
              if ((int)raw.x != img) break;
              const int hd = __popcll(bits ^ ((unsigned long long)raw.z | ((unsigned long long)raw.w << 32)));
              if (hd <= kMaxHamming) { sc += s_lut[hd]; votes += 1; }
            }
          }
          if (head && votes > 0) {
            float v = sc / sqrtf((float)votes);
            v *= sq;
            atomicAdd(scores + img, v);
            hit[img] = 1;
          }
          carry_img = last_img;
        }
      }
    }
    if (lane == 0) s_self[warp] = self;
    __syncthreads();
    double self_all = 0.0;
    for (int k = 0; k < kQueryThreads / 32; ++k) self_all += s_self[k];
    const float self_similarity = (float)self_all;
    const float nw = self_similarity > 0.0f ? 1.0f / sqrtf(self_similarity) : 1.0f;
    for (int e = tid; e < N; e += kQueryThreads) scores[e] = hit[e] ? scores[e] * (nw * A.norm[e]) : -1.0f;
    __syncthreads();
    // top max_num_images by (score desc, image id asc): repeated block-wide arg-max over the score array
    int count = 0;
    for (int r = 0; r < A.max_num_images; ++r) {
      float best = -1.0f;
      int besti = 0x7fffffff;
      for (int e = tid; e < N; e += kQueryThreads) {
        const float v = scores[e];
        if (v > best || (v == best && v >= 0.0f && e < besti)) { best = v; besti = e; }
      }
      s_best[tid] = best;
      s_besti[tid] = besti;
      __syncthreads();
      for (int o = kQueryThreads / 2; o > 0; o >>= 1) {
        if (tid < o) {
          const float v = s_best[tid + o];
          const int vi = s_besti[tid + o];
          if (v > s_best[tid] || (v == s_best[tid] && vi < s_besti[tid])) { s_best[tid] = v; s_besti[tid] = vi; }
        }
        __syncthreads();
      }
      const float top = s_best[0];
      const int topi = s_besti[0];
      __syncthreads();
      if (top < 0.0f) break;
      if (tid == 0) {
        A.out_ids[(size_t)q * A.max_num_images + r] = topi;
        A.out_scores[(size_t)q * A.max_num_images + r] = top;
        scores[topi] = -1.0f;
      }
      count = r + 1;
      __syncthreads();
    }
    if (tid == 0) A.out_count[q] = count;
  }
}

}  // namespace rt
}  // namespace b2

using namespace b2;

struct b2_retrieval {
  int device = 0, n_sm = 148;
  cudaStream_t stream = nullptr;
  int n_words = 0;
  uint8_t* d_words = nullptr;   // [n_blk * 128][128], rows >= n_words zero
  int32_t* d_word_sq = nullptr; // [n_blk * 128], INT_MAX at the padding rows
  uint32_t n_blk = 0;
  CUtensorMap tmap_words;
  float *d_proj = nullptr, *d_thr = nullptr, *d_idf = nullptr, *d_norm = nullptr, *d_lut = nullptr;
  uint8_t* d_has = nullptr;
  uint32_t* d_word_start = nullptr;  // [n_words + 1]
  rt::Entry* d_entries = nullptr;
  int n_images = 0;
  int64_t n_entries = 0;
  // the indexed images' descriptors and nearest words (query_all re-uses them)
  uint8_t* d_desc = nullptr;
  bool own_desc = false;
  int64_t* d_img_off = nullptr;
  int64_t n_desc = 0;
  int32_t* d_nn = nullptr;  // [n_desc * nn_k]
  int nn_k = 0;
  float* d_P = nullptr;        // projections of one batch of images' descriptors [batch_desc * 64]
  size_t p_floats = 0;
  std::vector<int64_t> h_img_off;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_seconds[3] = {0, 0, 0};  // word search, index build, query
};

namespace {
template <class T> void fr(T*& p) { if (p) cudaFree(p); p = nullptr; }
// Call-scoped device buffers: freed on every way out of the function (the pointer VARIABLES are tracked, so a buffer that
// was swapped with a handle-owned one is the one released).
struct ScopedDev {
  std::vector<void**> vars;
  std::vector<cudaEvent_t> events;
  template <class T> cudaError_t alloc(T** var, size_t bytes) {
    *var = nullptr;
    const cudaError_t e = cudaMalloc(reinterpret_cast<void**>(var), bytes);
    if (e == cudaSuccess) vars.push_back(reinterpret_cast<void**>(var));
    return e;
  }
  cudaError_t event(cudaEvent_t* ev) {
    const cudaError_t e = cudaEventCreate(ev);
    if (e == cudaSuccess) events.push_back(*ev);
    return e;
  }
  ~ScopedDev() {
    for (void** v : vars) if (*v) cudaFree(*v);
    for (cudaEvent_t ev : events) cudaEventDestroy(ev);
  }
};

// SIMT statement of the word search (test seam)
int knn_simt(b2_retrieval* r, const uint8_t* d_desc, int64_t n, int k, int32_t* d_out) {
  if (n == 0) return B2_OK;
  const unsigned g = (unsigned)((n + rt::kKnnThreads - 1) / rt::kKnnThreads);
  cudaStream_t s = r->stream;
  switch (k) {
    case 1: rt::word_knn_kernel<1><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 2: rt::word_knn_kernel<2><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 3: rt::word_knn_kernel<3><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 4: rt::word_knn_kernel<4><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 5: rt::word_knn_kernel<5><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 6: rt::word_knn_kernel<6><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    case 7: rt::word_knn_kernel<7><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
    default: rt::word_knn_kernel<8><<<g, rt::kKnnThreads, 0, s>>>(d_desc, n, r->d_words, r->d_word_sq, r->n_words, d_out); break;
  }
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  return B2_OK;
}
int knn(b2_retrieval* r, const uint8_t* d_desc, int64_t n, int k, int32_t* d_out) {
  if (n == 0) return B2_OK;
  B2_CUDA(launch_word_knn_tc(r->tmap_words, d_desc, (long long)n, r->d_word_sq, r->n_blk, k, d_out, r->n_sm, r->stream));
  count_launches(1);
  return B2_OK;
}

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// [rows][128] uint8, box 128 rows x 128 bytes, SWIZZLE_128B (one K-major operand block of the contraction)
int make_words_tmap(CUtensorMap* tm, void* base, uint64_t rows) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) return set_error(B2_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)rt::kDim, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)rt::kDim};
  const cuuint32_t box[2] = {(cuuint32_t)rt::kDim, 128u};
  const cuuint32_t estr[2] = {1, 1};
  CUresult res = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (res != CUDA_SUCCESS) return set_error(B2_ERR_CUDA, "cuTensorMapEncodeTiled failed for the word pool");
  return B2_OK;
}

// Image batches whose descriptors' projections fit the projection buffer (<= kProjBudget bytes); a batch holds at least
// one image.  Returns the image boundaries.
constexpr size_t kProjBudget = (size_t)2 << 30;
std::vector<int> image_batches(const std::vector<int64_t>& off) {
  std::vector<int> b{0};
  const int64_t cap = (int64_t)(kProjBudget / (rt::kEmb * 4));
  const int n = (int)off.size() - 1;
  int i = 0;
  while (i < n) {
    int j = i + 1;
    while (j < n && off[j + 1] - off[i] <= cap) ++j;
    b.push_back(j);
    i = j;
  }
  return b;
}
int project_batch(b2_retrieval* r, int64_t f0, int64_t f1) {
  const size_t need = (size_t)std::max<int64_t>(f1 - f0, 1) * rt::kEmb;
  if (need > r->p_floats) {
    fr(r->d_P);
    r->p_floats = 0;
    B2_CUDA(cudaMalloc(&r->d_P, need * 4));
    r->p_floats = need;
  }
  if (f1 > f0) {
    const int64_t warps = f1 - f0;
    const int grid = (int)std::min<int64_t>((warps + rt::kProjWarps - 1) / rt::kProjWarps, (int64_t)r->n_sm * 8);
    rt::project_kernel<<<grid, 32 * rt::kProjWarps, 0, r->stream>>>(r->d_desc, f0, f1, r->d_proj, r->d_P);
    B2_CUDA(cudaGetLastError());
    count_launches(1);
  }
  return B2_OK;
}
}  // namespace

int b2_retrieval_create(int device, b2_retrieval** out) {
  if (!out) return set_error(B2_ERR_INVALID, "NULL argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return set_error(B2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
  if (device < 0 || device >= n_dev) return set_error(B2_ERR_INVALID, "bad device index");
  b2_retrieval* r = new b2_retrieval;
  r->device = device;
  cudaDeviceProp prop;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  r->n_sm = prop.multiProcessorCount;
  B2_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  B2_CUDA(cudaEventCreate(&r->ev0));
  B2_CUDA(cudaEventCreate(&r->ev1));
  *out = r;
  return B2_OK;
}

int b2_retrieval_destroy(b2_retrieval* r) {
  if (!r) return B2_OK;
  cudaSetDevice(r->device);
  if (r->stream) cudaStreamSynchronize(r->stream);
  fr(r->d_words); fr(r->d_word_sq); fr(r->d_proj); fr(r->d_thr); fr(r->d_idf); fr(r->d_norm); fr(r->d_lut); fr(r->d_has);
  fr(r->d_word_start); fr(r->d_entries); fr(r->d_img_off); fr(r->d_nn); fr(r->d_P);
  if (r->own_desc) fr(r->d_desc);
  if (r->ev0) cudaEventDestroy(r->ev0);
  if (r->ev1) cudaEventDestroy(r->ev1);
  if (r->stream) cudaStreamDestroy(r->stream);
  delete r;
  return B2_OK;
}

int b2_retrieval_set_vocabulary(b2_retrieval* r, int32_t n_words, const uint8_t* words, const float* proj, const float* thresholds,
                                const uint8_t* has_embedding) {
  if (!r || n_words <= 0 || !words || !proj || !thresholds || !has_embedding) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  cudaStream_t s = r->stream;
  fr(r->d_words); fr(r->d_word_sq); fr(r->d_proj); fr(r->d_thr); fr(r->d_idf); fr(r->d_lut); fr(r->d_has); fr(r->d_word_start);
  r->n_words = n_words;
  r->n_blk = (uint32_t)((n_words + 127) / 128);
  const size_t pad_rows = (size_t)r->n_blk * 128;
  B2_CUDA(cudaMalloc(&r->d_words, pad_rows * rt::kDim));
  B2_CUDA(cudaMalloc(&r->d_word_sq, pad_rows * 4));
  B2_CUDA(cudaMemsetAsync(r->d_words, 0, pad_rows * rt::kDim, s));
  B2_TRY(make_words_tmap(&r->tmap_words, r->d_words, pad_rows));
  B2_CUDA(cudaMalloc(&r->d_proj, rt::kEmb * rt::kDim * 4));
  B2_CUDA(cudaMalloc(&r->d_thr, (size_t)n_words * rt::kEmb * 4));
  B2_CUDA(cudaMalloc(&r->d_idf, (size_t)n_words * 4));
  B2_CUDA(cudaMalloc(&r->d_has, (size_t)n_words));
  B2_CUDA(cudaMalloc(&r->d_lut, (rt::kEmb + 1) * 4));
  B2_CUDA(cudaMalloc(&r->d_word_start, ((size_t)n_words + 1) * 4));
  B2_CUDA(cudaMemcpyAsync(r->d_words, words, (size_t)n_words * rt::kDim, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(r->d_proj, proj, rt::kEmb * rt::kDim * 4, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(r->d_thr, thresholds, (size_t)n_words * rt::kEmb * 4, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(r->d_has, has_embedding, (size_t)n_words, cudaMemcpyHostToDevice, s));
  float lut[rt::kEmb + 1];  // HammingDistWeightFunctor<64, 16>: exp(-h^2 / sigma^2) up to 1.5 sigma, host libm as the reference
  for (int n = 0; n <= rt::kEmb; ++n) {
    const float h = (float)n;
    lut[n] = h <= (float)rt::kMaxHamming ? std::exp(-h * h / 256.0f) : 0.0f;
  }
  B2_CUDA(cudaMemcpyAsync(r->d_lut, lut, sizeof lut, cudaMemcpyHostToDevice, s));
  rt::word_norms_kernel<<<(unsigned)((pad_rows + 255) / 256), 256, 0, s>>>(r->d_words, n_words, (int)pad_rows, r->d_word_sq);
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  B2_CUDA(cudaStreamSynchronize(s));
  return B2_OK;
}

// VisualIndex::Add for every image (IndexOptions::num_neighbors = 1) + Prepare().  k_query = the number of nearest
// words kept per descriptor for b2_retrieval_query_all (QueryOptions::num_neighbors; the first of them is the indexing word).
static int index_impl(b2_retrieval* r, int32_t n_images, const uint8_t* d_desc, bool own, const int64_t* desc_off_host, int32_t k_query,
                      const int32_t* d_word_ids = nullptr) {
  if (r->n_words == 0) return set_error(B2_ERR_INVALID, "no vocabulary set");
  if (k_query < 1 || k_query > rt::kMaxK) return set_error(B2_ERR_INVALID, "num_neighbors must be in [1, 8]");
  cudaStream_t s = r->stream;
  const int64_t n = desc_off_host[n_images];
  if (r->own_desc) fr(r->d_desc);
  fr(r->d_img_off); fr(r->d_nn); fr(r->d_entries); fr(r->d_norm);
  r->d_desc = const_cast<uint8_t*>(d_desc);
  r->own_desc = own;
  r->n_images = n_images;
  r->n_desc = n;
  r->nn_k = k_query;
  B2_CUDA(cudaMalloc(&r->d_img_off, ((size_t)n_images + 1) * 8));
  B2_CUDA(cudaMemcpyAsync(r->d_img_off, desc_off_host, ((size_t)n_images + 1) * 8, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMalloc(&r->d_nn, std::max<size_t>((size_t)n * k_query, 1) * 4));
  B2_CUDA(cudaMalloc(&r->d_norm, std::max<size_t>(n_images, 1) * 4));
  B2_CUDA(cudaEventRecord(r->ev0, s));
  if (d_word_ids) {  // the word search ran elsewhere (sharded over the ranks of a multi-GPU job, all-gathered by the caller)
    if (n > 0) B2_CUDA(cudaMemcpyAsync(r->d_nn, d_word_ids, (size_t)n * k_query * 4, cudaMemcpyDeviceToDevice, s));
  } else {
    B2_TRY(knn(r, r->d_desc, n, k_query, r->d_nn));
  }
  B2_CUDA(cudaEventRecord(r->ev1, s));
  // per-feature image / index-in-image (host: cheap, once)
  std::vector<int32_t> fimg((size_t)std::max<int64_t>(n, 1)), fidx((size_t)std::max<int64_t>(n, 1));
  int n_nonempty = 0;
  for (int i = 0; i < n_images; ++i) {
    n_nonempty += desc_off_host[i + 1] > desc_off_host[i] ? 1 : 0;
    for (int64_t f = desc_off_host[i]; f < desc_off_host[i + 1]; ++f) { fimg[f] = i; fidx[f] = (int32_t)(f - desc_off_host[i]); }
  }
  int32_t *d_fimg = nullptr, *d_fidx = nullptr;
  uint32_t *d_count = nullptr, *d_cursor = nullptr;
  rt::Entry* d_staged = nullptr;
  ScopedDev tmp;
  B2_CUDA(tmp.alloc(&d_fimg, fimg.size() * 4));
  B2_CUDA(tmp.alloc(&d_fidx, fidx.size() * 4));
  B2_CUDA(tmp.alloc(&d_count, ((size_t)r->n_words + 1) * 4));
  B2_CUDA(tmp.alloc(&d_cursor, (size_t)r->n_words * 4));
  B2_CUDA(tmp.alloc(&d_staged, std::max<size_t>(n, 1) * sizeof(rt::Entry)));
  B2_CUDA(cudaMalloc(&r->d_entries, std::max<size_t>(n, 1) * sizeof(rt::Entry)));
  B2_CUDA(cudaMemcpyAsync(d_fimg, fimg.data(), fimg.size() * 4, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemcpyAsync(d_fidx, fidx.data(), fidx.size() * 4, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaMemsetAsync(d_count, 0, ((size_t)r->n_words + 1) * 4, s));
  B2_CUDA(cudaMemsetAsync(d_cursor, 0, (size_t)r->n_words * 4, s));
  cudaEvent_t e2, e3;
  B2_CUDA(tmp.event(&e2));
  B2_CUDA(tmp.event(&e3));
  B2_CUDA(cudaEventRecord(e2, s));
  r->h_img_off.assign(desc_off_host, desc_off_host + n_images + 1);
  if (n > 0) {
    const std::vector<int> bat = image_batches(r->h_img_off);
    for (size_t b = 0; b + 1 < bat.size(); ++b) {
      const int64_t f0 = r->h_img_off[bat[b]], f1 = r->h_img_off[bat[b + 1]];
      if (f1 == f0) continue;
      B2_TRY(project_batch(r, f0, f1));
      const int grid = (int)std::min<int64_t>((f1 - f0 + rt::kSigWarps - 1) / rt::kSigWarps, (int64_t)r->n_sm * 8);
      rt::index_sig_kernel<<<grid, 32 * rt::kSigWarps, 0, s>>>(r->d_P, f0, f1, r->d_nn, k_query, r->d_thr, d_fimg, d_fidx, d_count, d_staged);
      B2_CUDA(cudaGetLastError());
      count_launches(1);
    }
  }
  uint32_t* d_total = d_count + r->n_words;  // the scan writes the grand total behind the last count's slot
  B2_CUDA(launch_scan_u32(d_count, r->n_words, r->d_word_start, d_total, s));
  B2_CUDA(cudaMemcpyAsync(r->d_word_start + r->n_words, d_total, 4, cudaMemcpyDeviceToDevice, s));
  if (n > 0) {
    rt::scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_staged, r->d_nn, k_query, n, r->d_word_start, d_cursor, r->d_entries);
    B2_CUDA(cudaGetLastError());
    // rank sort out of place: entries -> staged (sorted), then staged is the index
    rt::sort_words_kernel<<<(unsigned)(((size_t)r->n_words * 32 + 255) / 256), 256, 0, s>>>(r->d_entries, d_staged, r->d_word_start, r->n_words, n_nonempty, r->d_idf);
    B2_CUDA(cudaGetLastError());
    std::swap(r->d_entries, d_staged);
    rt::image_norm_kernel<<<n_images, 256, 0, s>>>(r->d_nn, k_query, r->d_img_off, n_images, r->d_idf, r->d_norm);
    B2_CUDA(cudaGetLastError());
  } else {
    B2_CUDA(cudaMemsetAsync(r->d_idf, 0, (size_t)r->n_words * 4, s));
  }
  count_launches(4);
  B2_CUDA(cudaEventRecord(e3, s));
  B2_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  r->last_seconds[0] = ms * 1e-3;
  B2_CUDA(cudaEventElapsedTime(&ms, e2, e3));
  r->last_seconds[1] = ms * 1e-3;
  r->n_entries = n;
  return B2_OK;
}

int b2_retrieval_index_images(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors, const int64_t* desc_offsets, int32_t num_neighbors_query) {
  if (!r || n_images < 0 || !desc_offsets || (desc_offsets[n_images] > 0 && !descriptors)) return set_error(B2_ERR_INVALID, "bad argument");
  if (r->n_words == 0) return set_error(B2_ERR_INVALID, "no vocabulary set");
  if (num_neighbors_query < 1 || num_neighbors_query > rt::kMaxK) return set_error(B2_ERR_INVALID, "num_neighbors must be in [1, 8]");
  B2_CUDA(cudaSetDevice(r->device));
  const int64_t n = desc_offsets[n_images];
  uint8_t* d = nullptr;
  B2_CUDA(cudaMalloc(&d, std::max<size_t>((size_t)n * rt::kDim, 16)));
  if (n && cudaMemcpyAsync(d, descriptors, (size_t)n * rt::kDim, cudaMemcpyHostToDevice, r->stream) != cudaSuccess) {
    cudaFree(d);
    return set_error(B2_ERR_CUDA, "descriptor upload failed");
  }
  return index_impl(r, n_images, d, true, desc_offsets, num_neighbors_query);   // the handle owns d from here on
}

int b2_retrieval_index_images_device(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors_dev, const int64_t* desc_offsets_host,
                                     int32_t num_neighbors_query) {
  if (!r || n_images < 0 || !desc_offsets_host || !descriptors_dev) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  return index_impl(r, n_images, descriptors_dev, false, desc_offsets_host, num_neighbors_query);
}

// Multi-GPU: the word search of a rank's share of the descriptors (n_desc rows at descriptors_dev) into the caller's device
// buffer, and the index built from word ids that the ranks exchanged (all-gather; the one collective of the stage).
int b2_retrieval_word_search_device(b2_retrieval* r, const uint8_t* descriptors_dev, int64_t n_desc, int32_t num_neighbors,
                                    int32_t* out_word_ids_dev) {
  if (!r || n_desc < 0 || (n_desc > 0 && (!descriptors_dev || !out_word_ids_dev))) return set_error(B2_ERR_INVALID, "bad argument");
  if (r->n_words == 0) return set_error(B2_ERR_INVALID, "no vocabulary set");
  if (num_neighbors < 1 || num_neighbors > rt::kMaxK) return set_error(B2_ERR_INVALID, "num_neighbors must be in [1, 8]");
  B2_CUDA(cudaSetDevice(r->device));
  B2_CUDA(cudaEventRecord(r->ev0, r->stream));
  B2_TRY(knn(r, descriptors_dev, n_desc, num_neighbors, out_word_ids_dev));
  B2_CUDA(cudaEventRecord(r->ev1, r->stream));
  B2_CUDA(cudaStreamSynchronize(r->stream));
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  r->last_seconds[0] = ms * 1e-3;
  return B2_OK;
}
int b2_retrieval_index_images_words_device(b2_retrieval* r, int32_t n_images, const uint8_t* descriptors_dev, const int64_t* desc_offsets_host,
                                           int32_t num_neighbors_query, const int32_t* word_ids_dev) {
  if (!r || n_images < 0 || !desc_offsets_host || !descriptors_dev || !word_ids_dev) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  const double searched = r->last_seconds[0];
  const int rc = index_impl(r, n_images, descriptors_dev, false, desc_offsets_host, num_neighbors_query, word_ids_dev);
  r->last_seconds[0] = searched;   // the word search was this rank's b2_retrieval_word_search_device call
  return rc;
}

// VisualIndex::Query of every indexed image against the index (VocabSimilarityGraph::Run's retrieval loop).
int b2_retrieval_query_all(b2_retrieval* r, int32_t max_num_images, int32_t* out_ids, float* out_scores, int32_t* out_counts) {
  if (!r) return set_error(B2_ERR_INVALID, "bad argument");
  return b2_retrieval_query_range(r, 0, r->n_images, max_num_images, out_ids, out_scores, out_counts);
}

// The same for the query images [q0, q1) only (a rank's share of a multi-GPU job); outputs are indexed from q0.
int b2_retrieval_query_range(b2_retrieval* r, int32_t q0, int32_t q1, int32_t max_num_images, int32_t* out_ids, float* out_scores,
                             int32_t* out_counts) {
  if (!r || max_num_images <= 0 || !out_ids || !out_scores || !out_counts) return set_error(B2_ERR_INVALID, "bad argument");
  if (!r->d_entries || r->n_images == 0) return set_error(B2_ERR_INVALID, "no images indexed");
  if (q0 < 0 || q1 < q0 || q1 > r->n_images) return set_error(B2_ERR_INVALID, "query range outside the indexed images");
  if (q1 == q0) return B2_OK;
  B2_CUDA(cudaSetDevice(r->device));
  cudaStream_t s = r->stream;
  const int N = r->n_images, Q = q1 - q0;
  int32_t *d_ids = nullptr, *d_cnt = nullptr;
  float* d_sc = nullptr;
  ScopedDev tmp;
  B2_CUDA(tmp.alloc(&d_ids, (size_t)Q * max_num_images * 4));
  B2_CUDA(tmp.alloc(&d_sc, (size_t)Q * max_num_images * 4));
  B2_CUDA(tmp.alloc(&d_cnt, (size_t)Q * 4));
  B2_CUDA(cudaMemsetAsync(d_ids, 0xff, (size_t)Q * max_num_images * 4, s));
  B2_CUDA(cudaMemsetAsync(d_sc, 0, (size_t)Q * max_num_images * 4, s));
  const size_t with_scores = (size_t)N * 4 + (size_t)N + 16;
  const bool in_smem = with_scores <= 200 * 1024;
  const size_t smem = in_smem ? with_scores : (size_t)N + 16;
  int ctas_per_sm = (int)std::min<size_t>(6, (220 * 1024) / (smem + 4 * 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int max_grid = r->n_sm * ctas_per_sm;
  float* d_scratch = nullptr;
  if (!in_smem) B2_CUDA(tmp.alloc(&d_scratch, (size_t)max_grid * N * 4));
  B2_CUDA(cudaFuncSetAttribute(rt::query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rt::QueryArgs A;
  A.q_off = r->d_img_off; A.nn = r->d_nn; A.K = r->nn_k; A.thr = r->d_thr; A.has_emb = r->d_has;
  A.idf = r->d_idf; A.word_start = r->d_word_start; A.entries = r->d_entries; A.norm = r->d_norm; A.lut = r->d_lut; A.n_index_images = N;
  // the kernel indexes its outputs by the absolute query image: bias the pointers so that image q0 lands at slot 0
  A.max_num_images = max_num_images; A.out_ids = d_ids - (size_t)q0 * max_num_images; A.out_scores = d_sc - (size_t)q0 * max_num_images;
  A.out_count = d_cnt - q0; A.score_scratch = d_scratch;
  B2_CUDA(cudaEventRecord(r->ev0, s));
  const std::vector<int64_t> sub(r->h_img_off.begin() + q0, r->h_img_off.begin() + q1 + 1);
  const std::vector<int> bat = image_batches(sub);
  for (size_t b = 0; b + 1 < bat.size(); ++b) {
    const int i0 = q0 + bat[b], i1 = q0 + bat[b + 1];
    const int64_t f0 = r->h_img_off[i0], f1 = r->h_img_off[i1];
    B2_TRY(project_batch(r, f0, f1));
    A.P = r->d_P; A.f_base = f0; A.q0 = i0; A.q1 = i1;
    const int grid = std::min(i1 - i0, max_grid);
    rt::query_kernel<<<grid, rt::kQueryThreads, smem, s>>>(A, in_smem ? 1 : 0);
    B2_CUDA(cudaGetLastError());
    count_launches(1);
  }
  B2_CUDA(cudaEventRecord(r->ev1, s));
  B2_CUDA(cudaMemcpyAsync(out_ids, d_ids, (size_t)Q * max_num_images * 4, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(out_scores, d_sc, (size_t)Q * max_num_images * 4, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaMemcpyAsync(out_counts, d_cnt, (size_t)Q * 4, cudaMemcpyDeviceToHost, s));
  B2_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, r->ev0, r->ev1));
  r->last_seconds[2] = ms * 1e-3;
  return B2_OK;
}

// Test seams: the nearest words of the indexed descriptors [n_desc * k], the inverted file of one word.
int b2_retrieval_debug_word_ids(b2_retrieval* r, int32_t* out) {
  if (!r || !out || !r->d_nn) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  B2_CUDA(cudaMemcpy(out, r->d_nn, (size_t)r->n_desc * r->nn_k * 4, cudaMemcpyDeviceToHost));
  return B2_OK;
}
int b2_retrieval_debug_word_ids_simt(b2_retrieval* r, int32_t* out) {
  if (!r || !out || !r->d_nn) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  int32_t* tmp = nullptr;
  B2_CUDA(cudaMalloc(&tmp, std::max<size_t>((size_t)r->n_desc * r->nn_k, 1) * 4));
  int rc = knn_simt(r, r->d_desc, r->n_desc, r->nn_k, tmp);
  if (rc == B2_OK && cudaStreamSynchronize(r->stream) != cudaSuccess) rc = set_error(B2_ERR_CUDA, "word search (SIMT seam) failed");
  if (rc == B2_OK && cudaMemcpy(out, tmp, (size_t)r->n_desc * r->nn_k * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    rc = set_error(B2_ERR_CUDA, "copy failed");
  cudaFree(tmp);
  return rc;
}
int b2_retrieval_debug_index(b2_retrieval* r, uint32_t* word_start, int32_t* entry_image, int32_t* entry_feature, uint64_t* entry_bits, float* idf,
                             float* norm) {
  if (!r || !r->d_entries) return set_error(B2_ERR_INVALID, "bad argument");
  B2_CUDA(cudaSetDevice(r->device));
  std::vector<rt::Entry> e((size_t)std::max<int64_t>(r->n_entries, 1));
  std::vector<uint32_t> ws((size_t)r->n_words + 1);
  B2_CUDA(cudaMemcpy(ws.data(), r->d_word_start, ws.size() * 4, cudaMemcpyDeviceToHost));
  const uint32_t total = ws[r->n_words];
  B2_CUDA(cudaMemcpy(e.data(), r->d_entries, (size_t)total * sizeof(rt::Entry), cudaMemcpyDeviceToHost));
  if (word_start) memcpy(word_start, ws.data(), ws.size() * 4);
  for (uint32_t k = 0; k < total; ++k) {
    if (entry_image) entry_image[k] = e[k].image;
    if (entry_feature) entry_feature[k] = e[k].feature;
    if (entry_bits) entry_bits[k] = e[k].bits;
  }
  if (idf) B2_CUDA(cudaMemcpy(idf, r->d_idf, (size_t)r->n_words * 4, cudaMemcpyDeviceToHost));
  if (norm) B2_CUDA(cudaMemcpy(norm, r->d_norm, (size_t)r->n_images * 4, cudaMemcpyDeviceToHost));
  return B2_OK;
}
int b2_retrieval_last_timing(b2_retrieval* r, double* word_search_s, double* index_build_s, double* query_s) {
  if (!r) return set_error(B2_ERR_INVALID, "NULL argument");
  if (word_search_s) *word_search_s = r->last_seconds[0];
  if (index_build_s) *index_build_s = r->last_seconds[1];
  if (query_s) *query_s = r->last_seconds[2];
  return B2_OK;
}
