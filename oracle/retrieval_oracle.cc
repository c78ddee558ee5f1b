// CPU oracle of the vocabulary-tree retrieval stage (candidate image pairs) -- TEST INFRASTRUCTURE, never linked into or
// called by the product (dagsfm_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it.
//
// Restates, function by function:
//   VisualIndex::Add / Query / QueryAndFindWordIds          src/retrieval/visual_index.h:207-262, 505-540
//   InvertedIndex::AddEntry / Finalize / Query /
//     ComputeSelfSimilarity / ComputeWeightsAndNormalizationConstants   src/retrieval/inverted_index.h:229-283, 307-319, 417-441
//   InvertedFile::ConvertToBinaryDescriptor / ComputeIDFWeight / ScoreFeature / ComputeImageSelfSimilarities
//                                                           src/retrieval/inverted_file.h:256-264, 266-277, 305-366, 376-382
//   HammingDistWeightFunctor<64, 16>                        src/retrieval/utils.h:52-82
//   VocabSimilarityGraph::Run (pairs image_id < other)       src/graph/similarity_graph.cpp:101-200
//
// Two stated differences from the reference:
//   * FindWordIds asks a flann::AutotunedIndex (the reference vendors FLANN under lib/FLANN): its algorithm and parameters
//     come from TIMING experiments on the host (target_precision 0.95), its trees are randomised and the search stops after
//     num_checks = 256 leaves -- approximate and not reproducible.  Here the nearest words are EXACT (squared L2 over the
//     uint8 descriptors, ties -> lower word id).  PINNED: this function returns the same ids in the same order, ties
//     included, as the reference's own FLANN in exact mode (flann::LinearIndex over flann::L2<uint8>, built from lib/FLANN
//     into oracle/_ref/libflann_ref.so; golden vectors tests/golden/retrieval_flann_linear.npz), and on the synthetic
//     collection of the tests the reference's autotuned index returns exactly these words too (2 000 / 2 000 queries,
//     k = 5); on uniform random descriptors it finds 39 % of the nearest words.
//   * the Hamming projection proj * descriptor is an Eigen float product whose summation order is Eigen's (Eigen is not
//     vendored and not installed); here it is the plain left-to-right float sum (both the GPU path and this oracle), which
//     can flip a signature bit only when a projected value equals its threshold to the last ulp.
// Parity status: word search pinned as above.  For the voting stage the reference's tests (visual_index_test.cc) pin
// structure only -- ranking of an image against itself, result sizes under max_num_images -- and those are replayed in
// tests/test_oracle_retrieval.py; there are no numeric golden vectors for scores: "parity unpinned" for the score values.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

constexpr int kDim = 128, kEmb = 64, kSigma = 16, kMaxHamming = 24;  // static_cast<size_t>(1.5f * kSigma)

struct Entry { int image, feature; uint64_t bits; };

struct Index {
  int n_words = 0;
  std::vector<uint8_t> words;        // [n_words * 128]
  std::vector<float> proj;           // [64 * 128] row-major
  std::vector<float> thr;            // [n_words * 64]
  std::vector<uint8_t> has_emb;      // [n_words]
  std::vector<std::vector<Entry>> files;
  std::vector<float> idf;            // [n_words]
  std::unordered_map<int, float> norm;
  int n_images = 0;
  float lut[kEmb + 1];
};

void nearest_words(const Index& X, const uint8_t* d, int k, int* out) {
  std::vector<std::pair<int64_t, int>> best;  // (distance, word), ascending
  for (int w = 0; w < X.n_words; ++w) {
    const uint8_t* c = &X.words[(size_t)w * kDim];
    int64_t s = 0;
    for (int j = 0; j < kDim; ++j) { const int e = (int)d[j] - (int)c[j]; s += e * e; }
    if ((int)best.size() < k) { best.emplace_back(s, w); std::push_heap(best.begin(), best.end()); }
    else if (std::make_pair(s, w) < best.front()) { std::pop_heap(best.begin(), best.end()); best.back() = {s, w}; std::push_heap(best.begin(), best.end()); }
  }
  std::sort(best.begin(), best.end());
  for (int n = 0; n < k; ++n) out[n] = n < (int)best.size() ? best[n].second : 0x7fffffff;  // kInvalidWordId
}

uint64_t signature(const Index& X, int word, const uint8_t* d) {
  uint64_t b = 0;
  for (int i = 0; i < kEmb; ++i) {
    float s = 0.0f;
    for (int j = 0; j < kDim; ++j) s += X.proj[(size_t)i * kDim + j] * (float)d[j];
    if (s > X.thr[(size_t)word * kEmb + i]) b |= 1ull << i;
  }
  return b;
}

}  // namespace

extern "C" {

void* orc_retrieval_create(int n_words, const uint8_t* words, const float* proj, const float* thr, const uint8_t* has_emb) {
  Index* X = new Index;
  X->n_words = n_words;
  X->words.assign(words, words + (size_t)n_words * kDim);
  X->proj.assign(proj, proj + kEmb * kDim);
  X->thr.assign(thr, thr + (size_t)n_words * kEmb);
  X->has_emb.assign(has_emb, has_emb + n_words);
  X->files.resize(n_words);
  X->idf.assign(n_words, 0.0f);
  const float sigma_squared = kSigma * kSigma;
  for (int n = 0; n <= kEmb; ++n) {
    const float h = (float)n;
    X->lut[n] = h <= kMaxHamming ? std::exp(-h * h / sigma_squared) : 0.0f;
  }
  return X;
}
void orc_retrieval_destroy(void* p) { delete (Index*)p; }
void orc_retrieval_lut(void* p, float* out) { memcpy(out, ((Index*)p)->lut, sizeof(float) * (kEmb + 1)); }

// FindWordIds (exact): word_ids [n * k]
void orc_retrieval_word_ids(void* p, int n, const uint8_t* desc, int k, int* word_ids) {
  const Index& X = *(Index*)p;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n; ++i) nearest_words(X, desc + (size_t)i * kDim, k, word_ids + (size_t)i * k);
}
void orc_retrieval_signatures(void* p, int n, const uint8_t* desc, const int* word, uint64_t* out) {
  const Index& X = *(Index*)p;
  for (int i = 0; i < n; ++i) out[i] = signature(X, word[i], desc + (size_t)i * kDim);
}

// VisualIndex::Add with IndexOptions::num_neighbors = 1
void orc_retrieval_add(void* p, int image_id, int n, const uint8_t* desc) {
  Index& X = *(Index*)p;
  X.n_images += 1;
  std::vector<int> w(n);
  orc_retrieval_word_ids(p, n, desc, 1, w.data());
  for (int i = 0; i < n; ++i)
    if (w[i] != 0x7fffffff) X.files[w[i]].push_back({image_id, i, signature(X, w[i], desc + (size_t)i * kDim)});
}

// VisualIndex::Prepare -> InvertedIndex::Finalize
void orc_retrieval_prepare(void* p) {
  Index& X = *(Index*)p;
  std::unordered_map<int, double> self;
  std::vector<int> ids;
  for (auto& f : X.files) {
    std::stable_sort(f.begin(), f.end(), [](const Entry& a, const Entry& b) { return a.image < b.image; });
    for (const Entry& e : f) ids.push_back(e.image);
  }
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  const int num_total = (int)ids.size();  // image_ids.size() of GetImageIds: images with at least one entry
  for (int w = 0; w < X.n_words; ++w) {
    auto& f = X.files[w];
    if (f.empty()) continue;
    int distinct = 0, last = -1;
    for (const Entry& e : f) if (e.image != last) { ++distinct; last = e.image; }
    X.idf[w] = (float)std::log((double)num_total / (double)distinct);
  }
  for (int w = 0; w < X.n_words; ++w) {
    const double sq = (double)(X.idf[w] * X.idf[w]);
    for (const Entry& e : X.files[w]) self[e.image] += sq;
  }
  X.norm.clear();
  for (const auto& s : self) X.norm[s.first] = s.second > 0.0 ? (float)(1.0 / std::sqrt(s.second)) : 0.0f;
}

// VisualIndex::Query without spatial verification: scores of all images hit, sorted by descending score (ties: lower
// image id first -- the reference leaves ties to std::sort), truncated to max_num_images (< 0: all).  Returns the count.
int orc_retrieval_query(void* p, int n, const uint8_t* desc, int num_neighbors, int max_num_images, int* out_ids, float* out_scores) {
  const Index& X = *(Index*)p;
  std::vector<int> wid((size_t)n * num_neighbors);
  orc_retrieval_word_ids(p, n, desc, num_neighbors, wid.data());
  double self = 0.0;
  for (int w : wid) if (w != 0x7fffffff) self += (double)(X.idf[w] * X.idf[w]);
  const float self_similarity = (float)self;
  const float normalization_weight = self_similarity > 0.0f ? 1.0f / std::sqrt(self_similarity) : 1.0f;
  std::unordered_map<int, int> slot;
  std::vector<std::pair<int, float>> scores;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < num_neighbors; ++k) {
      const int w = wid[(size_t)i * num_neighbors + k];
      if (w == 0x7fffffff) continue;
      const auto& f = X.files[w];
      if (!X.has_emb[w] || f.empty()) continue;  // IsUsable(): embedding learnt and entries sorted
      const float sq = X.idf[w] * X.idf[w];
      const uint64_t b = signature(X, w, desc + (size_t)i * kDim);
      int img = f.front().image, votes = 0;
      float sc = 0.0f;
      auto flush = [&]() {
        if (votes > 0) {
          float v = sc / std::sqrt((float)votes);
          v *= sq;
          auto it = slot.find(img);
          if (it == slot.end()) { slot.emplace(img, (int)scores.size()); scores.emplace_back(img, v); }
          else scores[it->second].second += v;
        }
      };
      for (const Entry& e : f) {
        if (img < e.image) { flush(); img = e.image; sc = 0.0f; votes = 0; }
        const int hd = __builtin_popcountll(b ^ e.bits);
        if (hd <= kMaxHamming) { sc += X.lut[hd]; votes += 1; }
      }
      flush();
    }
  }
  for (auto& s : scores) s.second *= normalization_weight * X.norm.at(s.first);
  std::stable_sort(scores.begin(), scores.end(), [](const std::pair<int, float>& a, const std::pair<int, float>& b) {
    return a.second > b.second || (a.second == b.second && a.first < b.first);
  });
  int m = (int)scores.size();
  if (max_num_images >= 0 && max_num_images < m) m = max_num_images;
  for (int k = 0; k < m; ++k) { out_ids[k] = scores[k].first; out_scores[k] = scores[k].second; }
  return m;
}

}  // extern "C"
