// C++ adaptor of the retrieval seam: VocabSimilaritySearchOptions and the retrieval loop of VocabSimilarityGraph::Run
// (src/graph/similarity_graph.h:41-75, similarity_graph.cpp:101-200) over the C ABI (b2_retrieval_*).  The reference's Run()
// reads the vocabulary tree and the images' descriptors from disk / the database cache and fills image_pairs_ and scores_;
// here the caller hands over what those reads leave in memory and gets the same two vectors.
// Header-only; no Eigen, no FLANN.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../dagsfm_b200.h"

namespace dagsfm_b200 {

typedef uint32_t image_t;

struct VocabSimilaritySearchOptions {   // similarity_graph.h:41-66, same members and defaults
  int num_images = 100;
  int num_nearest_neighbors = 5;
  int num_checks = 256;                   // FLANN's leaf budget; the word search here is exact
  int num_images_after_verification = 0;  // spatial re-ranking: not part of this seam, must stay 0
  int max_num_features = -1;              // applied by the caller (ExtractTopScaleFeatures) before handing descriptors over
  int num_threads = 8;
  std::string vocab_tree_path = "";
  int gpu_index = 0;
  void Check() const {                    // similarity_graph.cpp: CHECK_GT(num_images, 0) etc.
    if (num_images <= 0 || num_nearest_neighbors <= 0 || num_nearest_neighbors > 8 || num_images_after_verification != 0)
      throw std::invalid_argument("VocabSimilaritySearchOptions::Check failed");
  }
};

// What retrieval::VisualIndex<>::Read leaves in memory (visual_index.h:541-600): visual_words_, the inverted index's
// proj_matrix_ (64 x 128 row-major) and, per inverted file, thresholds_ and (status_ & HAS_EMBEDDING).
struct VocabularyTree {
  int32_t n_words = 0;
  std::vector<uint8_t> words;          // [n_words * 128]
  std::vector<float> proj;             // [64 * 128]
  std::vector<float> thresholds;       // [n_words * 64]
  std::vector<uint8_t> has_embedding;  // [n_words]
};

class VocabSimilarityGraph {
 public:
  explicit VocabSimilarityGraph(const VocabSimilaritySearchOptions& options) : options_(options) { options_.Check(); }

  // image_ids[i] owns descriptors[i] (n_desc[i] x 128 uint8, row-major).  Fills ImagePairs() / Scores() exactly as
  // similarity_graph.cpp:183-191 does: (image_id, other) with image_id < other, score * 1e3, in query order.
  void Run(const VocabularyTree& vocab, const std::vector<image_t>& image_ids, const std::vector<const uint8_t*>& descriptors,
           const std::vector<int32_t>& n_desc) {
    image_pairs_.clear();
    scores_.clear();
    const size_t n = image_ids.size();
    if (descriptors.size() != n || n_desc.size() != n) throw std::invalid_argument("one descriptor array per image");
    b2_retrieval* r = nullptr;
    if (b2_retrieval_create(options_.gpu_index, &r) != B2_OK) throw std::runtime_error(b2_last_error());
    struct Guard { b2_retrieval* r; ~Guard() { b2_retrieval_destroy(r); } } guard{r};
    if (b2_retrieval_set_vocabulary(r, vocab.n_words, vocab.words.data(), vocab.proj.data(), vocab.thresholds.data(),
                                    vocab.has_embedding.data()) != B2_OK)
      throw std::runtime_error(b2_last_error());
    std::vector<int64_t> off(n + 1, 0);
    for (size_t i = 0; i < n; ++i) off[i + 1] = off[i] + n_desc[i];
    std::vector<uint8_t> all((size_t)off[n] * 128);
    for (size_t i = 0; i < n; ++i)
      if (n_desc[i] > 0) std::copy(descriptors[i], descriptors[i] + (size_t)n_desc[i] * 128, all.begin() + (size_t)off[i] * 128);
    if (b2_retrieval_index_images(r, (int32_t)n, all.data(), off.data(), options_.num_nearest_neighbors) != B2_OK)
      throw std::runtime_error(b2_last_error());
    const int K = options_.num_images;
    std::vector<int32_t> ids(n * (size_t)K), cnt(n);
    std::vector<float> sc(n * (size_t)K);
    if (n > 0 && b2_retrieval_query_all(r, K, ids.data(), sc.data(), cnt.data()) != B2_OK) throw std::runtime_error(b2_last_error());
    for (size_t q = 0; q < n; ++q)
      for (int k = 0; k < cnt[q]; ++k) {
        const image_t other = image_ids[(size_t)ids[q * K + k]];
        if (image_ids[q] < other) {
          image_pairs_.emplace_back(image_ids[q], other);
          scores_.push_back(sc[q * K + k] * 1e3f);
        }
      }
  }

  const std::vector<std::pair<image_t, image_t>>& ImagePairs() const { return image_pairs_; }
  const std::vector<float>& Scores() const { return scores_; }

 private:
  VocabSimilaritySearchOptions options_;
  std::vector<std::pair<image_t, image_t>> image_pairs_;
  std::vector<float> scores_;
};

}  // namespace dagsfm_b200
