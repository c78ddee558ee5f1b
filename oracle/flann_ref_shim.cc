// Thin C entry points over the reference's VENDORED FLANN (lib/FLANN, compiled from its own sources where they lie under
// /root/reference -- nothing is copied) -- TEST INFRASTRUCTURE: the checker of the retrieval oracle's word search and the
// measure of how far the reference's approximate search is from the exact one.  Built into oracle/_ref/libflann_ref.so by
// oracle/Makefile.
//
//   flann_ref_knn_linear     the k nearest visual words by FLANN's own exhaustive index (flann::LinearIndex, flann::L2<uint8>,
//                            KNNResultSet): the reference's distance functor and result handling in EXACT mode.
//   flann_ref_knn_autotuned  what VisualIndex::Build + FindWordIds do (src/retrieval/visual_index.h:517-521, 701-744):
//                            flann::AutotunedIndex with target_precision (BuildOptions default 0.9... passed in), built on
//                            the words, knnSearch with SearchParams(num_checks).  AutotunedIndex picks its algorithm and
//                            parameters from TIMING experiments on this host and its trees are randomised: its answers
//                            are approximate and not reproducible -- which is why parity of the word search is defined
//                            against the exact answer.
#include <cstdint>
#include <vector>

#include "FLANN/flann.hpp"

extern "C" {

int flann_ref_knn_linear(const uint8_t* words, int n_words, const uint8_t* desc, int n, int k, int* out_ids, float* out_dist) {
  flann::Matrix<uint8_t> W(const_cast<uint8_t*>(words), n_words, 128), Q(const_cast<uint8_t*>(desc), n, 128);
  flann::Index<flann::L2<uint8_t>> index(W, flann::LinearIndexParams());
  index.buildIndex();
  std::vector<size_t> idx((size_t)n * k, (size_t)-1);
  std::vector<float> dist((size_t)n * k, 0.0f);
  flann::Matrix<size_t> I(idx.data(), n, k);
  flann::Matrix<float> D(dist.data(), n, k);
  index.knnSearch(Q, I, D, k, flann::SearchParams(flann::FLANN_CHECKS_UNLIMITED));
  for (size_t i = 0; i < idx.size(); ++i) {
    out_ids[i] = idx[i] == (size_t)-1 ? 0x7fffffff : (int)idx[i];
    if (out_dist) out_dist[i] = dist[i];
  }
  return 0;
}

int flann_ref_knn_autotuned(const uint8_t* words, int n_words, const uint8_t* desc, int n, int k, int num_checks,
                            float target_precision, int cores, int* out_ids) {
  flann::Matrix<uint8_t> W(const_cast<uint8_t*>(words), n_words, 128), Q(const_cast<uint8_t*>(desc), n, 128);
  flann::AutotunedIndexParams index_params;
  index_params["target_precision"] = target_precision;
  flann::AutotunedIndex<flann::L2<uint8_t>> index(index_params);
  index.buildIndex(W);
  std::vector<size_t> idx((size_t)n * k, (size_t)-1);
  std::vector<float> dist((size_t)n * k, 0.0f);
  flann::Matrix<size_t> I(idx.data(), n, k);
  flann::Matrix<float> D(dist.data(), n, k);
  flann::SearchParams sp(num_checks);
  sp.cores = cores > 0 ? cores : 1;
  index.knnSearch(Q, I, D, k, sp);
  for (size_t i = 0; i < idx.size(); ++i) out_ids[i] = idx[i] == (size_t)-1 ? 0x7fffffff : (int)idx[i];
  return 0;
}

}  // extern "C"
