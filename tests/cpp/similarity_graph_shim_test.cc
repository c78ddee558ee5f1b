// The retrieval seam through the C++ adaptor (include/dagsfm_b200/similarity_graph_shim.hpp): the structure the reference's own
// test pins (src/retrieval/visual_index_test.cc:84-112 -- an indexed image retrieved with its own descriptors ranks first with
// a strictly larger score, result sizes follow max_num_images) expressed on VocabSimilarityGraph::Run's outputs, plus the
// pair-list rules of similarity_graph.cpp:183-191 (image_id < other, score * 1e3, query order).  Built by
// tests/test_similarity_graph_shim.py against the CUDA-emulator build of retrieval.cu (CPU) and against the product library.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <vector>

#include "dagsfm_b200/similarity_graph_shim.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

static unsigned g_s = 12345u;
static unsigned rnd() { g_s = g_s * 1664525u + 1013904223u; return g_s >> 8; }

int main() {
  using namespace dagsfm_b200;
  // a small collection with structure: image i holds 40 descriptors around "scene" prototypes i*8 .. i*8+39 (neighbours share 32)
  const int n_img = 9, n_kp = 40, n_proto = 8 * n_img + 40;
  std::vector<std::vector<uint8_t>> proto(n_proto, std::vector<uint8_t>(128));
  for (auto& p : proto) for (auto& v : p) v = (uint8_t)(rnd() % 160);
  std::vector<std::vector<uint8_t>> desc(n_img, std::vector<uint8_t>((size_t)n_kp * 128));
  for (int i = 0; i < n_img; ++i)
    for (int k = 0; k < n_kp; ++k)
      for (int j = 0; j < 128; ++j) {
        int v = proto[8 * i + k][j] + (int)(rnd() % 7) - 3;
        desc[i][(size_t)k * 128 + j] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
  // vocabulary: one word per prototype, a projection of +-1 / sqrt(128) rows, thresholds at the projection of the mean descriptor
  // (80 per component): the sign pattern of proj (d - mean) is the prototype's, the +-3 noise of a view does not flip it
  VocabularyTree voc;
  voc.n_words = n_proto;
  voc.words.resize((size_t)n_proto * 128);
  for (int w = 0; w < n_proto; ++w) memcpy(&voc.words[(size_t)w * 128], proto[w].data(), 128);
  voc.proj.resize(64 * 128);
  for (auto& x : voc.proj) x = ((rnd() & 1) ? 1.0f : -1.0f) / std::sqrt(128.0f);
  voc.thresholds.resize((size_t)n_proto * 64);
  for (int w = 0; w < n_proto; ++w)
    for (int i = 0; i < 64; ++i) {
      float s = 0.0f;
      for (int j = 0; j < 128; ++j) s += voc.proj[(size_t)i * 128 + j] * 80.0f;
      voc.thresholds[(size_t)w * 64 + i] = s;
    }
  voc.has_embedding.assign(n_proto, 1);

  std::vector<image_t> ids;
  std::vector<const uint8_t*> ptrs;
  std::vector<int32_t> cnt;
  for (int i = 0; i < n_img; ++i) { ids.push_back(100 + 3 * i); ptrs.push_back(desc[i].data()); cnt.push_back(n_kp); }

  VocabSimilaritySearchOptions opt;
  CHECK(opt.num_images == 100 && opt.num_nearest_neighbors == 5 && opt.num_checks == 256 && opt.num_images_after_verification == 0);
  opt.num_images = 4;
  opt.num_nearest_neighbors = 1;   // every descriptor votes through the word of its own prototype only: no chance votes in so small a set
  VocabSimilarityGraph graph(opt);
  graph.Run(voc, ids, ptrs, cnt);
  const auto& pairs = graph.ImagePairs();
  const auto& scores = graph.Scores();
  CHECK(pairs.size() == scores.size() && !pairs.empty());
  std::set<std::pair<image_t, image_t>> seen;
  for (size_t k = 0; k < pairs.size(); ++k) {
    CHECK(pairs[k].first < pairs[k].second);            // similarity_graph.cpp:186
    CHECK(scores[k] > 0.0f);
    CHECK(k == 0 || pairs[k - 1].first <= pairs[k].first);   // query order
    seen.insert(pairs[k]);
  }
  for (int i = 0; i + 1 < n_img; ++i) CHECK(seen.count({ids[i], ids[i + 1]}) == 1);   // neighbours share 32 of 40 prototypes

  // visual_index_test.cc:84-112 through the C ABI the adaptor uses: self first, strictly ahead; sizes follow max_num_images
  b2_retrieval* r = nullptr;
  CHECK(b2_retrieval_create(0, &r) == B2_OK);
  CHECK(b2_retrieval_set_vocabulary(r, voc.n_words, voc.words.data(), voc.proj.data(), voc.thresholds.data(), voc.has_embedding.data()) == B2_OK);
  std::vector<uint8_t> two(desc[0]);
  two.insert(two.end(), desc[5].begin(), desc[5].end());
  const int64_t off[3] = {0, n_kp, 2 * n_kp};
  CHECK(b2_retrieval_index_images(r, 2, two.data(), off, 5) == B2_OK);
  int32_t out_ids[6], out_cnt[2];
  float out_sc[6];
  CHECK(b2_retrieval_query_all(r, 1, out_ids, out_sc, out_cnt) == B2_OK);
  CHECK(out_cnt[0] == 1 && out_cnt[1] == 1 && out_ids[0] == 0 && out_ids[1] == 1);
  CHECK(b2_retrieval_query_all(r, 3, out_ids, out_sc, out_cnt) == B2_OK);
  CHECK(out_cnt[0] >= 1 && out_cnt[0] <= 2 && out_ids[0] == 0 && out_ids[3] == 1);
  if (out_cnt[0] == 2) CHECK(out_sc[0] > out_sc[1]);
  CHECK(b2_retrieval_query_all(r, 0, out_ids, out_sc, out_cnt) != B2_OK);   // CHECK_GT(max_num_images, 0)
  b2_retrieval_destroy(r);

  bool threw = false;
  try { VocabSimilaritySearchOptions bad; bad.num_images = 0; VocabSimilarityGraph g(bad); } catch (const std::invalid_argument&) { threw = true; }
  CHECK(threw);
  std::printf("similarity graph shim ok: %zu pairs\n", pairs.size());
  return 0;
}
