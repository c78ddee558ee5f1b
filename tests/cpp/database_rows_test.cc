// Reads a batch in the C ABI's layout from argv[1], writes the rows of include/dagsfm_b200/database_rows.hpp to
// argv[2].  Driven by tests/test_database_rows.py, which inserts the rows into a SQLite database with the
// reference's schema and reads them back the way Database::ReadMatches / ReadTwoViewGeometry do.
#include <cstdio>
#include <vector>

#include "dagsfm_b200/database_rows.hpp"

using namespace dagsfm_b200;

template <class T> static bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
template <class T> static void wr(FILE* f, const T* p, size_t n) { if (n) fwrite(p, sizeof(T), n, f); }
static void blob(FILE* f, const std::vector<uint8_t>& b) { const int64_t n = (int64_t)b.size(); wr(f, &n, 1); wr(f, b.data(), b.size()); }

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t n_pairs = 0, n_images = 0, min_inl = 0, has_pose = 0;
  if (!rd(f, &n_pairs, 1) || !rd(f, &n_images, 1) || !rd(f, &min_inl, 1) || !rd(f, &has_pose, 1)) return 2;
  std::vector<uint32_t> pairs(2 * n_pairs), ids(n_images);
  std::vector<int64_t> off(n_pairs + 1);
  if (!rd(f, pairs.data(), pairs.size()) || !rd(f, ids.data(), ids.size()) || !rd(f, off.data(), off.size())) return 2;
  const int64_t total = off[n_pairs];
  std::vector<uint32_t> matches(2 * total), inl(2 * total);
  std::vector<b2_two_view_result> res(n_pairs);
  std::vector<b2_relative_pose> poses(n_pairs);
  if (!rd(f, matches.data(), matches.size()) || !rd(f, inl.data(), inl.size()) || !rd(f, res.data(), res.size()) ||
      !rd(f, poses.data(), poses.size()))
    return 2;
  fclose(f);
  std::vector<MatchesRow> mrows;
  std::vector<TwoViewGeometryRow> grows;
  MakeRowsBatch(n_pairs, pairs.data(), ids.data(), off.data(), matches.data(), res.data(), inl.data(),
                has_pose ? poses.data() : nullptr, (int)min_inl, &mrows, &grows);
  FILE* o = fopen(argv[2], "wb");
  if (!o) return 2;
  for (int64_t p = 0; p < n_pairs; ++p) {
    wr(o, &mrows[p].pair_id, 1); wr(o, &mrows[p].rows, 1); wr(o, &mrows[p].cols, 1); blob(o, mrows[p].data);
    wr(o, &grows[p].pair_id, 1); wr(o, &grows[p].rows, 1); wr(o, &grows[p].cols, 1); blob(o, grows[p].data);
    wr(o, &grows[p].config, 1); blob(o, grows[p].F); blob(o, grows[p].E);
  }
  fclose(o);
  // pair-id helpers (database_test.cc: TestImagePairToPairId / TestSwapImagePair semantics)
  uint32_t a = 0, b = 0;
  PairIdToImagePair(ImagePairToPairId(7, 3), &a, &b);
  if (a != 3 || b != 7 || ImagePairToPairId(3, 7) != ImagePairToPairId(7, 3) || !SwapImagePair(7, 3) || SwapImagePair(3, 7)) return 3;
  const auto u = UniqueImagePairs({{1, 2}, {2, 1}, {3, 3}, {1, 3}, {1, 2}});
  if (u.size() != 2 || u[0] != std::make_pair(1u, 2u) || u[1] != std::make_pair(1u, 3u)) return 4;
  std::printf("rows ok\n");
  return 0;
}
