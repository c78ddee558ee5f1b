// Shared device-side structures of the descriptor matcher.
//
// HBM layout
//   pool            uint8 [pool_rows][128]   all images' descriptors; image i owns rows
//                                            [img_row[i], img_row[i] + pad256(n_desc[i])),
//                                            rows >= n_desc[i] are zero (a zero row has dot 0,
//                                            which can never be a best/second-best: the reference
//                                            starts both at 0 and compares with strict '>',
//                                            sift.cc:121-133)
//   items           MatchItem [n_items]      one 256-row "supertile" of the X image against the
//                                            whole Y image, per direction of a pair
//   midx            int32 [n_items*256]      per X row: matched Y row or -1
//   cands           uint4 [...]              rows that passed the integer threshold tests and need
//                                            the exact in-chunk rescan (match_fixup kernel)
#pragma once
#include <cstdint>

namespace b2 {

constexpr int kDescBytes = 128;   // FeatureDescriptors column count (types.h:102)
constexpr int kSuperRows = 256;   // X rows per work item (two M=128 accumulators)
constexpr int kTileRows = 128;    // UMMA M and N
constexpr int kChunk = 32;        // columns whose max is tracked as one unit
constexpr int kDotClamp = 262144; // 512*512: kDistNorm * dot >= 1 from here on (sift.cc:115)

struct MatchItem {
  uint32_t x_row;   // pool row of the first X row of this supertile
  uint32_t y_row;   // pool row of the first row of the Y image
  uint32_t y_nblk;  // number of 128-row blocks of the Y image (padded)
  uint32_t pad;
};

struct PairMeta {
  uint32_t item_start;  // first item of this pair (direction 0 supertiles, then direction 1)
  uint32_t nt1;         // supertiles of image 1 (direction 0 items)
  uint32_t n1, n2;      // descriptor counts
};

}  // namespace b2
