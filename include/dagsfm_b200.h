/*
 * dagsfm_b200 -- C ABI of the B200-native hot path of AIBluefisher/DAGSfM.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and returns an
 * int status (B2_OK == 0).  Nothing throws across this boundary and no torch /
 * Eigen / STL type appears in a signature.  Each block cites the reference
 * interface (path:line under the reference tree) it replaces.
 *
 * There is NO CPU fallback behind these calls: if no CUDA device / sm_100a
 * kernel image is available they return B2_ERR_CUDA / B2_ERR_NO_DEVICE.
 */
#ifndef DAGSFM_B200_H_
#define DAGSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
  B2_OK = 0,
  B2_ERR_INVALID = 1,   /* bad argument (the reference CHECK()-aborts here)     */
  B2_ERR_CUDA = 2,      /* a CUDA call failed; b2_last_error() has the text     */
  B2_ERR_NO_DEVICE = 3, /* no sm_100 device visible                            */
  B2_ERR_CAPACITY = 4,  /* caller's output buffer too small; nothing written    */
  B2_ERR_INTERNAL = 5   /* device-side consistency check tripped                */
};

/* Text of the last error raised on the calling thread ("" if none). */
const char* b2_last_error(void);
/* Library version, e.g. "dagsfm_b200 0.1 (sm_100a)". */
const char* b2_version(void);
/* Number of CUDA kernels this library has launched since load (all threads);
 * bench.py reports the delta over the timed region as `gpu_launches`. */
uint64_t b2_kernel_launch_count(void);

/* ================================================================== MATCH ==
 * Replaces: MatchSiftFeaturesGPU (src/feature/sift.h:235-239,
 * src/feature/sift.cc:941-985) and the SiftMatchGPU object behind it
 * (lib/SiftGPU/SiftGPU.h:276-373: SetDescriptors / GetSiftMatch), with the
 * *semantics of the CPU path* MatchSiftFeaturesCPU (sift.cc:76-198,810-822),
 * which is what the reference's own test treats as the truth
 * (src/feature/sift_test.cc:448-578).
 *
 * Descriptors: row-major uint8, 128 bytes per feature (FeatureDescriptors,
 * src/feature/types.h:102-103).  Matches: {uint32 idx1, uint32 idx2}
 * (FeatureMatch, types.h:86-98), ascending idx1.
 */
typedef struct b2_matcher b2_matcher;

typedef struct b2_match_options {
  float max_ratio;        /* SiftMatchingOptions::max_ratio    (sift.h:131) default 0.8 */
  float max_distance;     /* SiftMatchingOptions::max_distance (sift.h:134) default 0.7 */
  int32_t cross_check;    /* SiftMatchingOptions::cross_check  (sift.h:137) default 1   */
  int32_t max_num_matches;/* SiftMatchingOptions::max_num_matches (sift.h:140) 32768:
                             features beyond it are clamped (SiftMatchCU.cpp:108) */
} b2_match_options;

/* Fills the reference defaults (src/feature/sift.h:116-165). */
void b2_match_default_options(b2_match_options* opt);

/* CreateSiftGPUMatcher (sift.cc:877-939).  `device` = CUDA ordinal. */
int b2_match_create(int device, b2_matcher** out);
int b2_match_destroy(b2_matcher* m);

/* -- image store: the batched seam ----------------------------------------
 * Uploads the descriptors of `n_images` images into one HBM pool (each image
 * zero-padded to a multiple of 256 rows).  `desc[i]` -> n_desc[i] x 128 bytes
 * in HOST memory.  Replaces the per-pair SetDescriptors H2D copies
 * (SiftMatchCU.cpp:99-112) and the FeatureMatcherCache::GetDescriptors calls
 * of SiftGPUFeatureMatcher::Run (src/feature/matching.cc:376-427). */
int b2_match_set_images(b2_matcher* m, int32_t n_images,
                        const uint8_t* const* desc, const int32_t* n_desc);
/* Same, from one contiguous DEVICE buffer: image i occupies rows
 * [row_offset[i], row_offset[i]+n_desc[i]) of `desc_dev` (row = 128 bytes). */
int b2_match_set_images_device(b2_matcher* m, int32_t n_images,
                               const uint8_t* desc_dev,
                               const int64_t* row_offset, const int32_t* n_desc);

/* Matches `n_pairs` image pairs (indices into the image store), all on the
 * device, and copies the result to HOST memory:
 *   out_offsets[n_pairs+1]  prefix offsets into out_matches (in matches)
 *   out_matches[capacity]   {idx1, idx2}
 * Returns B2_ERR_CAPACITY (and the needed count in *n_total) if capacity is
 * too small.  Replaces the loop of SiftGPUFeatureMatcher::Run over
 * MatchSiftFeaturesGPU (matching.cc:392-424). */
int b2_match_pairs(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs /*[n][2]*/,
                   const b2_match_options* opt, int64_t* out_offsets,
                   uint32_t* out_matches /*[capacity][2]*/, int64_t capacity,
                   int64_t* n_total);

/* Device-resident variant (inputs already in HBM, results stay in HBM): only
 * the total match count and the per-pair counts' checksum come back.  Used by
 * the bench's `value` leg and by callers that chain into b2_verify_*.
 * `pairs_dev` is a DEVICE pointer to n_pairs x 2 uint32.  `out_offsets_dev`
 * (n_pairs+1 int64) and `out_matches_dev` (capacity x 2 uint32) are DEVICE
 * buffers owned by the caller. */
int b2_match_pairs_device(b2_matcher* m, int64_t n_pairs, const uint32_t* pairs_dev,
                          const b2_match_options* opt, int64_t* out_offsets_dev,
                          uint32_t* out_matches_dev, int64_t capacity,
                          int64_t* n_total);

/* -- two-slot seam: SiftMatchGPU::SetDescriptors / GetSiftMatch ------------
 * (lib/SiftGPU/SiftGPU.h:312-326).  slot in {0,1}; `desc == NULL` keeps the
 * previous upload (sift.h:232-234).  b2_match_run returns the number of
 * matches in *n_out (<= max_num_matches) -- the reference returns -1 on a
 * device error (SiftMatchCU.cpp:193-196); here that is a non-zero status. */
int b2_match_set_descriptors(b2_matcher* m, int slot, int32_t n, const uint8_t* desc);
int b2_match_run(b2_matcher* m, const b2_match_options* opt,
                 uint32_t* out_matches /*[max_num_matches][2]*/, int32_t* n_out);

/* Test / profiling hooks (no reference counterpart). */
/* Seconds of device time (CUDA events on the matcher's stream) spent in the
 * tensor-core kernel and in all kernels during the last b2_match_pairs* call,
 * and the number of tensor-core launches. */
int b2_match_last_timing(b2_matcher* m, double* tc_kernel_s, double* all_kernels_s,
                         int64_t* tc_launches, int64_t* fixup_candidates);

#ifdef __cplusplus
}
#endif
#endif /* DAGSFM_B200_H_ */
