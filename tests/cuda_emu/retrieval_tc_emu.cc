// Host stand-in for the tcgen05 word-search kernel (retrieval_tc.cu), which has no CPU meaning: it produces what that
// kernel is CONTRACTED to produce -- per descriptor the k words with the largest 2 d.w - |w|^2 (word_sq as given, padding
// rows carry INT_MAX), ties to the lower word id, kInvalidWord where fewer than k words exist -- so that the rest of
// retrieval.cu (real code) runs against the oracle on the CPU.  TEST INFRASTRUCTURE ONLY.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace b2 {

cudaError_t launch_word_knn_tc(const CUtensorMap& tm, const uint8_t* desc, long long n_desc, const int* word_sq, uint32_t n_blk,
                               int k, int32_t* out, int, cudaStream_t) {
  cuda_emu::DeviceWindow window;
  const uint8_t* words = reinterpret_cast<const uint8_t*>(tm.opaque[0]);   // the emulator's tensor map keeps its base address
  for (long long i = 0; i < n_desc; ++i) {
    int bd[8], bw[8];
    for (int j = 0; j < k; ++j) { bd[j] = -0x7fffffff; bw[j] = 0x7fffffff; }
    for (uint32_t w = 0; w < n_blk * 128u; ++w) {
      int dot = 0;
      for (int j = 0; j < 128; ++j) dot += (int)desc[i * 128 + j] * (int)words[(size_t)w * 128 + j];
      const int s = 2 * dot - word_sq[w];
      if (s > bd[k - 1]) {
        int cd = s, cw = (int)w;
        bool placed = false;
        for (int j = 0; j < k; ++j)
          if (placed || cd > bd[j]) { const int td = bd[j], tw = bw[j]; bd[j] = cd; bw[j] = cw; cd = td; cw = tw; placed = true; }
      }
    }
    for (int j = 0; j < k; ++j) out[i * k + j] = bw[j];
  }
  return cudaSuccess;
}

}  // namespace b2
