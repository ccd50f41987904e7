// Host-only harness around pvr_habitat_b200/csrc/preprocess_aa_core.cuh (test infrastructure): the same
// __host__ __device__ functions the CUDA kernel uses, driven on the CPU so that tests/test_preprocess_aa_core.py can
// compare them bit for bit with the oracle without a GPU. Built by the test with g++ -ffp-contract=off.
#include <cstdint>
#include <vector>

#include "preprocess_aa_core.cuh"

extern "C" {

int aa_max_taps(void) { return pvr::AA_MAX_TAPS; }

// per output index: xmin[i], size[i], w[i * AA_MAX_TAPS + j]
void aa_weights(int in_size, int out_size, int* xmin, int* size, float* w) {
  for (int i = 0; i < out_size; ++i) pvr::aa_index_weights(i, in_size, out_size, &xmin[i], &size[i], w + i * pvr::AA_MAX_TAPS);
}

// (planes, H, W) uint8 -> (planes, rh, rw) float32: horizontal pass into a float32 intermediate, then vertical pass
void aa_resize(const uint8_t* in, int planes, int H, int W, int rh, int rw, float* out) {
  std::vector<int> xmin(rw), xsz(rw), ymin(rh), ysz(rh);
  std::vector<float> wx((size_t)rw * pvr::AA_MAX_TAPS), wy((size_t)rh * pvr::AA_MAX_TAPS);
  aa_weights(W, rw, xmin.data(), xsz.data(), wx.data());
  aa_weights(H, rh, ymin.data(), ysz.data(), wy.data());
  std::vector<float> tmp((size_t)H * rw);
  for (int p = 0; p < planes; ++p) {
    const uint8_t* src = in + (size_t)p * H * W;
    for (int r = 0; r < H; ++r)
      for (int x = 0; x < rw; ++x) {
        const uint8_t* s = src + (size_t)r * W + xmin[x];
        tmp[(size_t)r * rw + x] =
            pvr::aa_accumulate(xsz[x], wx.data() + (size_t)x * pvr::AA_MAX_TAPS, [&](int j) { return (float)s[j]; });
      }
    for (int y = 0; y < rh; ++y)
      for (int x = 0; x < rw; ++x) {
        const float* t = tmp.data() + (size_t)ymin[y] * rw + x;
        out[((size_t)p * rh + y) * rw + x] = pvr::aa_accumulate(
            ysz[y], wy.data() + (size_t)y * pvr::AA_MAX_TAPS, [&](int j) { return t[(size_t)j * rw]; });
      }
  }
}

}  // extern "C"
