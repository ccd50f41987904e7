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

// The kernel's own decomposition (aa_band / aa_build_lut / aa_horizontal / aa_vertical of preprocess_aa_core.cuh) run
// block by block with a single "thread": (N, H, W, 3*nf) uint8 observations -> (nf*N, 3, crop, crop) float32, the
// PVR_FMT_NCHW_F32 output of pvr_preprocess_u8_aa. The shared-memory staging is replaced by a pointer into the input.
void aa_preprocess(const uint8_t* in, int N, int H, int W, int nf, int rh, int rw, int top, int left, int crop,
                   const float* mean, const float* stdv, float* out, int sample_major, int rows) {
  std::vector<int> xmin(rw), xsz(rw), ymin(rh), ysz(rh);
  std::vector<float> wx((size_t)rw * pvr::AA_MAX_TAPS), wy((size_t)rh * pvr::AA_MAX_TAPS);
  aa_weights(W, rw, xmin.data(), xsz.data(), wx.data());
  aa_weights(H, rh, ymin.data(), ysz.data(), wy.data());
  pvr::AAGeom g;
  g.N = N; g.H = H; g.W = W; g.nf = nf; g.CH = 3 * nf;
  g.top = top; g.left = left; g.crop = crop; g.rows = rows; g.bands = (crop + rows - 1) / rows;
  g.ymin = ymin.data(); g.ysize = ysz.data(); g.xmin = xmin.data(); g.xsize = xsz.data();
  g.wy = wy.data(); g.wx = wx.data();
  g.sample_major = sample_major;
  float lut[768];
  pvr::aa_build_lut(lut, mean, stdv, 0, 1);
  const long long row_bytes = (long long)W * g.CH, plane = (long long)crop * crop;
  std::vector<float> tmp;
  for (int block = 0; block < g.bands * N; ++block) {
    const pvr::AABand b = pvr::aa_band(g, block);
    tmp.assign((size_t)b.rows_in * crop * 3, -1.f);
    const uint8_t* s = in + ((long long)b.img * H + b.r_lo) * row_bytes;
    for (int f = 0; f < nf; ++f) {
      pvr::aa_horizontal(g, b, s, tmp.data(), f, 0, 1);
      pvr::aa_vertical(g, b, tmp.data(), lut, f, 0, 1, [&](long long image, int y, int x, const float* o) {
        float* dst = out + image * 3 * plane + (long long)y * crop + x;
        dst[0] = o[0];
        dst[plane] = o[1];
        dst[2 * plane] = o[2];
      });
    }
  }
}

// upper bound the host side of pvr_preprocess_u8_aa uses to size shared memory; must cover every band's rows_in
int aa_max_rows_in(int H, int rh, int top, int crop, int rows) {
  std::vector<int> ymin(rh), ysz(rh);
  std::vector<float> wy((size_t)rh * pvr::AA_MAX_TAPS);
  aa_weights(H, rh, ymin.data(), ysz.data(), wy.data());
  int worst = 0;
  for (int y0 = 0; y0 < crop; y0 += rows) {
    const int y1 = (y0 + rows < crop ? y0 + rows : crop) - 1 + top;
    const int n = ymin[y1] + ysz[y1] - ymin[y0 + top];
    if (n > worst) worst = n;
  }
  return worst;
}

}  // extern "C"
