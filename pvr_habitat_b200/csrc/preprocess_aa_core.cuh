// Arithmetic core of the antialiased bicubic resize (CLIP transforms, src/embeddings.py:309-310:
// T.Resize(res, BICUBIC, antialias=True) -> ATen's separable _upsample_bicubic2d_aa on the float image, a = -0.5).
//
// Everything that decides the bits of the result lives here as __host__ __device__ functions, so that the same code is
// compiled into the CUDA kernel (preprocess_aa.cu) and into a host-only harness that tests/test_preprocess_aa_core.py
// checks bit for bit against the oracle (oracle/restate.py:_aa_weights / _aa_apply) without a GPU. Which operations
// are fused and which sub-expressions run in double follows the x86 build of ATen (probed, see the oracle).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDA_ARCH__
#define PVR_AA_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define PVR_AA_MUL(a, b) __fmul_rn((a), (b))
#define PVR_AA_ADD(a, b) __fadd_rn((a), (b))
#define PVR_AA_SUB(a, b) __fsub_rn((a), (b))
#define PVR_AA_DIV(a, b) __fdiv_rn((a), (b))
#define PVR_AA_DMUL(a, b) __dmul_rn((a), (b))
#define PVR_AA_DADD(a, b) __dadd_rn((a), (b))
#define PVR_AA_DDIV(a, b) __ddiv_rn((a), (b))
#else  // host harness: compiled with -ffp-contract=off, fmaf is the correctly rounded libm function
#define PVR_AA_FMA(a, b, c) std::fmaf((a), (b), (c))
#define PVR_AA_MUL(a, b) ((a) * (b))
#define PVR_AA_ADD(a, b) ((a) + (b))
#define PVR_AA_SUB(a, b) ((a) - (b))
#define PVR_AA_DIV(a, b) ((a) / (b))
#define PVR_AA_DMUL(a, b) ((a) * (b))
#define PVR_AA_DADD(a, b) ((a) + (b))
#define PVR_AA_DDIV(a, b) ((a) / (b))
#endif

#ifndef PVR_HD
#ifdef __CUDACC__
#define PVR_HD __host__ __device__ __forceinline__
#else
#define PVR_HD inline
#endif
#endif

namespace pvr {

constexpr int AA_MAX_TAPS = 32;  // ceil(2 * scale) * 2 + 1 taps: down-scaling by up to 7.5x

// ATen aa_filter for bicubic, a = -0.5 (every a*b+c fused)
PVR_HD float aa_cubic_filter(float x) {
  x = fabsf(x);
  if (x < 1.f) {
    const float t = PVR_AA_FMA(1.5f, x, -2.5f);
    return PVR_AA_FMA(PVR_AA_MUL(t, x), x, 1.f);
  }
  if (x < 2.f) {
    float u = PVR_AA_FMA(PVR_AA_SUB(x, 5.f), x, 8.f);
    u = PVR_AA_FMA(u, x, -4.f);
    return PVR_AA_MUL(u, -0.5f);
  }
  return 0.f;
}

// ATen _compute_indices_min_size_weights_aa for output index i of one dimension: first input index, number of taps and
// the normalised weights w[0 .. size). Float variables meet double literals in the C++ (`+ 0.5`, `1.0 / scale`): those
// sub-expressions are evaluated in double and rounded once.
PVR_HD void aa_index_weights(int i, int in_size, int out_size, int* xmin_out, int* size_out, float* w) {
  const float scale = PVR_AA_DIV((float)in_size, (float)out_size);
  const bool down = scale >= 1.f;
  const float support = down ? PVR_AA_MUL(2.f, scale) : 2.f;
  const float invscale = down ? (float)PVR_AA_DDIV(1.0, (double)scale) : 1.f;
  const float center = (float)PVR_AA_DMUL((double)scale, PVR_AA_DADD((double)i, 0.5));
  long long lo = (long long)PVR_AA_DADD((double)PVR_AA_SUB(center, support), 0.5);
  if (lo < 0) lo = 0;
  long long hi = (long long)PVR_AA_DADD((double)PVR_AA_ADD(center, support), 0.5);
  if (hi > in_size) hi = in_size;
  int size = (int)(hi - lo);
  if (size < 0) size = 0;
  if (size > AA_MAX_TAPS) size = AA_MAX_TAPS;
  float total = 0.f;
  for (int j = 0; j < size; ++j) {
    const float d = PVR_AA_SUB((float)(j + lo), center);
    const float arg = (float)PVR_AA_DMUL(PVR_AA_DADD((double)d, 0.5), (double)invscale);
    w[j] = aa_cubic_filter(arg);
    total = PVR_AA_ADD(total, w[j]);
  }
  if (total != 0.f)
    for (int j = 0; j < size; ++j) w[j] = PVR_AA_DIV(w[j], total);
  *xmin_out = (int)lo;
  *size_out = size;
}

// ATen interpolate_aa_single_dim: out = v[0] * w[0]; out += v[j] * w[j]. As compiled for x86 the loop runs in groups of
// four iterations with a rounded product and a separate add; the remaining (n - 1) mod 4 iterations are fused.
// `get(j)` returns the j-th input value as float.
template <class Get>
PVR_HD float aa_accumulate(int n, const float* w, Get get) {
  float o = PVR_AA_MUL(get(0), w[0]);
  const int grouped = (n - 1) / 4 * 4;
  int j = 1;
  for (; j <= grouped; ++j) o = PVR_AA_ADD(o, PVR_AA_MUL(get(j), w[j]));
  for (; j < n; ++j) o = PVR_AA_FMA(get(j), w[j], o);
  return o;
}

// ---- the kernel's work decomposition, also host/device so that the harness runs the very same index arithmetic
struct AAGeom {
  int N, H, W, CH, nf;              // observations (N, H, W, CH = 3 * nf) uint8
  int top, left, crop, rows, bands; // crop window inside the resized image; output rows per band; bands per image
  const int *ymin, *ysize, *xmin, *xsize;  // tap ranges per resized row / column
  const float *wy, *wx;                    // weights, AA_MAX_TAPS per row / column
  int sample_major;                        // image index of (sample i, frame f): i*nf + f instead of f*N + i
};

struct AABand {
  int img, y_first, y_count;  // observation, first output row (inside the crop) and number of rows of this band
  int r_lo, rows_in;          // input rows the band depends on: [r_lo, r_lo + rows_in)
};

PVR_HD AABand aa_band(const AAGeom& g, int block) {
  AABand b;
  b.img = block / g.bands;
  const int band = block - b.img * g.bands;
  b.y_first = band * g.rows;
  b.y_count = g.crop - b.y_first < g.rows ? g.crop - b.y_first : g.rows;
  const int y_last = b.y_first + b.y_count - 1 + g.top;
  b.r_lo = g.ymin[b.y_first + g.top];  // tap ranges move monotonically with the output row
  b.rows_in = g.ymin[y_last] + g.ysize[y_last] - b.r_lo;
  return b;
}

// ((u / 255) - mean) / std for every uint8 value, each operation rounded to float32 like the reference
PVR_HD void aa_build_lut(float* lut, const float* mean, const float* stdv, int tid, int nthreads) {
  for (int t = tid; t < 768; t += nthreads) {
    const int c = t >> 8, u = t & 255;
    lut[t] = PVR_AA_DIV(PVR_AA_SUB(PVR_AA_DIV((float)u, 255.0f), mean[c]), stdv[c]);
  }
}

// horizontal pass of frame f: tmp[r][x][c] for the band's input rows (s = first byte of input row r_lo) and the crop's
// columns
PVR_HD void aa_horizontal(const AAGeom& g, const AABand& b, const uint8_t* s, float* tmp, int f, int tid, int nthreads) {
  const long long row_bytes = (long long)g.W * g.CH;
  const int row_vals = g.crop * 3;
  const int CH = g.CH;
  for (int idx = tid; idx < b.rows_in * row_vals; idx += nthreads) {
    const int r = idx / row_vals;
    const int rem = idx - r * row_vals;
    const int x = rem / 3;
    const int c = rem - x * 3;
    const int X = x + g.left;
    const uint8_t* src = s + (long long)r * row_bytes + (long long)g.xmin[X] * CH + 3 * f + c;
    tmp[idx] = aa_accumulate(g.xsize[X], g.wx + (long long)X * AA_MAX_TAPS, [&](int j) { return (float)src[j * CH]; });
  }
}

// vertical pass of frame f + clamp + half-even round + normalisation table; store(image, y, x, o[3]) writes the pixel
template <class Store>
PVR_HD void aa_vertical(const AAGeom& g, const AABand& b, const float* tmp, const float* lut, int f, int tid,
                        int nthreads, Store store) {
  const int row_vals = g.crop * 3;
  const long long image = g.sample_major ? (long long)b.img * g.nf + f : (long long)f * g.N + b.img;
  for (int idx = tid; idx < b.y_count * g.crop; idx += nthreads) {
    const int yy = idx / g.crop;
    const int x = idx - yy * g.crop;
    const int y = b.y_first + yy;
    const int Y = y + g.top;
    const float* w = g.wy + (long long)Y * AA_MAX_TAPS;
    const int n = g.ysize[Y];
    const float* t0 = tmp + ((long long)(g.ymin[Y] - b.r_lo) * g.crop + x) * 3;
    float o[3];
    for (int c = 0; c < 3; ++c) {
      float v = aa_accumulate(n, w, [&](int j) { return t0[(long long)j * row_vals + c]; });
      v = fminf(fmaxf(v, 0.f), 255.f);  // torchvision clamps the overshoot before the rounding cast
      o[c] = lut[c * 256 + (int)rintf(v)];
    }
    store(image, y, x, o);
  }
}

}  // namespace pvr
