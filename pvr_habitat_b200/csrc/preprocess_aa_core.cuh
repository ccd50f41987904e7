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

}  // namespace pvr
