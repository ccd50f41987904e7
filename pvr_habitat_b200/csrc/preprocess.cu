// K1 — fused uint8 HWC decode + Resize(bilinear, half-even round back to uint8) + CenterCrop + /255 + Normalize
//      + frame split, one pass over HBM.
//
// Reference semantics: src/embeddings.py:80-85 (Resize(256) -> CenterCrop(224) -> ConvertImageDtype(float) ->
// Normalize) applied after the NHWC->NCHW transpose of src/embeddings.py:391-394, on each 3-channel frame of an
// (N, H, W, 3n) observation (frame split: main_bc_1.py:134, behavioral_cloning/save_embedded_obs.py:153).
// torchvision resizes the uint8 image by casting to fp32, bilinear interpolation (align_corners=False), torch.round
// (half to even) and a cast back to uint8; then x/255, (x - mean)/std as three separately rounded fp32 operations.
// Because the intermediate is a uint8 value, the last three operations are a 3 x 256 table built with exactly those
// fp32 operations (no FMA contraction), so the fp32 output is bit-identical.
//
// Data movement: one CTA produces `rows` output rows of all n frames of one observation. The input rows it needs are
// one contiguous byte range of the HWC image, staged into shared memory by a single 1-D bulk async copy (TMA engine,
// UBLKCP) that completes on an mbarrier; every input byte is read from HBM once per band (bands overlap by <= 2 rows).
// Output stores are coalesced along x (fp32 NCHW) or 8-byte pixels (bf16 NHWC4).
#include "pvr_b200.h"
#include "ptx.cuh"

namespace pvr {

struct PreParams {
  const uint8_t* in;
  void* out;
  long long total_bytes;  // N*H*W*CH
  int N, H, W, CH, nf;
  int top, left, crop, rows, bands;
  float scale_y, scale_x;
  float mean[3], stdv[3];
  int fmt;
  int sample_major;  // image index of (sample i, frame f): i*nf + f instead of f*N + i
  int pix_off;       // byte offset of the bf16 pixel buffer inside dynamic smem (PVR_FMT_STEM_BF16)
  int flt;           // PVR_RESIZE_FLOAT: no rounding back to uint8, no /255: (v - mean) / std on the fp32 interpolant
  int swap02;        // PVR_SWAP_ROWS_0_2: image rows 0 and 2 trade places before the resize
};

// bilinear interpolant v of channel c -> output value
__device__ __forceinline__ float finish_bilinear(const PreParams& p, float v, int c, const float* lut) {
  if (p.flt) return __fdiv_rn(__fsub_rn(v, p.mean[c]), p.stdv[c]);  // Normalize: sub_(mean).div_(std)
  int u = (int)rintf(v);  // half to even, as torch.round
  u = min(max(u, 0), 255);
  return lut[c * 256 + u];
}
__device__ __forceinline__ int swap_row(const PreParams& p, int r) {
  return p.swap02 ? (r == 0 ? 2 : (r == 2 ? 0 : r)) : r;
}

__device__ __forceinline__ void src_index(float scale, int dst, int size, int& i0, int& i1, float& l) {
  // ATen area_pixel_compute_source_index(align_corners=false) + guard_index_and_lambda. The x86 build of ATen
  // contracts scale*(dst+0.5)-0.5 into one FMA (probed bit-exact, oracle/restate.py:_src_index).
  float s = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > size - 1) i0 = size - 1;
  i1 = i0 + (i0 < size - 1 ? 1 : 0);
  l = __fsub_rn(s, (float)i0);
  l = fminf(fmaxf(l, 0.f), 1.f);
}

// Bicubic (A = -0.75, align_corners=False, no antialias): T.Resize(256, interpolation=3) of the MAE encoders
// (src/embeddings.py:81). Index, coefficient and summation arithmetic follow the x86 build of ATen operation by
// operation (which products are fused is probed bit-exact in oracle/restate.py:_cubic_coefficients / _cubic_sum).
__device__ __forceinline__ float cubic_cc1(float x) {  // ((A+2)x - (A+3)) x x + 1
  const float t1 = __fmaf_rn(1.25f, x, -2.25f);
  return __fadd_rn(__fmul_rn(__fmul_rn(t1, x), x), 1.f);
}
__device__ __forceinline__ float cubic_cc2(float x) {  // ((A x - 5A) x + 8A) x - 4A
  const float t2 = __fmaf_rn(__fmaf_rn(-0.75f, x, 3.75f), x, -6.f);
  return __fadd_rn(__fmul_rn(t2, x), 3.f);
}
__device__ __forceinline__ void cubic_index(float scale, int dst, int size, int& i0, float w[4]) {
  const float s = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);  // cubic: not clamped at zero
  i0 = (int)floorf(s);
  if (i0 > size - 1) i0 = size - 1;
  float t = __fsub_rn(s, (float)i0);
  t = fminf(fmaxf(t, 0.f), 1.f);
  const float x2 = __fsub_rn(1.f, t);
  w[0] = cubic_cc2(__fadd_rn(t, 1.f));
  w[1] = cubic_cc1(t);
  w[2] = cubic_cc1(x2);
  w[3] = cubic_cc2(__fadd_rn(x2, 1.f));
}
__device__ __forceinline__ float cubic_sum(float t0, float t1, float t2, float t3, const float w[4]) {
  return __fmaf_rn(t3, w[3], __fmaf_rn(t2, w[2], __fmaf_rn(t0, w[0], __fmul_rn(t1, w[1]))));
}
// one output pixel, 3 channels starting at byte `ch` of the staged HWC rows (row r of the image = s + (r - r_lo) * row_bytes)
__device__ __forceinline__ void cubic_pixel(const PreParams& p, const uint8_t* s, int r_lo, long long row_bytes, int iy,
                                            const float wy[4], int ix, const float wx[4], int ch, const float* lut,
                                            float o[3]) {
  int col[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) col[k] = min(max(ix + k - 1, 0), p.W - 1) * p.CH + ch;
  float rows[3][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint8_t* r = s + (long long)(min(max(iy + j - 1, 0), p.H - 1) - r_lo) * row_bytes;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      rows[c][j] = cubic_sum((float)r[col[0] + c], (float)r[col[1] + c], (float)r[col[2] + c], (float)r[col[3] + c], wx);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v = cubic_sum(rows[c][0], rows[c][1], rows[c][2], rows[c][3], wy);
    v = fminf(fmaxf(v, 0.f), 255.f);  // torchvision clamps the overshoot before the rounding cast
    o[c] = lut[c * 256 + (int)rintf(v)];
  }
}

template <bool CUBIC>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* lut = reinterpret_cast<float*>(smem + 16);  // [3][256]
  uint8_t* stage = smem + 16 + 3 * 256 * 4;

  const int img = blockIdx.x / p.bands;
  const int band = blockIdx.x - img * p.bands;
  const int y_first = band * p.rows;
  const int y_count = min(p.rows, p.crop - y_first);

  // input row range of this band
  int r_lo, r_hi, tmp;
  float lf;
  if (CUBIC) {
    float wtmp[4];
    cubic_index(p.scale_y, y_first + p.top, p.H, r_lo, wtmp);
    cubic_index(p.scale_y, y_first + y_count - 1 + p.top, p.H, r_hi, wtmp);
    r_lo = max(r_lo - 1, 0);
    r_hi = min(r_hi + 2, p.H - 1);
  } else {
    src_index(p.scale_y, y_first + p.top, p.H, r_lo, tmp, lf);
    src_index(p.scale_y, y_first + y_count - 1 + p.top, p.H, tmp, r_hi, lf);
    if (p.swap02 && r_lo <= 2) {  // a band that touches rows 0..2 stages all three
      r_lo = 0;
      r_hi = max(r_hi, 2);
    }
  }
  const long long row_bytes = (long long)p.W * p.CH;
  const long long g0 = (long long)img * p.H * row_bytes + (long long)r_lo * row_bytes;
  const long long nbytes = (long long)(r_hi - r_lo + 1) * row_bytes;
  const long long a0 = g0 & ~15ll;
  const int head = (int)(g0 - a0);
  long long want = (head + nbytes + 15) & ~15ll;
  long long avail = (p.total_bytes - a0) & ~15ll;  // never read past the tensor with the bulk engine
  const uint32_t bulk = (uint32_t)(want < avail ? want : avail);

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, bulk);
    bulk_load_1d(stage, p.in + a0, bulk, bar);
  }
  // tail bytes not covered by the 16-byte granular bulk copy (only at the very end of the tensor)
  for (long long t = bulk + threadIdx.x; t < head + nbytes; t += blockDim.x) stage[t] = p.in[a0 + t];
  // normalisation table: ((u / 255) - mean) / std, each op rounded to fp32 like the reference
  for (int t = threadIdx.x; t < 768; t += blockDim.x) {
    const int c = t >> 8, u = t & 255;
    lut[t] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.0f), p.mean[c]), p.stdv[c]);
  }
  __syncthreads();
  mbar_wait(bar, 0);

  const uint8_t* s = stage + head;
  const int npix = y_count * p.crop;
  const long long plane = (long long)p.crop * p.crop;
  if (p.fmt == PVR_FMT_STEM_BF16 || p.fmt == PVR_FMT_STEM_PAD_BF16) {
    // W-expanded stem input: out[image][y][q][8 columns 2q-3..2q+4][4 ch] bf16 (64 B per output column of the 7x7/2
    // stem), so that the stem conv is a 7x1-tap implicit GEMM with 64-byte TMA rows. Pixels are first written to a
    // zero-margined bf16 row buffer in smem, then copied out 16 B per thread, fully coalesced.
    uint2* pix = reinterpret_cast<uint2*>(smem + p.pix_off);
    const int prow = p.crop + 8;  // entry x+4 holds column x; 4 zero entries on each side
    const int Q = p.crop >> 1;
    for (int t = threadIdx.x; t < y_count * 8; t += blockDim.x) {
      const int yy = t >> 3, e = t & 7;
      pix[yy * prow + (e < 4 ? e : p.crop + e)] = make_uint2(0u, 0u);
    }
    for (int f = 0; f < p.nf; ++f) {
      for (int idx = threadIdx.x; idx < npix; idx += blockDim.x) {
        const int yy = idx / p.crop;
        const int x = idx - yy * p.crop;
        float o[3];
        if (CUBIC) {
          int iy, ix;
          float wy[4], wx[4];
          cubic_index(p.scale_y, y_first + yy + p.top, p.H, iy, wy);
          cubic_index(p.scale_x, x + p.left, p.W, ix, wx);
          cubic_pixel(p, s, r_lo, row_bytes, iy, wy, ix, wx, 3 * f, lut, o);
        } else {
          int r0, r1, c0, c1;
          float ly, lx;
          src_index(p.scale_y, y_first + yy + p.top, p.H, r0, r1, ly);
          src_index(p.scale_x, x + p.left, p.W, c0, c1, lx);
          r0 = swap_row(p, r0);
          r1 = swap_row(p, r1);
          const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
          const uint8_t* q00 = s + (long long)(r0 - r_lo) * row_bytes + c0 * p.CH + 3 * f;
          const uint8_t* q01 = s + (long long)(r0 - r_lo) * row_bytes + c1 * p.CH + 3 * f;
          const uint8_t* q10 = s + (long long)(r1 - r_lo) * row_bytes + c0 * p.CH + 3 * f;
          const uint8_t* q11 = s + (long long)(r1 - r_lo) * row_bytes + c1 * p.CH + 3 * f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float top = __fmaf_rn((float)q00[c], hx, __fmul_rn((float)q01[c], lx));
            const float bot = __fmaf_rn((float)q10[c], hx, __fmul_rn((float)q11[c], lx));
            const float v = __fmaf_rn(top, hy, __fmul_rn(bot, ly));
            o[c] = finish_bilinear(p, v, c, lut);
          }
        }
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], 0.f);
        pix[yy * prow + x + 4] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
      __syncthreads();
      const long long image = p.sample_major ? (long long)img * p.nf + f : (long long)f * p.N + img;
      if (p.fmt == PVR_FMT_STEM_PAD_BF16) {
        // compact variant: the padded NHWC4 row itself (buffer pixel j = column j - 3 = row-buffer entry j + 1),
        // 16 B (two pixels) per thread; the stem's tensor map does the W-expansion (include/pvr_b200.h)
        const int R = prow >> 1;  // uint4 per row
        uint4* dst = reinterpret_cast<uint4*>(p.out) + (image * p.crop + y_first) * (long long)R;
        for (int t = threadIdx.x; t < y_count * R; t += blockDim.x) {
          const int yy = t / R, k = t - yy * R;
          const uint2 e0 = pix[yy * prow + 2 * k + 1];
          const uint2 e1 = 2 * k + 2 < prow ? pix[yy * prow + 2 * k + 2] : make_uint2(0u, 0u);
          dst[t] = make_uint4(e0.x, e0.y, e1.x, e1.y);
        }
        __syncthreads();
        continue;
      }
      uint4* dst = reinterpret_cast<uint4*>(p.out) + (image * p.crop + y_first) * (long long)Q * 4;
      for (int t = threadIdx.x; t < y_count * Q * 4; t += blockDim.x) {
        const int yy = t / (Q * 4);
        const int rem = t - yy * Q * 4;
        const int q = rem >> 2, k = rem & 3;
        const uint2 e0 = pix[yy * prow + 2 * q + 1 + 2 * k];
        const uint2 e1 = pix[yy * prow + 2 * q + 2 + 2 * k];
        dst[t] = make_uint4(e0.x, e0.y, e1.x, e1.y);
      }
      __syncthreads();
    }
    return;
  }
  for (int idx = threadIdx.x; idx < npix; idx += blockDim.x) {
    const int yy = idx / p.crop;
    const int x = idx - yy * p.crop;
    const int y = y_first + yy;
    int r0 = 0, r1 = 0, c0 = 0, c1 = 0;
    float ly = 0.f, lx = 0.f, wy[4], wx[4];
    if (CUBIC) {
      cubic_index(p.scale_y, y + p.top, p.H, r0, wy);
      cubic_index(p.scale_x, x + p.left, p.W, c0, wx);
    } else {
      src_index(p.scale_y, y + p.top, p.H, r0, r1, ly);
      src_index(p.scale_x, x + p.left, p.W, c0, c1, lx);
      r0 = swap_row(p, r0);
      r1 = swap_row(p, r1);
    }
    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    const uint8_t* q00 = s + (long long)(r0 - r_lo) * row_bytes + c0 * p.CH;
    const uint8_t* q01 = s + (long long)(r0 - r_lo) * row_bytes + c1 * p.CH;
    const uint8_t* q10 = s + (long long)(r1 - r_lo) * row_bytes + c0 * p.CH;
    const uint8_t* q11 = s + (long long)(r1 - r_lo) * row_bytes + c1 * p.CH;
    for (int f = 0; f < p.nf; ++f) {
      float o[3];
      if (CUBIC) cubic_pixel(p, s, r_lo, row_bytes, r0, wy, c0, wx, 3 * f, lut, o);
#pragma unroll
      for (int c = 0; c < 3 && !CUBIC; ++c) {
        const int ch = 3 * f + c;
        // ATen Interpolate<>::eval as compiled for x86: fma(t0, w0, round(t1 * w1)) per dimension
        const float top = __fmaf_rn((float)q00[ch], hx, __fmul_rn((float)q01[ch], lx));
        const float bot = __fmaf_rn((float)q10[ch], hx, __fmul_rn((float)q11[ch], lx));
        const float v = __fmaf_rn(top, hy, __fmul_rn(bot, ly));
        o[c] = finish_bilinear(p, v, c, lut);
      }
      // frame-major is the reference's np.concatenate(np.split(o, n, 3), 0) order
      const long long image = p.sample_major ? (long long)img * p.nf + f : (long long)f * p.N + img;
      if (p.fmt == PVR_FMT_NCHW_F32) {
        float* dst = reinterpret_cast<float*>(p.out) + image * 3 * plane + (long long)y * p.crop + x;
        dst[0] = o[0];
        dst[plane] = o[1];
        dst[2 * plane] = o[2];
      } else if (p.fmt == PVR_FMT_NHWC4_F32) {
        float4* dst = reinterpret_cast<float4*>(p.out) + image * plane + (long long)y * p.crop + x;
        *dst = make_float4(o[0], o[1], o[2], 0.f);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], 0.f);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t*>(&a);
        v.y = *reinterpret_cast<uint32_t*>(&b);
        uint2* dst = reinterpret_cast<uint2*>(p.out) + image * plane + (long long)y * p.crop + x;
        *dst = v;
      }
    }
  }
}

}  // namespace pvr

extern void pvr_set_error(const char* fmt, ...);

extern "C" int pvr_preprocess_u8(const uint8_t* in, int N, int H, int W, int n_frames, int rh, int rw, int top,
                                 int left, int crop, const float* mean, const float* stdv, void* out, int out_fmt,
                                 int sample_major, void* stream) {
  using namespace pvr;
  const bool cubic = (out_fmt & PVR_RESIZE_BICUBIC) != 0;
  const bool flt = (out_fmt & PVR_RESIZE_FLOAT) != 0, swap02 = (out_fmt & PVR_SWAP_ROWS_0_2) != 0;
  out_fmt &= ~(PVR_RESIZE_BICUBIC | PVR_RESIZE_FLOAT | PVR_SWAP_ROWS_0_2);
  if ((flt || swap02) && (cubic || H < 3)) {
    pvr_set_error("pvr_preprocess_u8: PVR_RESIZE_FLOAT / PVR_SWAP_ROWS_0_2 are bilinear-only and need >= 3 rows");
    return PVR_ERR_ARG;
  }
  if (!in || !out || N <= 0 || H <= 0 || W <= 0 || n_frames <= 0 || rh <= 0 || rw <= 0 || crop <= 0 || top < 0 ||
      left < 0 || top + crop > rh || left + crop > rw || !mean || !stdv ||
      (out_fmt != PVR_FMT_NCHW_F32 && out_fmt != PVR_FMT_NHWC4_BF16 && out_fmt != PVR_FMT_STEM_BF16 &&
       out_fmt != PVR_FMT_NHWC4_F32 && out_fmt != PVR_FMT_STEM_PAD_BF16) ||
      ((out_fmt == PVR_FMT_STEM_BF16 || out_fmt == PVR_FMT_STEM_PAD_BF16) && (crop & 1))) {
    pvr_set_error("pvr_preprocess_u8: invalid argument");
    return PVR_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    pvr_set_error("pvr_preprocess_u8: in/out must be 16-byte aligned");
    return PVR_ERR_ARG;
  }
  PreParams p;
  p.in = in;
  p.out = out;
  p.N = N; p.H = H; p.W = W; p.nf = n_frames; p.CH = 3 * n_frames;
  p.total_bytes = (long long)N * H * W * p.CH;
  p.top = top; p.left = left; p.crop = crop;
  p.scale_y = (float)H / (float)rh;  // ATen area_pixel_compute_scale with an explicit output size
  p.scale_x = (float)W / (float)rw;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  p.fmt = out_fmt;
  p.sample_major = sample_major ? 1 : 0;
  p.flt = flt ? 1 : 0;
  p.swap02 = swap02 ? 1 : 0;
  // rows per band: keep the staged input under ~48 KiB so several CTAs share an SM
  const long long row_bytes = (long long)W * p.CH;
  int rows = 16;
  auto stage_bytes = [&](int r) { return ((long long)(p.scale_y * r) + (cubic ? 5 : (swap02 ? 6 : 3))) * row_bytes + 48; };
  while (rows > 1 && stage_bytes(rows) > 48 * 1024) rows >>= 1;
  long long smem = 16 + 3072 + stage_bytes(rows);
  p.pix_off = 0;
  if (out_fmt == PVR_FMT_STEM_BF16 || out_fmt == PVR_FMT_STEM_PAD_BF16) {
    p.pix_off = (int)((smem + 15) & ~15ll);
    smem = p.pix_off + (long long)rows * (crop + 8) * 8;
  }
  if (smem > 200 * 1024) {
    pvr_set_error("pvr_preprocess_u8: input rows too wide for shared-memory staging (%lld bytes)", smem);
    return PVR_ERR_ARG;
  }
  p.rows = rows;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(preprocess_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    attr = true;
  }
  p.bands = (crop + rows - 1) / rows;
  if ((long long)p.bands * N > 0x7fffffffll) {
    pvr_set_error("pvr_preprocess_u8: batch too large for one launch");
    return PVR_ERR_ARG;
  }
  dim3 grid((unsigned)(p.bands * N));
  if (cubic) preprocess_kernel<true><<<grid, 256, (size_t)smem, static_cast<cudaStream_t>(stream)>>>(p);
  else preprocess_kernel<false><<<grid, 256, (size_t)smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}
