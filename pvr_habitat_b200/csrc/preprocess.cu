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
};

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
  src_index(p.scale_y, y_first + p.top, p.H, r_lo, tmp, lf);
  src_index(p.scale_y, y_first + y_count - 1 + p.top, p.H, tmp, r_hi, lf);
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
  if (p.fmt == PVR_FMT_STEM_BF16) {
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
        int r0, r1, c0, c1;
        float ly, lx;
        src_index(p.scale_y, y_first + yy + p.top, p.H, r0, r1, ly);
        src_index(p.scale_x, x + p.left, p.W, c0, c1, lx);
        const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
        const uint8_t* q00 = s + (long long)(r0 - r_lo) * row_bytes + c0 * p.CH + 3 * f;
        const uint8_t* q01 = s + (long long)(r0 - r_lo) * row_bytes + c1 * p.CH + 3 * f;
        const uint8_t* q10 = s + (long long)(r1 - r_lo) * row_bytes + c0 * p.CH + 3 * f;
        const uint8_t* q11 = s + (long long)(r1 - r_lo) * row_bytes + c1 * p.CH + 3 * f;
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float top = __fmaf_rn((float)q00[c], hx, __fmul_rn((float)q01[c], lx));
          const float bot = __fmaf_rn((float)q10[c], hx, __fmul_rn((float)q11[c], lx));
          const float v = __fmaf_rn(top, hy, __fmul_rn(bot, ly));
          int u = (int)rintf(v);
          u = min(max(u, 0), 255);
          o[c] = lut[c * 256 + u];
        }
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(o[2], 0.f);
        pix[yy * prow + x + 4] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
      }
      __syncthreads();
      const long long image = p.sample_major ? (long long)img * p.nf + f : (long long)f * p.N + img;
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
    int r0, r1, c0, c1;
    float ly, lx;
    src_index(p.scale_y, y + p.top, p.H, r0, r1, ly);
    src_index(p.scale_x, x + p.left, p.W, c0, c1, lx);
    const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
    const uint8_t* q00 = s + (long long)(r0 - r_lo) * row_bytes + c0 * p.CH;
    const uint8_t* q01 = s + (long long)(r0 - r_lo) * row_bytes + c1 * p.CH;
    const uint8_t* q10 = s + (long long)(r1 - r_lo) * row_bytes + c0 * p.CH;
    const uint8_t* q11 = s + (long long)(r1 - r_lo) * row_bytes + c1 * p.CH;
    for (int f = 0; f < p.nf; ++f) {
      float o[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int ch = 3 * f + c;
        // ATen Interpolate<>::eval as compiled for x86: fma(t0, w0, round(t1 * w1)) per dimension
        const float top = __fmaf_rn((float)q00[ch], hx, __fmul_rn((float)q01[ch], lx));
        const float bot = __fmaf_rn((float)q10[ch], hx, __fmul_rn((float)q11[ch], lx));
        const float v = __fmaf_rn(top, hy, __fmul_rn(bot, ly));
        int u = (int)rintf(v);  // half to even, as torch.round
        u = min(max(u, 0), 255);
        o[c] = lut[c * 256 + u];
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
  if (!in || !out || N <= 0 || H <= 0 || W <= 0 || n_frames <= 0 || rh <= 0 || rw <= 0 || crop <= 0 || top < 0 ||
      left < 0 || top + crop > rh || left + crop > rw || !mean || !stdv ||
      (out_fmt != PVR_FMT_NCHW_F32 && out_fmt != PVR_FMT_NHWC4_BF16 && out_fmt != PVR_FMT_STEM_BF16 &&
       out_fmt != PVR_FMT_NHWC4_F32) ||
      (out_fmt == PVR_FMT_STEM_BF16 && (crop & 1))) {
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
  // rows per band: keep the staged input under ~48 KiB so several CTAs share an SM
  const long long row_bytes = (long long)W * p.CH;
  int rows = 16;
  auto stage_bytes = [&](int r) { return ((long long)(p.scale_y * r) + 3) * row_bytes + 48; };
  while (rows > 1 && stage_bytes(rows) > 48 * 1024) rows >>= 1;
  long long smem = 16 + 3072 + stage_bytes(rows);
  p.pix_off = 0;
  if (out_fmt == PVR_FMT_STEM_BF16) {
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
    cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    attr = true;
  }
  p.bands = (crop + rows - 1) / rows;
  if ((long long)p.bands * N > 0x7fffffffll) {
    pvr_set_error("pvr_preprocess_u8: batch too large for one launch");
    return PVR_ERR_ARG;
  }
  dim3 grid((unsigned)(p.bands * N));
  preprocess_kernel<<<grid, 256, (size_t)smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}
