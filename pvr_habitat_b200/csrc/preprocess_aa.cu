// K1, antialiased bicubic variant — CLIP transforms (src/embeddings.py:309-314): Resize(res, BICUBIC, antialias=True)
// -> CenterCrop(res) -> ConvertImageDtype(float) -> Normalize(CLIP mean / std) + frame split, one pass over HBM.
//
// torchvision casts the uint8 image to float32 and calls ATen's separable antialiased bicubic resize (a = -0.5,
// normalised weights, support 2 * max(scale, 1)): a horizontal pass over every input row into a float32 intermediate,
// then a vertical pass; the result is clamped to [0, 255], rounded half to even to uint8, then x/255 and
// (x - mean)/std (the 3 x 256 table of preprocess.cu). The bit-deciding arithmetic is preprocess_aa_core.cuh, shared
// with a host harness that is checked against the oracle on CPU (tests/test_preprocess_aa_core.py).
//
// One CTA produces `rows` output rows of all frames of one observation: the input rows they depend on (one contiguous
// byte range of the HWC image) are staged in shared memory by a 1-D bulk async copy, the horizontal pass writes the
// float32 intermediate rows (crop columns only) to shared memory, the vertical pass reads them back.
// Tap ranges and weights per output row / column come from two small tables in global memory, computed once per
// (input size, output size) by aa_weights_kernel and cached.
//
// Verified on B200 (round 2): bit-exact against the oracle and the reference's own transforms on 14 geometries
// (tests/test_gpu_preprocess_aa.py), compute-sanitizer clean (tools/sanitize_aa.py).
#include <mutex>
#include <vector>

#include "preprocess_aa_core.cuh"
#include "pvr_b200.h"
#include "ptx.cuh"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

struct AAParams {
  AAGeom g;
  const uint8_t* in;
  void* out;
  long long total_bytes;
  int max_in_rows;
  float mean[3], stdv[3];
  int fmt;
  int tmp_off;  // byte offset of the float32 intermediate inside dynamic smem
};

__global__ void aa_weights_kernel(int in_size, int out_size, int* xmin, int* size, float* w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < out_size) aa_index_weights(i, in_size, out_size, &xmin[i], &size[i], w + (long long)i * AA_MAX_TAPS);
}

__global__ void __launch_bounds__(256) preprocess_aa_kernel(const AAParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  float* lut = reinterpret_cast<float*>(smem + 16);  // [3][256]
  uint8_t* stage = smem + 16 + 3 * 256 * 4;
  float* tmp = reinterpret_cast<float*>(smem + p.tmp_off);
  const AAGeom& g = p.g;

  const AABand b = aa_band(g, blockIdx.x);
  if (b.rows_in > p.max_in_rows) __trap();  // the host sized the shared memory from an upper bound

  const long long row_bytes = (long long)g.W * g.CH;
  const long long g0 = ((long long)b.img * g.H + b.r_lo) * row_bytes;
  const long long nbytes = (long long)b.rows_in * row_bytes;
  const long long a0 = g0 & ~15ll;
  const int head = (int)(g0 - a0);
  const long long want = (head + nbytes + 15) & ~15ll;
  const long long avail = (p.total_bytes - a0) & ~15ll;  // never read past the tensor with the bulk engine
  const uint32_t bulk = (uint32_t)(want < avail ? want : avail);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, bulk);
    bulk_load_1d(stage, p.in + a0, bulk, bar);
  }
  for (long long t = bulk + threadIdx.x; t < head + nbytes; t += blockDim.x) stage[t] = p.in[a0 + t];
  aa_build_lut(lut, p.mean, p.stdv, threadIdx.x, blockDim.x);
  __syncthreads();
  mbar_wait(bar, 0);

  const uint8_t* s = stage + head;
  const long long plane = (long long)g.crop * g.crop;
  const int crop = g.crop, fmt = p.fmt;
  void* out = p.out;
  for (int f = 0; f < g.nf; ++f) {
    aa_horizontal(g, b, s, tmp, f, threadIdx.x, blockDim.x);
    __syncthreads();
    aa_vertical(g, b, tmp, lut, f, threadIdx.x, blockDim.x, [&](long long image, int y, int x, const float* o) {
      if (fmt == PVR_FMT_NCHW_F32) {
        float* dst = reinterpret_cast<float*>(out) + image * 3 * plane + (long long)y * crop + x;
        dst[0] = o[0];
        dst[plane] = o[1];
        dst[2 * plane] = o[2];
      } else if (fmt == PVR_FMT_NHWC4_F32) {
        float4* dst = reinterpret_cast<float4*>(out) + image * plane + (long long)y * crop + x;
        *dst = make_float4(o[0], o[1], o[2], 0.f);
      } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]);
        __nv_bfloat162 bb = __floats2bfloat162_rn(o[2], 0.f);
        uint2 v;
        v.x = *reinterpret_cast<uint32_t*>(&a);
        v.y = *reinterpret_cast<uint32_t*>(&bb);
        uint2* dst = reinterpret_cast<uint2*>(out) + image * plane + (long long)y * crop + x;
        *dst = v;
      }
    });
    __syncthreads();
  }
}

// ---- weight tables, one per (device, input size, output size), built on first use
struct AATable {
  int dev, in_size, out_size;
  int *xmin, *size;
  float* w;
};
std::mutex g_tables_mutex;
std::vector<AATable> g_tables;

bool aa_table(int in_size, int out_size, cudaStream_t stream, AATable* out, cudaError_t* err) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  for (const AATable& t : g_tables)
    if (t.dev == dev && t.in_size == in_size && t.out_size == out_size) {
      *out = t;
      return true;
    }
  AATable t{dev, in_size, out_size, nullptr, nullptr, nullptr};
  *err = cudaMalloc(&t.xmin, sizeof(int) * out_size);
  if (*err == cudaSuccess) *err = cudaMalloc(&t.size, sizeof(int) * out_size);
  if (*err == cudaSuccess) *err = cudaMalloc(&t.w, sizeof(float) * (size_t)out_size * AA_MAX_TAPS);
  if (*err != cudaSuccess) return false;
  // first use of a geometry (not capturable: cudaMalloc + synchronisation); the tables are complete before any
  // consumer on any stream is launched and are never written again
  aa_weights_kernel<<<(out_size + 127) / 128, 128, 0, stream>>>(in_size, out_size, t.xmin, t.size, t.w);
  *err = cudaGetLastError();
  if (*err == cudaSuccess) *err = cudaStreamSynchronize(stream);
  if (*err != cudaSuccess) return false;
  g_tables.push_back(t);
  *out = t;
  return true;
}

// upper bound of the taps of one output index, and of the input rows a band of `rows` output rows depends on
int max_taps(int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.f ? 2.f * scale : 2.f;
  return (int)ceilf(support) * 2 + 1;
}

}  // namespace
}  // namespace pvr

extern "C" int pvr_preprocess_u8_aa(const uint8_t* in, int N, int H, int W, int n_frames, int rh, int rw, int top,
                                    int left, int crop, const float* mean, const float* stdv, void* out, int out_fmt,
                                    int sample_major, void* stream_) {
  using namespace pvr;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!in || !out || N <= 0 || H <= 0 || W <= 0 || n_frames <= 0 || rh <= 0 || rw <= 0 || crop <= 0 || top < 0 ||
      left < 0 || top + crop > rh || left + crop > rw || !mean || !stdv ||
      (out_fmt != PVR_FMT_NCHW_F32 && out_fmt != PVR_FMT_NHWC4_BF16 && out_fmt != PVR_FMT_NHWC4_F32)) {
    pvr_set_error("pvr_preprocess_u8_aa: invalid argument");
    return PVR_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) {
    pvr_set_error("pvr_preprocess_u8_aa: in/out must be 16-byte aligned");
    return PVR_ERR_ARG;
  }
  const int taps_y = max_taps(H, rh), taps_x = max_taps(W, rw);
  if (taps_y > AA_MAX_TAPS || taps_x > AA_MAX_TAPS) {
    pvr_set_error("pvr_preprocess_u8_aa: down-scaling factor too large (%d x %d -> %d x %d)", H, W, rh, rw);
    return PVR_ERR_ARG;
  }
  cudaError_t err = cudaSuccess;
  AATable ty, tx;
  if (!aa_table(H, rh, stream, &ty, &err) || !aa_table(W, rw, stream, &tx, &err)) {
    pvr_set_error("pvr_preprocess_u8_aa: weight tables: %s", cudaGetErrorString(err));
    return PVR_ERR_CUDA;
  }
  AAParams p;
  AAGeom& g = p.g;
  p.in = in;
  p.out = out;
  g.N = N; g.H = H; g.W = W; g.nf = n_frames; g.CH = 3 * n_frames;
  p.total_bytes = (long long)N * H * W * g.CH;
  g.top = top; g.left = left; g.crop = crop;
  g.ymin = ty.xmin; g.ysize = ty.size; g.wy = ty.w;
  g.xmin = tx.xmin; g.xsize = tx.size; g.wx = tx.w;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  p.fmt = out_fmt;
  g.sample_major = sample_major ? 1 : 0;
  // rows per band: staged input rows + float32 intermediate rows must fit in shared memory
  const long long row_bytes = (long long)W * g.CH;
  const float scale_y = (float)H / (float)rh;
  auto rows_in = [&](int r) { return (int)ceilf(scale_y * (float)r) + taps_y + 1; };
  auto smem_bytes = [&](int r, int* tmp_off) {
    long long off = 16 + 3072 + (long long)rows_in(r) * row_bytes + 48;
    off = (off + 15) & ~15ll;
    *tmp_off = (int)off;
    return off + (long long)rows_in(r) * crop * 3 * 4;
  };
  int rows = 16, tmp_off = 0;
  while (rows > 1 && smem_bytes(rows, &tmp_off) > 160 * 1024) rows >>= 1;
  const long long smem = smem_bytes(rows, &tmp_off);
  if (smem > 200 * 1024) {
    pvr_set_error("pvr_preprocess_u8_aa: input rows too wide for shared-memory staging (%lld bytes)", smem);
    return PVR_ERR_ARG;
  }
  g.rows = rows;
  p.max_in_rows = rows_in(rows);
  p.tmp_off = tmp_off;
  g.bands = (crop + rows - 1) / rows;
  if ((long long)g.bands * N > 0x7fffffffll) {
    pvr_set_error("pvr_preprocess_u8_aa: batch too large for one launch");
    return PVR_ERR_ARG;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(preprocess_aa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8_aa: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    attr = true;
  }
  preprocess_aa_kernel<<<(unsigned)(g.bands * N), 256, (size_t)smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_preprocess_u8_aa: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}
