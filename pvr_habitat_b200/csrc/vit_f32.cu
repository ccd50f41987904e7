// fp32 parity mode of the ViT encoders (CLIP ViT-B, MAE ViT-B / ViT-L): the north star's second embedding tolerance
// (relative L2 <= 1e-5 "in the fp32 mode") for the transformer encoders of src/embeddings.py:137-148, 298-314.
// Everything stays float32 on the CUDA cores: GEMMs through the blocked-summation kernel of conv_f32.cu (a 1x1
// "convolution" over M rows), attention with one warp per query row, LayerNorm through pvr_layernorm_f32 (vit.cu).
// A checking mode, not a performance path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"
#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// x[row] = [cls | patch] + pos, optionally LayerNorm'ed (CLIP's ln_pre); one warp per token row, any width % 4 == 0.
__global__ void __launch_bounds__(256) vit_embed_f32_kernel(const float* __restrict__ patches,
                                                             const float* __restrict__ cls,
                                                             const float* __restrict__ pos, int tokens, int W,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, long long rows,
                                                             float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const long long img = row / tokens;
  const int tok = (int)(row - img * tokens);
  const float* src = tok == 0 ? cls : patches + (img * (tokens - 1) + tok - 1) * W;
  const float* pe = pos + (long long)tok * W;
  float* o = out + row * W;
  float s = 0.f;
  for (int c = lane; c < W; c += 32) {
    const float v = src[c] + pe[c];
    o[c] = v;
    s += v;
  }
  if (!gamma) return;
  const float mean = warp_sum(s) / W;
  float ss = 0.f;
  for (int c = lane; c < W; c += 32) {
    const float d = o[c] - mean;
    ss += d * d;
  }
  const float rstd = rsqrtf(warp_sum(ss) / W + eps);
  for (int c = lane; c < W; c += 32) o[c] = (o[c] - mean) * rstd * gamma[c] + beta[c];
}

// One block per (image, head): K and V of the head staged in shared memory, one warp per query row at a time.
// head_dim D <= 128: lane l owns dimensions l, l + 32, ... (2 values at D = 64; 3 with the last one guarded at the
// D = 80 of mae_huge).
template <int D>
__global__ void __launch_bounds__(256) vit_attention_f32_kernel(const float* __restrict__ qkv, int S, int W, int heads,
                                                                 float scale, float* __restrict__ out) {
  constexpr int PER = (D + 31) / 32;
  extern __shared__ float sm[];
  float* Ks = sm;                 // (S, D)
  float* Vs = sm + (size_t)S * D;
  float* Ps = Vs + (size_t)S * D;  // (8 warps, S) probabilities of the row a warp is working on
  const int img = blockIdx.x / heads, head = blockIdx.x % heads;
  const float* base = qkv + (long long)img * S * 3 * W + head * D;
  for (int i = threadIdx.x; i < S * D; i += blockDim.x) {
    const int t = i / D, d = i - t * D;
    Ks[i] = base[(long long)t * 3 * W + W + d];
    Vs[i] = base[(long long)t * 3 * W + 2 * W + d];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* P = Ps + (size_t)warp * S;
  for (int r = warp; r < S; r += 8) {
    float q[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) q[k] = lane + 32 * k < D ? base[(long long)r * 3 * W + lane + 32 * k] * scale : 0.f;
    float mx = -INFINITY;
    for (int j = 0; j < S; ++j) {
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (lane + 32 * k < D) part = fmaf(q[k], Ks[j * D + lane + 32 * k], part);
      const float s = warp_sum(part);
      if (lane == 0) P[j] = s;
      mx = fmaxf(mx, s);
    }
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const float e = expf(P[j] - mx);
      P[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) o[k] = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = P[j];
#pragma unroll
      for (int k = 0; k < PER; ++k)
        if (lane + 32 * k < D) o[k] = fmaf(pj, Vs[j * D + lane + 32 * k], o[k]);
    }
    float* dst = out + ((long long)img * S + r) * W + head * D;
#pragma unroll
    for (int k = 0; k < PER; ++k)
      if (lane + 32 * k < D) dst[lane + 32 * k] = o[k] / sum;
    __syncwarp();
  }
}

template <int D>
int launch_attention_f32(const float* qkv, int n_img, int tokens, int width, int heads, float* out, cudaStream_t stream) {
  const size_t smem = ((size_t)tokens * 2 * D + 8 * (size_t)tokens) * sizeof(float);
  if (smem > 200 * 1024) {
    pvr_set_error("pvr_attention_f32: sequence too long for shared-memory staging (%d tokens)", tokens);
    return PVR_ERR_ARG;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(vit_attention_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) { pvr_set_error("pvr_attention_f32: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    attr = true;
  }
  vit_attention_f32_kernel<D><<<n_img * heads, 256, smem, stream>>>(qkv, tokens, width, heads, 1.f / sqrtf((float)D), out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_attention_f32: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

}  // namespace
}  // namespace pvr

// out (M x N) = act(a (M x K) w^T (N x K) + bias (+ res)); everything float32, w dense. act: 0 none, 2 QuickGELU,
// 3 erf GELU. `res` may alias `out` (in-place residual stream).
extern "C" int pvr_gemm_f32(const float* a, int64_t lda, const float* w, const float* bias, const float* res,
                            int64_t ldr, float* out, int64_t ldo, int64_t M, int N, int K, int act, void* stream) {
  static float* ones = nullptr;
  static int ones_n = 0;
  if (!a || !w || !bias || !out || M <= 0 || N <= 0 || K <= 0 || K % 4 || lda % 4 || (act != 0 && act != 2 && act != 3)) {
    pvr_set_error("pvr_gemm_f32: invalid argument");
    return PVR_ERR_ARG;
  }
  if (N > ones_n) {  // unit scale vector (the conv kernel applies folded-BN scale / bias)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &cap);
    if (cap != cudaStreamCaptureStatusNone) {
      pvr_set_error("pvr_gemm_f32: first use inside a stream capture");
      return PVR_ERR_ARG;
    }
    const int n = N < 4096 ? 4096 : N;
    float* host = new float[n];
    for (int i = 0; i < n; ++i) host[i] = 1.f;
    if (ones) cudaFree(ones);
    if (cudaMalloc(&ones, n * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(ones, host, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      delete[] host;
      ones = nullptr;
      ones_n = 0;
      pvr_set_error("pvr_gemm_f32: cudaMalloc failed");
      return PVR_ERR_CUDA;
    }
    delete[] host;
    ones_n = n;
  }
  pvr::ConvF32Params p;
  p.in = a; p.w = w; p.scale = ones; p.bias = bias; p.res = res; p.out = out;
  p.M = M; p.N = N; p.C = K; p.H = 1; p.W = 1; p.P = 1; p.Q = 1; p.R = 1; p.S = 1;
  p.stride_h = 1; p.stride_w = 1; p.lower_h = 0; p.lower_w = 0;
  p.in_pitch = (int)lda; p.out_pitch = (int)ldo; p.out_coff = 0; p.res_pitch = (int)ldr; p.res_coff = 0;
  p.relu_n = 0; p.elu = act;
  cudaError_t e = pvr::launch_conv_f32(p, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { pvr_set_error("pvr_gemm_f32: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_vit_embed_f32(const float* patches, const float* cls, const float* pos, int n_img, int tokens,
                                 int width, const float* gamma, const float* beta, float eps, float* x_out,
                                 void* stream) {
  if (!patches || !cls || !pos || (!gamma != !beta) || !x_out || n_img <= 0 || tokens <= 1 || width <= 0) {
    pvr_set_error("pvr_vit_embed_f32: invalid argument");
    return PVR_ERR_ARG;
  }
  const long long rows = (long long)n_img * tokens;
  pvr::vit_embed_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      patches, cls, pos, tokens, width, gamma, beta, eps, rows, x_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_vit_embed_f32: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_attention_f32(const float* qkv, int n_img, int tokens, int width, int heads, float* out,
                                 void* stream) {
  if (!qkv || !out || n_img <= 0 || tokens <= 0 || heads <= 0 || width <= 0 || width % heads) {
    pvr_set_error("pvr_attention_f32: invalid argument");
    return PVR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (width / heads) {
    case 64: return pvr::launch_attention_f32<64>(qkv, n_img, tokens, width, heads, out, st);
    case 80: return pvr::launch_attention_f32<80>(qkv, n_img, tokens, width, heads, out, st);
    default:
      pvr_set_error("pvr_attention_f32: head_dim %d is not built (64, 80)", width / heads);
      return PVR_ERR_ARG;
  }
}
