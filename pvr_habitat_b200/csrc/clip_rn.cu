// The two pieces of CLIP's ModifiedResNet (`clip.load("RN50")`, src/embeddings.py:305-306; openai/CLIP clip/model.py
// ModifiedResNet / Bottleneck / AttentionPool2d) that the ResNet program did not have:
//   * 2 x 2 / stride 2 average pooling of an NHWC activation (the anti-aliased stride of CLIP's Bottleneck: avgpool
//     after conv2 and in front of the shortcut's 1x1 conv, and the stem's avgpool) — encoder op PVR_OP_AVGPOOL2;
//   * token assembly of the attention pool: [mean over the 7 x 7 positions | positions] + positional_embedding.
// Both are HBM bound (one read, a quarter / one write) and vectorised 16 bytes per thread. The attention pool's
// projections are plain GEMMs and its 32-head attention runs in attention_mma.cu (50 tokens, head_dim 64).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    f[2 * t] = __uint_as_float(w[t] << 16);
    f[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]),
                 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c);
  o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

// out[img][p][q][c] = (in[2p][2q] + in[2p][2q+1] + in[2p+1][2q] + in[2p+1][2q+1]) / 4 (nn.AvgPool2d(2): floor mode).
// bf16: 8 channels per thread, fp32 accumulation; C % 8 == 0.
__global__ void __launch_bounds__(256) avgpool2_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                            __nv_bfloat16* __restrict__ out, long long total, int H,
                                                            int W, int C, int P, int Q) {
  const int cg = C >> 3;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const long long img = t / P;
    const __nv_bfloat16* base = in + ((img * H + 2 * p) * W + 2 * q) * C + 8 * g;
    float acc[8], f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base)), acc);
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)W * C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
    unpack8(__ldg(reinterpret_cast<const uint4*>(base + (long long)W * C + C)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (acc[j] + f[j]) * 0.25f;
    *reinterpret_cast<uint4*>(out + ((img * P + p) * Q + q) * C + 8 * g) = pack8(acc);
  }
}

// fp32 parity mode: 4 channels per thread; C % 4 == 0
__global__ void __launch_bounds__(256) avgpool2_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           long long total, int H, int W, int C, int P, int Q) {
  const int cg = C >> 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int q = (int)(t % Q);
    t /= Q;
    const int p = (int)(t % P);
    const long long img = t / P;
    const float* base = in + ((img * H + 2 * p) * W + 2 * q) * C + 4 * g;
    const float4 a = __ldg(reinterpret_cast<const float4*>(base));
    const float4 b = __ldg(reinterpret_cast<const float4*>(base + C));
    const float4 c = __ldg(reinterpret_cast<const float4*>(base + (long long)W * C));
    const float4 d = __ldg(reinterpret_cast<const float4*>(base + (long long)W * C + C));
    float4 o;
    o.x = (((a.x + b.x) + c.x) + d.x) * 0.25f;
    o.y = (((a.y + b.y) + c.y) + d.y) * 0.25f;
    o.z = (((a.z + b.z) + c.z) + d.z) * 0.25f;
    o.w = (((a.w + b.w) + c.w) + d.w) * 0.25f;
    *reinterpret_cast<float4*>(out + ((img * P + p) * Q + q) * C + 4 * g) = o;
  }
}

// tokens[img][0] = mean_i x[img][i] + pos[0]; tokens[img][i + 1] = x[img][i] + pos[i + 1]. One block per
// (image, 256-channel slab of 4-channel groups): thread = 4 channels, loops over the hw positions once.
template <typename T>
__global__ void __launch_bounds__(256) attnpool_tokens_kernel(const T* __restrict__ x, const float* __restrict__ pos,
                                                              int hw, int C, T* __restrict__ tok) {
  const long long img = blockIdx.y;
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= C) return;
  const T* xi = x + img * hw * C + c;
  T* ti = tok + img * (hw + 1) * C + c;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < hw; ++i) {
    float v[4];
    if constexpr (sizeof(T) == 2) {
      const uint2 u = *reinterpret_cast<const uint2*>(xi + (long long)i * C);
      v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
      v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
    } else {
      const float4 u = *reinterpret_cast<const float4*>(xi + (long long)i * C);
      v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    }
    const float4 pe = *reinterpret_cast<const float4*>(pos + (long long)(i + 1) * C + c);
#pragma unroll
    for (int k = 0; k < 4; ++k) s[k] += v[k];
    const float o[4] = {v[0] + pe.x, v[1] + pe.y, v[2] + pe.z, v[3] + pe.w};
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
      *reinterpret_cast<uint2*>(ti + (long long)(i + 1) * C) =
          make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    } else {
      *reinterpret_cast<float4*>(ti + (long long)(i + 1) * C) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  const float4 p0 = *reinterpret_cast<const float4*>(pos + c);
  const float inv = 1.f / (float)hw;
  const float o[4] = {s[0] * inv + p0.x, s[1] * inv + p0.y, s[2] * inv + p0.z, s[3] * inv + p0.w};
  if constexpr (sizeof(T) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(o[0], o[1]), b = __floats2bfloat162_rn(o[2], o[3]);
    *reinterpret_cast<uint2*>(ti) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  } else {
    *reinterpret_cast<float4*>(ti) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace

cudaError_t launch_avgpool2(const void* in, void* out, int n_img, int H, int W, int C, int f32, cudaStream_t stream) {
  const int P = H / 2, Q = W / 2;
  const int per = f32 ? 4 : 8;
  if (C % per || P <= 0 || Q <= 0) return cudaErrorInvalidValue;
  const long long total = (long long)n_img * P * Q * (C / per);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (f32)
    avgpool2_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out),
                                                              total, H, W, C, P, Q);
  else
    avgpool2_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(in),
                                                               static_cast<__nv_bfloat16*>(out), total, H, W, C, P, Q);
  return cudaGetLastError();
}

}  // namespace pvr

extern "C" int pvr_attnpool_tokens(const void* x, int n_img, int hw, int width, const float* pos, int f32, void* tokens,
                                   void* stream) {
  if (!x || !pos || !tokens || n_img <= 0 || hw <= 0 || width <= 0 || width % 4 || n_img > 65535) {
    pvr_set_error("pvr_attnpool_tokens: invalid argument");
    return PVR_ERR_ARG;
  }
  const dim3 grid((unsigned)((width / 4 + 255) / 256), (unsigned)n_img);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32)
    pvr::attnpool_tokens_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), pos, hw, width,
                                                             static_cast<float*>(tokens));
  else
    pvr::attnpool_tokens_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), pos, hw, width,
                                                                     static_cast<__nv_bfloat16*>(tokens));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pvr_set_error("pvr_attnpool_tokens: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}
