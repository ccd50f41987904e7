// CLIP-architecture ViT pieces around the tcgen05 GEMMs (patch embedding, QKV, projection and MLP run through
// pvr_gemm / the encoder's conv op): token assembly + ln_pre, LayerNorm, and multi-head attention on tcgen05.
//
// Reference call site: src/embeddings.py:303-304, 375-376 (`clip.load("ViT-B/32")`, `encode_image`); the arithmetic is
// openai/CLIP's VisionTransformer (not vendored in the reference): conv1 (stride = patch, no bias) -> [class token |
// patches] + positional embedding -> ln_pre -> 12 x [x += attn(ln_1(x)); x += mlp(ln_2(x))] with QuickGELU ->
// ln_post(x[:, 0]) @ proj.
#include "pvr_b200.h"
#include "conv_gemm.cuh"
#include "ptx.cuh"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per token row. W % 128 == 0 (each lane owns W/32 values as float4 groups).
// mode 0: x = LN(src fp32 row)                      -> bf16 (ln_1 / ln_2 / ln_post)
// mode 1: x = LN([cls | patch] + pos) (ln_pre)      -> fp32 residual stream
template <int W>
__global__ void __launch_bounds__(256) vit_layernorm_kernel(const float* __restrict__ src, long long src_row_step,
                                                             const __nv_bfloat16* __restrict__ patches,
                                                             const float* __restrict__ cls,
                                                             const float* __restrict__ pos, int tokens,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, long long rows,
                                                             float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16, int mode) {
  constexpr int PER = W / 32;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[PER];
  if (mode == 0) {
    const float* s = src + row * src_row_step * W;
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(s + (i * 32 + lane) * 4);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
    const long long img = row / tokens;
    const int tok = (int)(row - img * tokens);
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 pe = *reinterpret_cast<const float4*>(pos + (long long)tok * W + c);
      float4 t;
      if (tok == 0) {
        t = *reinterpret_cast<const float4*>(cls + c);
      } else {
        const uint2 pk = *reinterpret_cast<const uint2*>(patches + (img * (tokens - 1) + tok - 1) * W + c);
        t.x = __uint_as_float(pk.x << 16); t.y = __uint_as_float(pk.x & 0xFFFF0000u);
        t.z = __uint_as_float(pk.y << 16); t.w = __uint_as_float(pk.y & 0xFFFF0000u);
      }
      v[4 * i] = t.x + pe.x; v[4 * i + 1] = t.y + pe.y; v[4 * i + 2] = t.z + pe.z; v[4 * i + 3] = t.w + pe.w;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum_f(s) * (1.f / W);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; ss += d * d; }
  const float rstd = rsqrtf(warp_sum_f(ss) * (1.f / W) + eps);
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    const float o0 = (v[4 * i] - mean) * rstd * g.x + b.x, o1 = (v[4 * i + 1] - mean) * rstd * g.y + b.y;
    const float o2 = (v[4 * i + 2] - mean) * rstd * g.z + b.z, o3 = (v[4 * i + 3] - mean) * rstd * g.w + b.w;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * W + c) = make_float4(o0, o1, o2, o3);
    if (out_bf16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), bb = __floats2bfloat162_rn(o2, o3);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&a);
      o.y = *reinterpret_cast<uint32_t*>(&bb);
      *reinterpret_cast<uint2*>(out_bf16 + row * W + c) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// One work item = (image, head, 128-row query tile). head_dim = 64 (one 128-byte swizzle row per token and head).
//   S = Q K^T      tcgen05.mma, Q (128 x 64) and K (KP x 64) K-major from TMA, fp32 scores in TMEM (KP columns)
//   P = softmax    one thread per query row reads its TMEM lane: row max / sum are thread-local; P is written as
//                  bf16 into 128B-swizzled K-major shared memory (64-key chunks)
//   O = P V        tcgen05.mma, V (KP x 64, straight from TMA) is the MN-major B operand; 64 fp32 columns in TMEM
// Keys beyond the sequence are masked to p = 0; query rows beyond it are computed but never stored.
struct AttnParams {
  int n_img, S, W, heads, KP, mtiles;
  float scale_log2e;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(128, 1)
vit_attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t kv_bytes = (uint32_t)p.KP * 128;
  const uint32_t kv_alloc = (kv_bytes + 1023) & ~1023u;
  uint8_t* sQ = smem;                       // 128 x 128 B
  uint8_t* sK = sQ + 16384;                 // KP x 128 B
  uint8_t* sV = sK + kv_alloc;              // KP x 128 B
  uint8_t* sP = sV + kv_alloc;              // ceil(KP/64) chunks of 128 x 128 B
  const int pchunks = (p.KP + 63) / 64;
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(sP + pchunks * 16384);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_mma + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;  // query row inside the tile == TMEM lane
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS = tmem_base + ((uint32_t)(warp * 32) << 16);        // scores: columns [0, KP)
  const uint32_t tO = tS + 256;                                          // output: columns [256, 320)
  const uint32_t idesc_s = umma_idesc_bf16(128, p.KP);
  const uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);        // B (= V) is MN-major
  const uint32_t swz = (uint32_t)(row & 7);

  const long long items = (long long)p.n_img * p.heads * p.mtiles;
  uint32_t ph_load = 0, ph_mma = 0;
  for (long long it = blockIdx.x; it < items; it += gridDim.x) {
    const int mt = (int)(it % p.mtiles);
    const int head = (int)((it / p.mtiles) % p.heads);
    const long long img = it / ((long long)p.mtiles * p.heads);
    const int row0 = (int)(img * p.S);
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar_load, 16384 + 2 * kv_bytes);
      tma_load_2d(&tmap_q, bar_load, sQ, head * 64, row0 + mt * 128);
      tma_load_2d(&tmap_kv, bar_load, sK, p.W + head * 64, row0);
      tma_load_2d(&tmap_kv, bar_load, sV, 2 * p.W + head * 64, row0);
    }
    mbar_wait(bar_load, ph_load);
    ph_load ^= 1;
    if (threadIdx.x == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_base, umma_desc_sw128(smem_u32(sQ) + k * 32), umma_desc_sw128(smem_u32(sK) + k * 32), idesc_s,
                  k != 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    // ---- softmax over the valid keys of this thread's row
    float mx = -INFINITY;
    for (int c = 0; c < p.KP; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tS + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (c + j < p.S) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    float sum = 0.f;
    const uint32_t p_row = smem_u32(sP) + row * 128;
    for (int c = 0; c < p.KP; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tS + c, v);
      tmem_wait_ld();
      float e[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        e[j] = (c + j < p.S) ? exp2f((__uint_as_float(v[j]) - mx) * p.scale_log2e) : 0.f;
        // the probabilities enter P V as bf16: normalise with the sum of the rounded values
        e[j] = __bfloat162float(__float2bfloat16_rn(e[j]));
        sum += e[j];
      }
      const uint32_t chunk_base = p_row + (c >> 6) * 16384;
      const int k16 = (c & 63) >> 3;  // 16-byte chunk index inside the 128-byte row (two per 16 keys)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 o;
        __nv_bfloat162 a = __floats2bfloat162_rn(e[8 * h + 0], e[8 * h + 1]), b = __floats2bfloat162_rn(e[8 * h + 2], e[8 * h + 3]);
        __nv_bfloat162 cc = __floats2bfloat162_rn(e[8 * h + 4], e[8 * h + 5]), d = __floats2bfloat162_rn(e[8 * h + 6], e[8 * h + 7]);
        o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        o.z = *reinterpret_cast<uint32_t*>(&cc); o.w = *reinterpret_cast<uint32_t*>(&d);
        st_shared_v4(chunk_base + (((k16 + h) ^ swz) << 4), o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int ks = 0; ks < p.KP / 16; ++ks)
        umma_bf16(tmem_base + 256, umma_desc_sw128(smem_u32(sP) + (ks >> 2) * 16384 + (ks & 3) * 32),
                  umma_desc_sw128(smem_u32(sV) + ks * 2048), idesc_o, ks != 0);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1;
    tc_fence_after();
    {
      uint32_t v[64];
      tmem_ld_32x32b_x32(tO, v);
      tmem_ld_32x32b_x32(tO + 32, v + 32);
      tmem_wait_ld();
      const int tok = mt * 128 + row;
      if (tok < p.S) {
        const float inv = 1.f / sum;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((long long)row0 + tok) * p.W + head * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
          __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
          __nv_bfloat162 c = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
          __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
          o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
          o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
          dst[j] = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // smem / TMEM are reused by the next item
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace pvr

extern "C" int pvr_layernorm(const float* x, int64_t row_step, int64_t rows, int width, const float* gamma,
                             const float* beta, float eps, void* y_bf16, void* stream) {
  if (!x || !gamma || !beta || !y_bf16 || rows <= 0 || width != 768 || row_step <= 0) {
    pvr_set_error("pvr_layernorm: invalid argument (width must be 768)");
    return PVR_ERR_ARG;
  }
  pvr::vit_layernorm_kernel<768><<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, row_step, nullptr, nullptr, nullptr, 1, gamma, beta, eps, rows, nullptr, static_cast<__nv_bfloat16*>(y_bf16),
      0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_layernorm: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_vit_embed(const void* patches_bf16, const float* cls, const float* pos, int n_img, int tokens,
                             int width, const float* gamma, const float* beta, float eps, float* x_out, void* stream) {
  if (!patches_bf16 || !cls || !pos || !gamma || !beta || !x_out || n_img <= 0 || tokens <= 1 || width != 768) {
    pvr_set_error("pvr_vit_embed: invalid argument (width must be 768)");
    return PVR_ERR_ARG;
  }
  const long long rows = (long long)n_img * tokens;
  pvr::vit_layernorm_kernel<768><<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, 1, static_cast<const __nv_bfloat16*>(patches_bf16), cls, pos, tokens, gamma, beta, eps, rows, x_out,
      nullptr, 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_vit_embed: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_attention(const void* qkv_bf16, int n_img, int tokens, int width, int heads, void* out_bf16,
                             void* stream) {
  using namespace pvr;
  if (!qkv_bf16 || !out_bf16 || n_img <= 0 || tokens <= 0 || tokens > 256 || heads <= 0 || width != heads * 64) {
    pvr_set_error("pvr_attention: invalid argument (head_dim must be 64, at most 256 tokens)");
    return PVR_ERR_ARG;
  }
  AttnParams p;
  p.n_img = n_img; p.S = tokens; p.W = width; p.heads = heads;
  p.KP = (tokens + 15) / 16 * 16;
  p.mtiles = (tokens + 127) / 128;
  p.scale_log2e = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 folded into the exp2
  p.out = static_cast<__nv_bfloat16*>(out_bf16);
  CUtensorMap tq, tkv;
  const char* err = "";
  const uint64_t rows = (uint64_t)n_img * tokens;
  if (!make_tmap_2d(&tq, qkv_bf16, 3ull * width, rows, 3ull * width, 128, &err) ||
      !make_tmap_2d(&tkv, qkv_bf16, 3ull * width, rows, 3ull * width, (uint32_t)p.KP, &err)) {
    pvr_set_error("pvr_attention: %s", err);
    return PVR_ERR_CUDA;
  }
  const uint32_t kv_alloc = ((uint32_t)p.KP * 128 + 1023) & ~1023u;
  const size_t smem = 1024 + 16384 + 2 * kv_alloc + (size_t)((p.KP + 63) / 64) * 16384 + 64;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pvr_set_error("pvr_attention: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    configured = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)n_img * heads * p.mtiles;
  const int grid = (int)(items < sms ? items : sms);
  vit_attention_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(tq, tkv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_attention: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}
