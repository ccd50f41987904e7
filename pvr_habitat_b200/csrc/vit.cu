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

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per token row. W % 128 == 0 (each lane owns W/32 values as float4 groups).
// mode 0: x = LN(src fp32 row)                      -> bf16 (ln_1 / ln_2 / ln_post)
// mode 1: x = LN([cls | patch] + pos) (ln_pre)      -> fp32 residual stream
// mode 2: x = [cls | patch] + pos                    -> fp32 residual stream (MAE, mae.py:209-217: no ln_pre)
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
  if (mode == 2) {  // MAE: [cls | patch] + pos goes to the residual stream as it is (no ln_pre)
#pragma unroll
    for (int i = 0; i < PER / 4; ++i)
      *reinterpret_cast<float4*>(out_f32 + row * W + (i * 32 + lane) * 4) =
          make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
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
// One work item = (image, head): both 128-row query tiles of the sequence (<= 256 tokens), head_dim = 64.
//   S = Q K^T      tcgen05.mma, Q (128 x 64 per tile) and K (KP x 64) K-major from TMA, fp32 scores in TMEM
//   P = softmax    one thread per query row reads its TMEM lane (row max / sum are thread-local) and writes the
//                  probabilities back INTO TMEM as packed bf16 (tcgen05.st), over the scores it has consumed
//   O = P V        tcgen05.mma with the A operand in TMEM; V (KP x 64, straight from TMA) is the MN-major B operand
// Warps 0-3 / 4-7 are two softmax groups, each owning one 256-column TMEM buffer; query tiles alternate between them
// and are pipelined individually (see the controller); warp 8 is the controller (TMA prefetch two items ahead, MMA
// issue). Keys beyond the sequence get p = 0; query rows beyond it are computed but never stored.
// TMEM columns per tile (base 0 / 256): S [0, KP), P [0, KP/2) in place, O [128, 192).
struct AttnParams {
  int n_img, S, W, heads, KP, mtiles;
  float scale_log2e;
  __nv_bfloat16* out;
};

constexpr int ATT_THREADS = 288;

__global__ void __launch_bounds__(ATT_THREADS, 1)
vit_attention_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t kv_bytes = (uint32_t)p.KP * 128;
  const uint32_t kv_alloc = (kv_bytes + 1023) & ~1023u;
  const uint32_t stage_bytes = 32768 + 2 * kv_alloc;  // Q (2 tiles) | K | V
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* s_ready = bars + 4;     // [2] one per TMEM buffer (256 columns each)
  uint64_t* p_ready = bars + 6;     // [2]
  uint64_t* o_ready = bars + 8;     // [2]
  uint64_t* tmem_free = bars + 10;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_q);
    prefetch_tmap(&tmap_kv);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_ready[i], 1);
      mbar_init(&tmem_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);  // uniform register for tcgen05.mma
  const long long items = (long long)p.n_img * p.heads;
  const uint32_t q_bytes = 16384u * p.mtiles;

  if (warp == 8) {
    // ===================================================== controller: TMA prefetch + MMA issue. The whole warp walks
    // the loop (uniform loop state), one elected lane issues; descriptors are built once per stage and advanced by
    // adding to the (address >> 4) field (see conv_gemm.cu).
    const uint32_t idesc_s = umma_idesc_bf16(128, p.KP);
    const uint32_t idesc_o = umma_idesc_bf16(128, 64) | (1u << 16);  // B (= V) is MN-major
    auto issue_loads = [&](long long it, int st) {
      const int head = (int)(it % p.heads);
      const int row0 = (int)((it / p.heads) * p.S);
      uint8_t* base = smem + st * stage_bytes;
      mbar_expect_tx(&kv_full[st], q_bytes + 2 * kv_bytes);
      for (int mt = 0; mt < p.mtiles; ++mt)
        tma_load_2d(&tmap_q, &kv_full[st], base + mt * 16384, head * 64, row0 + mt * 128);
      tma_load_2d(&tmap_kv, &kv_full[st], base + 32768, p.W + head * 64, row0);
      tma_load_2d(&tmap_kv, &kv_full[st], base + 32768 + kv_alloc, 2 * p.W + head * 64, row0);
    };
    // Query tiles are pipelined individually: tile tt (item tt / mtiles, m-tile tt % mtiles) uses TMEM buffer tt & 1
    // and softmax group tt & 1. Per iteration the controller issues S(tt) as soon as that buffer is free and P V of
    // tile tt - 1 as soon as its probabilities are written, so one group's MMAs / epilogue overlap the other group's
    // softmax instead of all warps marching through S -> softmax -> PV -> output in lock step.
    const int mtl = p.mtiles;
    const long long my_items = (long long)blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_tiles = my_items * mtl;
    auto item_of = [&](long long n) { return (long long)blockIdx.x + n * gridDim.x; };
    for (int i = 0; i < 2 && i < my_items; ++i) {
      if (elect_one()) issue_loads(item_of(i), i);
      __syncwarp();
    }
    for (long long tt = 0; tt <= total_tiles; ++tt) {
      if (tt < total_tiles) {
        const long long n = tt / mtl;
        const int j = (int)(tt - n * mtl), st = (int)(n & 1), b = (int)(tt & 1);
        const long long u = tt >> 1;
        if (j == 0) mbar_wait(&kv_full[st], (uint32_t)((n >> 1) & 1));
        if (u >= 1) mbar_wait(&tmem_free[b], (uint32_t)((u - 1) & 1));  // tile tt - 2 has left this buffer
        tc_fence_after();
        const uint32_t sQ = smem_u32(smem + st * stage_bytes);
        const uint64_t q_desc = umma_desc_sw128(sQ + j * 16384), k_desc = umma_desc_sw128(sQ + 32768);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + b * 256, q_desc + ((k * 32) >> 4), k_desc + ((k * 32) >> 4), idesc_s, k != 0);
          umma_commit(&s_ready[b]);
        }
        __syncwarp();
      }
      if (tt >= 1) {
        const long long pt = tt - 1, pn = pt / mtl;
        const int pj = (int)(pt - pn * mtl), pst = (int)(pn & 1), pb = (int)(pt & 1);
        mbar_wait(&p_ready[pb], (uint32_t)((pt >> 1) & 1));
        tc_fence_after();
        const uint64_t v_desc = umma_desc_sw128(smem_u32(smem + pst * stage_bytes) + 32768 + kv_alloc);
        if (elect_one()) {
          for (int ks = 0; ks < p.KP / 16; ++ks)
            umma_bf16_ts(tmem_base + pb * 256 + 128, tmem_base + pb * 256 + ks * 8, v_desc + ((ks * 2048) >> 4), idesc_o,
                         ks != 0);
          umma_commit(&o_ready[pb]);
          if (pj == mtl - 1) umma_commit(&kv_empty[pst]);
        }
        __syncwarp();
        if (pj == mtl - 1 && pn + 2 < my_items) {  // the stage of item pn is free once its MMAs have completed
          mbar_wait(&kv_empty[pst], (uint32_t)((pn >> 1) & 1));
          if (elect_one()) issue_loads(item_of(pn + 2), pst);
          __syncwarp();
        }
      }
    }
  } else if (warp < 8) {
    // ===================================================== softmax + output (one thread per query row)
    const int g = warp >> 2;  // group = TMEM buffer
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tS = tmem_base + g * 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const int mtl = p.mtiles;
    const long long my_items = (long long)blockIdx.x < items ? (items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total_tiles = my_items * mtl;
    for (long long tt = g; tt < total_tiles; tt += 2) {
      const long long nn = tt / mtl;
      const int mt = (int)(tt - nn * mtl);
      const uint32_t n = (uint32_t)(tt >> 1);  // use count of this buffer
      const long long it = (long long)blockIdx.x + nn * gridDim.x;
      const int head = (int)(it % p.heads);
      const long long row0 = (it / p.heads) * p.S;
      mbar_wait(&s_ready[g], n & 1);
      tc_fence_after();
      // Both passes read S in 16-column chunks, double buffered: the tcgen05.ld of the next chunk is in flight while
      // the current one is reduced / exponentiated; only the chunk that crosses S (197 of KP = 208 columns) is masked.
      float mx = -INFINITY;
      {
        uint32_t va[16], vb[16];
        tmem_ld_32x32b_x16(tS, va);
        for (int c = 0; c < p.KP; c += 32) {
          const bool has_b = c + 16 < p.KP;
          tmem_wait_ld();
          if (has_b) tmem_ld_32x32b_x16(tS + c + 16, vb);
          if (c + 16 <= p.S) {
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(va[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < p.S) mx = fmaxf(mx, __uint_as_float(va[j]));
          }
          if (has_b) {
            tmem_wait_ld();
            if (c + 32 < p.KP) tmem_ld_32x32b_x16(tS + c + 32, va);
            if (c + 32 <= p.S) {
#pragma unroll
              for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(vb[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c + 16 + j < p.S) mx = fmaxf(mx, __uint_as_float(vb[j]));
            }
          }
        }
      }
      float sum = 0.f;
      {
        const float mxs = mx * p.scale_log2e;
        auto expo = [&](const uint32_t* v, int c0) {  // 16 columns -> 8 packed bf16 pairs, stored in place
          uint32_t pk[8];
          const bool full = c0 + 16 <= p.S;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float e0 = fast_exp2(fmaf(__uint_as_float(v[2 * j]), p.scale_log2e, -mxs));
            float e1 = fast_exp2(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2e, -mxs));
            if (!full) {
              if (c0 + 2 * j >= p.S) e0 = 0.f;
              if (c0 + 2 * j + 1 >= p.S) e1 = 0.f;
            }
            __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
            // the probabilities enter P V as bf16: normalise with the sum of the rounded values
            sum += __low2float(b2) + __high2float(b2);
            pk[j] = *reinterpret_cast<uint32_t*>(&b2);
          }
          tmem_st_32x32b_x8(tS + (c0 >> 1), pk);  // in place: columns [c0/2, c0/2+8) were consumed in earlier chunks
        };
        uint32_t va[16], vb[16];
        tmem_ld_32x32b_x16(tS, va);
        for (int c = 0; c < p.KP; c += 32) {
          const bool has_b = c + 16 < p.KP;
          tmem_wait_ld();
          if (has_b) tmem_ld_32x32b_x16(tS + c + 16, vb);
          expo(va, c);
          if (has_b) {
            tmem_wait_ld();
            if (c + 32 < p.KP) tmem_ld_32x32b_x16(tS + c + 32, va);
            expo(vb, c + 16);
          }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[g]);
      mbar_wait(&o_ready[g], n & 1);
      tc_fence_after();
      uint32_t o[64];
      tmem_ld_32x32b_x32(tS + 128, o);
      tmem_ld_32x32b_x32(tS + 160, o + 32);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&tmem_free[g]);
      const int tok = mt * 128 + row;
      if (tok < p.S) {
        const float inv = 1.f / sum;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (row0 + tok) * p.W + head * 64);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 ov;
          __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
          __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
          __nv_bfloat162 c = __floats2bfloat162_rn(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
          __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
          ov.x = *reinterpret_cast<uint32_t*>(&a); ov.y = *reinterpret_cast<uint32_t*>(&b);
          ov.z = *reinterpret_cast<uint32_t*>(&c); ov.w = *reinterpret_cast<uint32_t*>(&d);
          dst[j] = ov;
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace pvr

namespace pvr {
namespace {
template <int W>
cudaError_t launch_ln(const float* src, long long row_step, const __nv_bfloat16* patches, const float* cls,
                      const float* pos, int tokens, const float* gamma, const float* beta, float eps, long long rows,
                      float* out_f32, __nv_bfloat16* out_bf16, int mode, cudaStream_t stream) {
  vit_layernorm_kernel<W><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(src, row_step, patches, cls, pos, tokens, gamma,
                                                                          beta, eps, rows, out_f32, out_bf16, mode);
  return cudaGetLastError();
}
// widths: 768 (ViT-B: CLIP, mae_base), 1024 (ViT-L: mae_large), 1280 (ViT-H: mae_huge)
}  // namespace
inline bool ln_width_ok(int width) { return width == 768 || width == 1024 || width == 1280; }
namespace {
cudaError_t dispatch_ln(int width, const float* src, long long row_step, const __nv_bfloat16* patches, const float* cls,
                        const float* pos, int tokens, const float* gamma, const float* beta, float eps, long long rows,
                        float* out_f32, __nv_bfloat16* out_bf16, int mode, cudaStream_t stream) {
  if (width == 768)
    return launch_ln<768>(src, row_step, patches, cls, pos, tokens, gamma, beta, eps, rows, out_f32, out_bf16, mode, stream);
  if (width == 1280)
    return launch_ln<1280>(src, row_step, patches, cls, pos, tokens, gamma, beta, eps, rows, out_f32, out_bf16, mode, stream);
  return launch_ln<1024>(src, row_step, patches, cls, pos, tokens, gamma, beta, eps, rows, out_f32, out_bf16, mode, stream);
}
}  // namespace
}  // namespace pvr

extern "C" int pvr_layernorm(const float* x, int64_t row_step, int64_t rows, int width, const float* gamma,
                             const float* beta, float eps, void* y_bf16, void* stream) {
  if (!x || !gamma || !beta || !y_bf16 || rows <= 0 || !pvr::ln_width_ok(width) || row_step <= 0) {
    pvr_set_error("pvr_layernorm: invalid argument (width must be 768, 1024 or 1280)");
    return PVR_ERR_ARG;
  }
  cudaError_t e = pvr::dispatch_ln(width, x, row_step, nullptr, nullptr, nullptr, 1, gamma, beta, eps, rows, nullptr,
                                   static_cast<__nv_bfloat16*>(y_bf16), 0, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { pvr_set_error("pvr_layernorm: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_layernorm_f32(const float* x, int64_t row_step, int64_t rows, int width, const float* gamma,
                                 const float* beta, float eps, float* y, int64_t ldy, void* stream) {
  if (!x || !gamma || !beta || !y || rows <= 0 || !pvr::ln_width_ok(width) || row_step <= 0 || ldy != width) {
    pvr_set_error("pvr_layernorm_f32: invalid argument (width must be 768, 1024 or 1280, dense output rows)");
    return PVR_ERR_ARG;
  }
  cudaError_t e = pvr::dispatch_ln(width, x, row_step, nullptr, nullptr, nullptr, 1, gamma, beta, eps, rows, y, nullptr,
                                   0, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { pvr_set_error("pvr_layernorm_f32: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_vit_embed(const void* patches_bf16, const float* cls, const float* pos, int n_img, int tokens,
                             int width, const float* gamma, const float* beta, float eps, float* x_out, void* stream) {
  if (!patches_bf16 || !cls || !pos || (!gamma != !beta) || !x_out || n_img <= 0 || tokens <= 1 ||
      !pvr::ln_width_ok(width)) {
    pvr_set_error("pvr_vit_embed: invalid argument (width must be 768, 1024 or 1280)");
    return PVR_ERR_ARG;
  }
  const long long rows = (long long)n_img * tokens;
  cudaError_t e = pvr::dispatch_ln(width, nullptr, 1, static_cast<const __nv_bfloat16*>(patches_bf16), cls, pos, tokens,
                                   gamma, beta, eps, rows, x_out, nullptr, gamma ? 1 : 2,
                                   static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) { pvr_set_error("pvr_vit_embed: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

extern "C" int pvr_attention(const void* qkv_bf16, int n_img, int tokens, int width, int heads, void* out_bf16,
                             void* stream) {
  using namespace pvr;
  if (!qkv_bf16 || !out_bf16 || n_img <= 0 || tokens <= 0 || heads <= 0 || width <= 0 || width % heads) {
    pvr_set_error("pvr_attention: invalid argument");
    return PVR_ERR_ARG;
  }
  // the tensor-memory kernel below is built for head_dim 64 and at most 256 keys (one TMEM buffer per query tile);
  // everything else (mae_huge: 257 tokens, head_dim 80) takes the register-level kernel of attention_mma.cu
  if (tokens > 256 || width != heads * 64)
    return pvr_attention_mma(qkv_bf16, n_img, tokens, width, heads, out_bf16, stream);
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
  const size_t smem = 1024 + 2 * (size_t)(32768 + 2 * kv_alloc) + 128;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(vit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { pvr_set_error("pvr_attention: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    configured = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long items = (long long)n_img * heads;
  const int grid = (int)(items < sms ? items : sms);
  vit_attention_kernel<<<grid, ATT_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(tq, tkv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_attention: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}
