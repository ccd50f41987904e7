// Multi-head self-attention for the ViT shapes the tcgen05 kernel of vit.cu does not cover: head_dim != 64 or more than
// 256 tokens. The one user today is `mae_huge` (src/embeddings.py:145-148 -> mae_vit_huge_patch14,
// src/vision_models/mae.py:291-296: width 1280, 16 heads of 80, 16 x 16 patches + class token = 257 tokens), where
// attention is 4 % of the encoder's FLOPs — so this is the compact register-level formulation (warp-level
// mma.sync m16n8k16 bf16 -> fp32, online softmax), not a second tensor-memory pipeline:
//   one block per (image, head): K and V of the head are staged once in shared memory (cp.async, rows padded by 16
//   bytes so that every ldmatrix phase touches 8 distinct 16-byte bank groups); every warp owns 16 query rows at a time,
//   keeps its Q fragments in registers, and walks the keys 32 at a time:
//     S = Q K^T          B fragments = K rows via ldmatrix (K is (key, d) row-major = column-major B)
//     online softmax     running row max / sum in registers, exp2 with head_dim^-0.5 * log2(e) folded in
//     O += P V           P re-used from the S accumulators as A fragments (bf16), V via ldmatrix.trans
// Probabilities enter P V as bf16 and the normaliser is the sum of the ROUNDED values, like vit_attention_kernel.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b, float& sum) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  sum += __low2float(t) + __high2float(t);
  return *reinterpret_cast<const uint32_t*>(&t);
}

template <int D>
__global__ void __launch_bounds__(256, 2) vit_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int S, int W,
                                                                 int heads, float scale_log2e,
                                                                 __nv_bfloat16* __restrict__ out) {
  static_assert(D % 16 == 0 && D <= 128, "head_dim");
  constexpr int PITCH = D + 8;  // bf16 per shared-memory row
  constexpr int KS = D / 16;    // k16 steps of Q K^T
  constexpr int DN = D / 8;     // n8 tiles of the output
  extern __shared__ __align__(16) uint8_t att_smem[];
  const int SP = (S + 31) & ~31;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* Vs = Ks + (size_t)SP * PITCH;
  const int img = blockIdx.x / heads, head = blockIdx.x - img * heads;
  const __nv_bfloat16* base = qkv + (long long)img * S * 3 * W + head * D;
  for (int i = threadIdx.x; i < SP * (D / 8); i += blockDim.x) {
    const int t = i / (D / 8), c = i - t * (D / 8);
    if (t < S) {
      cp16(Ks + t * PITCH + 8 * c, base + (long long)t * 3 * W + W + 8 * c);
      cp16(Vs + t * PITCH + 8 * c, base + (long long)t * 3 * W + 2 * W + 8 * c);
    } else {  // keys past the sequence: masked below, but V must be finite (0 * NaN)
      *reinterpret_cast<uint4*>(Ks + t * PITCH + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(Vs + t * PITCH + 8 * c) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, tg = lane & 3, lr = lane & 7, mi = lane >> 3;
  const int rtiles = (S + 15) >> 4;
  for (int rt = warp; rt < rtiles; rt += nwarps) {
    const int r0 = rt * 16 + g, r1 = r0 + 8;
    uint32_t qf[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const __nv_bfloat16* q0 = base + (long long)r0 * 3 * W + 16 * ks + 2 * tg;
      const __nv_bfloat16* q1 = base + (long long)r1 * 3 * W + 16 * ks + 2 * tg;
      qf[ks][0] = r0 < S ? *reinterpret_cast<const uint32_t*>(q0) : 0u;
      qf[ks][1] = r1 < S ? *reinterpret_cast<const uint32_t*>(q1) : 0u;
      qf[ks][2] = r0 < S ? *reinterpret_cast<const uint32_t*>(q0 + 8) : 0u;
      qf[ks][3] = r1 < S ? *reinterpret_cast<const uint32_t*>(q1 + 8) : 0u;
    }
    float o[DN][4];
#pragma unroll
    for (int dn = 0; dn < DN; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dn][e] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lsum[2] = {0.f, 0.f};
    for (int kb = 0; kb < SP; kb += 32) {
      float s[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int ntp = 0; ntp < 2; ++ntp) {
          // matrices: (keys +0..7, d lo), (keys +0..7, d hi), (keys +8..15, d lo), (keys +8..15, d hi)
          uint32_t kf[4];
          ldsm_x4(kf, Ks + (kb + 16 * ntp + 8 * (mi >> 1) + lr) * PITCH + 16 * ks + 8 * (mi & 1));
          mma_16816(s[2 * ntp], qf[ks], kf[0], kf[1]);
          mma_16816(s[2 * ntp + 1], qf[ks], kf[2], kf[3]);
        }
      }
      // s[nt][e]: row g (e < 2) / g + 8 (e >= 2), key kb + 8 nt + 2 tg + (e & 1)
      if (kb + 32 > S) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (kb + 8 * nt + 2 * tg + (e & 1) >= S) s[nt][e] = -INFINITY;
      }
      uint32_t pa[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float mx = mrow[h];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mx = fmaxf(mx, fmaxf(s[nt][2 * h], s[nt][2 * h + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        // every 32-key block holds at least one real key (SP - S < 32), so mx is finite from the first block on
        const float corr = ex2((mrow[h] - mx) * scale_log2e);
        mrow[h] = mx;
        const float mxs = mx * scale_log2e;
        float part = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float p0 = ex2(fmaf(s[nt][2 * h], scale_log2e, -mxs));
          const float p1 = ex2(fmaf(s[nt][2 * h + 1], scale_log2e, -mxs));
          // A fragment of k16 step nt / 2: register h (+ 2 for the upper 8 keys of the step)
          pa[nt >> 1][h + 2 * (nt & 1)] = pack_bf16(p0, p1, part);
        }
        lsum[h] = lsum[h] * corr + part;
#pragma unroll
        for (int dn = 0; dn < DN; ++dn) {
          o[dn][2 * h] *= corr;
          o[dn][2 * h + 1] *= corr;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
        for (int dp = 0; dp < DN / 2; ++dp) {
          // matrices: (keys +0..7, d tile 2 dp), (keys +8..15, d tile 2 dp), (keys +0..7, 2 dp + 1), (keys +8..15, 2 dp + 1)
          uint32_t vf[4];
          ldsm_x4_t(vf, Vs + (kb + 16 * kk + 8 * (mi & 1) + lr) * PITCH + 8 * (2 * dp + (mi >> 1)));
          mma_16816(o[2 * dp], pa[kk], vf[0], vf[1]);
          mma_16816(o[2 * dp + 1], pa[kk], vf[2], vf[3]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float l = lsum[h];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const int r = h ? r1 : r0;
      if (r >= S) continue;
      const float inv = 1.f / l;
      __nv_bfloat16* dst = out + ((long long)img * S + r) * W + head * D + 2 * tg;
#pragma unroll
      for (int dn = 0; dn < DN; ++dn)
        *reinterpret_cast<__nv_bfloat162*>(dst + 8 * dn) = __floats2bfloat162_rn(o[dn][2 * h] * inv, o[dn][2 * h + 1] * inv);
    }
  }
}

template <int D>
int launch(const void* qkv, int n_img, int tokens, int width, int heads, void* out, cudaStream_t stream) {
  const int SP = (tokens + 31) & ~31;
  const size_t smem = (size_t)SP * (D + 8) * 2 * 2;
  if (smem > 200 * 1024) {
    pvr_set_error("pvr_attention_mma: %d tokens of head_dim %d do not fit shared memory", tokens, D);
    return PVR_ERR_ARG;
  }
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(vit_attention_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    // two blocks per SM (2 x 101 KB at mae_huge's shape) need the full shared-memory carve-out: with the default one
    // the driver sized it for ONE block (ncu: "Block Limit Shared Mem 1", 12 % occupancy, 68 % of the issue slots idle)
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vit_attention_mma_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { pvr_set_error("pvr_attention_mma: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
    configured = smem;
  }
  // as many warps (<= 8) as keep the 16-row tiles evenly spread: 257 tokens = 17 tiles -> 6 warps x 3 rounds. (Registers
  // are allocated per 4 warps: 9 warps x 2 rounds was tried and is slower, 587 vs 467 us per layer of mae_huge at 256
  // images, because it costs the second block per SM — profiles/r02_ncu_full_attention_mma.txt.)
  const int rtiles = (tokens + 15) / 16;
  const int rounds = (rtiles + 7) / 8;
  const int warps = (rtiles + rounds - 1) / rounds;
  vit_attention_mma_kernel<D><<<n_img * heads, 32 * warps, smem, stream>>>(
      static_cast<const __nv_bfloat16*>(qkv), tokens, width, heads, 1.4426950408889634f / sqrtf((float)D),
      static_cast<__nv_bfloat16*>(out));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { pvr_set_error("pvr_attention_mma: %s", cudaGetErrorString(e)); return PVR_ERR_CUDA; }
  return PVR_OK;
}

}  // namespace
}  // namespace pvr

extern "C" int pvr_attention_mma(const void* qkv_bf16, int n_img, int tokens, int width, int heads, void* out_bf16,
                                 void* stream) {
  if (!qkv_bf16 || !out_bf16 || n_img <= 0 || tokens <= 0 || heads <= 0 || width <= 0 || width % heads ||
      (long long)n_img * heads > 0x7fffffffll || ((reinterpret_cast<uintptr_t>(qkv_bf16) | (uintptr_t)width * 2) & 15)) {
    pvr_set_error("pvr_attention_mma: invalid argument");
    return PVR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (width / heads) {
    case 64: return pvr::launch<64>(qkv_bf16, n_img, tokens, width, heads, out_bf16, st);
    case 80: return pvr::launch<80>(qkv_bf16, n_img, tokens, width, heads, out_bf16, st);
    case 96: return pvr::launch<96>(qkv_bf16, n_img, tokens, width, heads, out_bf16, st);
    case 128: return pvr::launch<128>(qkv_bf16, n_img, tokens, width, heads, out_bf16, st);
    default:
      pvr_set_error("pvr_attention_mma: head_dim %d is not built (64, 80, 96, 128)", width / heads);
      return PVR_ERR_ARG;
  }
}
