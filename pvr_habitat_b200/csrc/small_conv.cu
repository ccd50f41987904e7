// First layer of the 5-layer "small conv" trunk (src/embeddings.py:90-106 'random' PVR, src/models.py:107-118
// PolicyNetWithConv.feat_extract): Conv2d(3, 32, 3x3, stride 2, padding 1) + ELU over NHWC4 bf16 frames, and its
// weight / bias gradient.
//
// With 3 (+1 pad) input channels this layer is not a tcgen05 problem: K = 27, and the implicit-GEMM kernel needs
// eight 16-byte-pixel im2col TMA instructions per 128-pixel tile — the SM's TMA pipe serves one instruction per ~700
// cycles, 5.4 k cycles per tile, 493 us for the 3200 frames of a finetune step (46 TFLOP/s, 0.85 TB/s). The layer
// moves 105 MB in and 210 MB out (HBM roof ~ 50 us; the weight gradient reads 735 MB, ~ 115 us). A first CUDA-core
// version (fp32 FMA from shared-memory weights) measured 386 us forward / 1043 us weight gradient
// (profiles/r02_launches_finetune_direct_v0.csv): shared-memory and latency bound. These kernels keep the operands in
// registers and issue warp-level mma.sync m16n8k16 (bf16 x bf16 -> fp32) instead, which takes the arithmetic off the
// critical path: measured 155 us forward (issue bound at IPC 2.2) / 171 us weight gradient (4.3 TB/s),
// profiles/r02_ncu_full_small_conv.txt. Same arithmetic as the GEMM path they replace: bf16 inputs and weights, fp32
// accumulation, fp32 bias, ELU (ex2-based, see elu_neg), bf16 output; the backward rounds dz = dy * ELU'(y) to bf16
// before it is multiplied, as the GEMM path's operand was.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem));
}
// 8-byte copy; src_bytes = 0 writes zeros (padding pixels)
__device__ __forceinline__ void cp_async_8z(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ELU for v <= 0: expm1f costs ~35 instructions per value and was the largest item of the first version (the kernel is
// issue bound: ncu, profiles/r02_ncu_full_small_conv.txt). The result is rounded to bf16 (relative 2^-9), so
// ex2.approx(v log2 e) - 1 is exact for this purpose: its absolute error (~1.2e-7) is below half a bf16 ulp of the
// result for |v| > 6e-5 and below 1.2e-7 in absolute terms everywhere.
__device__ __forceinline__ float elu_neg(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return e - 1.f;
}

// ---------------------------------------------------------------------------------------------- forward
// A warp computes 16 consecutive output pixels x 32 channels as three m16n8k16 steps (one per filter row r): the K
// index inside a step is 4 * slot + c with slot = input column 2 ox - 1 + slot (slot 3 carries zero weights), so every
// A register is one aligned 4-byte load of two channels of one input pixel, straight from global memory (the 16 pixels
// of a warp cover 264 contiguous bytes per input row; L1 serves the overlap). The fragments of the NEXT group are
// loaded before the current one is multiplied and stored, so a warp always has 12 loads in flight. The weights live in
// registers as B fragments whose columns are permuted (column g of n-tile nt = channel 8 (g / 2) + 2 nt + (g & 1)) so
// that a thread's 8 results of one pixel are the 8 consecutive channels 8 tg .. 8 tg + 7: one 16-byte store per pixel.
// `wpk` is the packed weight of program.pack_first_small_conv: bf16 (32, 64), K index 16 r + 4 + 4 j + c for filter
// tap (row r, column j) and input channel c.
__global__ void __launch_bounds__(256, 3) small_conv1_fwd_kernel(const __nv_bfloat16* __restrict__ x, int F, int Hi, int Wi,
                                                              int Ho, int Wo, const __nv_bfloat16* __restrict__ wpk,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ bias,
                                                              __nv_bfloat16* __restrict__ y) {
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  // B fragments: b[r][nt][0] = W(ch, r, slot tg / 2, c 2 (tg & 1) + {0, 1}), b[r][nt][1] = slot 2 + tg / 2 (3 -> 0)
  uint32_t b[3][4][2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int ch = 8 * (g >> 1) + 2 * nt + (g & 1);
      const int c0 = 2 * (tg & 1);
      const __nv_bfloat16* wr = wpk + ch * 64 + 16 * r + 4;
      const unsigned short z = 0;
      unsigned short w00 = __bfloat16_as_ushort(wr[4 * (tg >> 1) + c0]);
      unsigned short w01 = c0 + 1 < 3 ? __bfloat16_as_ushort(wr[4 * (tg >> 1) + c0 + 1]) : z;
      unsigned short w10 = tg < 2 ? __bfloat16_as_ushort(wr[8 + c0]) : z;
      unsigned short w11 = (tg < 2 && c0 + 1 < 3) ? __bfloat16_as_ushort(wr[8 + c0 + 1]) : z;
      b[r][nt][0] = (uint32_t)w00 | ((uint32_t)w01 << 16);
      b[r][nt][1] = (uint32_t)w10 | ((uint32_t)w11 << 16);
    }
  float sc[8], bi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = __ldg(scale + 8 * tg + k);
    bi[k] = __ldg(bias + 8 * tg + k);
  }
  const int M = F * Ho * Wo;  // < 2^31 (checked by the launcher)
  const int groups = (M + 15) >> 4;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* x32 = reinterpret_cast<const uint32_t*>(x);
  // A fragments of one group: a[r][0..3] = (pixel g, slot tg/2), (pixel g + 8, same), (pixel g, slot 2 + tg/2), (g + 8)
  // Pixel coordinates are carried along instead of divided out: a warp owns a contiguous run of groups.
  struct Pix { int m, ox, oy, f; };
  auto advance = [&](Pix& q, int n) {
    q.m += n;
    q.ox += n;
    while (q.ox >= Wo) {
      q.ox -= Wo;
      if (++q.oy == Ho) { q.oy = 0; ++q.f; }
    }
  };
  auto load_group = [&](Pix q, uint32_t (&a)[3][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ok = q.m < M;
      const int ix0 = 2 * q.ox - 1 + (tg >> 1), ix1 = ix0 + 2;
      const bool c0ok = ok && ix0 >= 0 && ix0 < Wi, c1ok = ok && tg < 2 && ix1 < Wi;
      const uint32_t* row = x32 + ((long long)(q.f * Hi + 2 * q.oy - 1) * Wi) * 2 + (tg & 1);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = 2 * q.oy - 1 + r;
        const bool rok = iy >= 0 && iy < Hi;
        a[r][h] = (rok && c0ok) ? __ldg(row + 2 * ix0) : 0u;
        a[r][2 + h] = (rok && c1ok) ? __ldg(row + 2 * ix1) : 0u;
        row += 2 * Wi;
      }
      advance(q, 8);  // second fragment row: pixel m + 8
    }
  };
  const int per = (groups + warps_total - 1) / warps_total;
  int grp = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * per;
  const int grp_end = min(groups, grp + per);
  Pix cur;
  cur.m = grp * 16 + g;
  cur.ox = cur.m % Wo;
  cur.oy = (cur.m / Wo) % Ho;
  cur.f = cur.m / (Wo * Ho);
  uint32_t a[3][4];
  if (grp < grp_end) load_group(cur, a);
  for (; grp < grp_end; ++grp) {
    uint32_t an[3][4];
    const bool more = grp + 1 < grp_end;
    const int m_cur = cur.m;
    if (more) {
      advance(cur, 16);
      load_group(cur, an);
    }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[nt], a[r], b[r][nt][0], b[r][nt][1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m_cur + 8 * h;
      if (m >= M) continue;
      float o[8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float v = fmaf(acc[nt][2 * h + e], sc[2 * nt + e], bi[2 * nt + e]);
          o[2 * nt + e] = v > 0.f ? v : elu_neg(v);
        }
      uint4 pk;
      pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
      *reinterpret_cast<uint4*>(y + (long long)m * 32 + 8 * tg) = pk;
    }
    if (more) {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) a[r][e] = an[r][e];
    }
  }
}

// ---------------------------------------------------------------------------------------------- weight gradient
// dW[co][a][b][c] += sum_px dz[px][co] * x[2 oy - 1 + a][2 ox - 1 + b][c],  dbias[co] += sum_px dz[px][co],
// dz = bf16(dy * (y > 0 ? 1 : y + 1)): a (48 x 32) = col^T (48 x px) . dz (px x 32) product whose reduction runs over
// the 3.3 M output pixels of a finetune step. A block walks tiles of 128 pixels through a two-stage cp.async pipeline:
// raw dy (fp32) and y rows plus the 3 x 3 input pixels around every output pixel (K index 16 a + 4 slot + c, slot 3
// stays zero) land in shared memory while the previous tile is multiplied. dz is formed shared -> shared. Both operands
// are pixel-major, the TRANSPOSE of what mma.sync wants for a reduction over pixels, so the fragments come from
// ldmatrix.trans. Each of the 8 warps multiplies its own 16 pixels (12 m16n8k16 per tile) into 48 fp32 accumulators
// kept for the whole kernel; the blocks add their sums into dw / dbias once at the end.
constexpr int WG_TILE = 128;   // pixels per tile
constexpr int WG_COLP = 56;    // bf16 per im2col row in shared memory (48 used; 112-byte rows: conflict-free ldmatrix)
constexpr int WG_DZP = 40;     // bf16 per dz row (32 used; 80-byte rows)
constexpr int WG_STAGE = WG_TILE * 32 * 4 + WG_TILE * 32 * 2 + WG_TILE * WG_COLP * 2;  // dy | y | cols = 38912 B
constexpr int WG_SMEM = 2 * WG_STAGE + WG_TILE * WG_DZP * 2 + (48 * 32 + 32) * 4;     // 94336 B

__global__ void __launch_bounds__(256, 2) small_conv1_wgrad_kernel(const float* __restrict__ dy,
                                                                   const __nv_bfloat16* __restrict__ y, int y_pitch,
                                                                   const __nv_bfloat16* __restrict__ x, int F, int Hi,
                                                                   int Wi, int Ho, int Wo, float* __restrict__ dw,
                                                                   float* __restrict__ dbias) {
  extern __shared__ __align__(16) uint8_t wg_smem[];
  __nv_bfloat16* dzs = reinterpret_cast<__nv_bfloat16*>(wg_smem + 2 * WG_STAGE);
  float* red = reinterpret_cast<float*>(wg_smem + 2 * WG_STAGE + WG_TILE * WG_DZP * 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int M = F * Ho * Wo;
  const int tiles = (M + WG_TILE - 1) / WG_TILE;
  float acc[3][4][4];
#pragma unroll
  for (int mt = 0; mt < 3; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
  float bacc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) bacc[k] = 0.f;
  for (int i = tid; i < 48 * 32 + 32; i += 256) red[i] = 0.f;
  // everything the copies never write (slot 3 of every filter row, the row padding) has to be finite: zero both stages
  for (int i = tid; i < 2 * WG_STAGE / 16; i += 256) reinterpret_cast<uint4*>(wg_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  auto issue = [&](int tile, int st) {
    uint8_t* base = wg_smem + st * WG_STAGE;
    float* sdy = reinterpret_cast<float*>(base);
    __nv_bfloat16* sy = reinterpret_cast<__nv_bfloat16*>(base + WG_TILE * 128);
    __nv_bfloat16* cols = reinterpret_cast<__nv_bfloat16*>(base + WG_TILE * 192);
    const int m0 = tile * WG_TILE;
    // dy: 8 x 16 B per pixel, y: 4 x 16 B per pixel. Pixels past M are clamped (their dz is forced to zero below).
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int i = tid + 256 * it, p = i >> 3, ch = i & 7;
      const int m = min(m0 + p, M - 1);
      cp_async_16(sdy + p * 32 + 4 * ch, dy + (long long)m * 32 + 4 * ch);
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int i = tid + 256 * it, p = i >> 2, ch = i & 3;
      const int m = min(m0 + p, M - 1);
      cp_async_16(sy + p * 32 + 8 * ch, y + (long long)m * y_pitch + 8 * ch);
    }
    // im2col: two threads per pixel, 9 input pixels of 8 bytes split 5 / 4 (taps 0-4 and 5-8, tap = 3 r + slot)
    {
      const int p = tid >> 1, half = tid & 1;
      const int m = min(m0 + p, M - 1);
      const int ox = m % Wo, t = m / Wo;
      const int oy = t % Ho, f = t / Ho;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const int tap = 5 * half + k;
        if (tap < 9) {
          const int r = tap / 3, s = tap - 3 * r;
          const int iy = 2 * oy - 1 + r, ix = 2 * ox - 1 + s;
          const bool ok = iy >= 0 && iy < Hi && ix >= 0 && ix < Wi;
          const __nv_bfloat16* src = ok ? x + (((long long)f * Hi + iy) * Wi + ix) * 4 : x;
          cp_async_8z(cols + p * WG_COLP + 16 * r + 4 * s, src, ok ? 8 : 0);
        }
      }
    }
    cp_async_commit();
  };

  int st = 0;
  if ((int)blockIdx.x < tiles) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, st ^= 1) {
    const int nxt = tile + gridDim.x;
    if (nxt < tiles) {
      issue(nxt, st ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    uint8_t* base = wg_smem + st * WG_STAGE;
    const float* sdy = reinterpret_cast<const float*>(base);
    const __nv_bfloat16* sy = reinterpret_cast<const __nv_bfloat16*>(base + WG_TILE * 128);
    const __nv_bfloat16* cols = reinterpret_cast<const __nv_bfloat16*>(base + WG_TILE * 192);
    const int m0 = tile * WG_TILE;
    // dz: item = (pixel, 8-channel quarter); a thread keeps the same quarter (tid & 3) for the bias sums
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int i = tid + 256 * it;
      const int p = i >> 2, q = i & 3;
      uint4 pk = make_uint4(0u, 0u, 0u, 0u);
      if (m0 + p < M) {
        const float4 d0 = *reinterpret_cast<const float4*>(sdy + p * 32 + 8 * q);
        const float4 d1 = *reinterpret_cast<const float4*>(sdy + p * 32 + 8 * q + 4);
        const uint4 yv = *reinterpret_cast<const uint4*>(sy + p * 32 + 8 * q);
        const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const float yy[8] = {bf_lo(yv.x), bf_hi(yv.x), bf_lo(yv.y), bf_hi(yv.y),
                             bf_lo(yv.z), bf_hi(yv.z), bf_lo(yv.w), bf_hi(yv.w)};
        float dz[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          dz[k] = __bfloat162float(__float2bfloat16_rn(dd[k] * (yy[k] > 0.f ? 1.f : yy[k] + 1.f)));
          bacc[k] += dz[k];
        }
        pk.x = pack2(dz[0], dz[1]); pk.y = pack2(dz[2], dz[3]); pk.z = pack2(dz[4], dz[5]); pk.w = pack2(dz[6], dz[7]);
      }
      *reinterpret_cast<uint4*>(dzs + p * WG_DZP + 8 * q) = pk;
    }
    __syncthreads();
    {
      const int p0 = warp * 16;  // this warp's 16 pixels = one k16 step
      const int lr = lane & 7, mi = lane >> 3;
      // B = dz (k = pixel, n = channel): matrices (k half mi & 1, n-tile 2 j + (mi >> 1))
      uint32_t bf[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        ldmatrix_x4_trans(bf[j], dzs + (p0 + 8 * (mi & 1) + lr) * WG_DZP + 16 * j + 8 * (mi >> 1));
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        // A = col^T (m = K index, k = pixel): a0 (m 0-7, k 0-7), a1 (m 8-15, k 0-7), a2 (m 0-7, k 8-15), a3 (m 8-15, k 8-15)
        uint32_t af[4];
        ldmatrix_x4_trans(af, cols + (p0 + 8 * (mi >> 1) + lr) * WG_COLP + 16 * mt + 8 * (mi & 1));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_bf16_16816(acc[mt][nt], af, bf[nt >> 1][2 * (nt & 1)], bf[nt >> 1][2 * (nt & 1) + 1]);
      }
    }
    __syncthreads();  // dzs and this stage are rewritten by the next iteration
  }
  // block sums: shared-memory atomics over the 8 warps, then one global atomic per element
  {
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kidx = 16 * mt + g + 8 * (e >> 1), co = 8 * nt + 2 * tg + (e & 1);
          atomicAdd(&red[kidx * 32 + co], acc[mt][nt][e]);
        }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = bacc[k];  // lanes with the same (lane & 3) hold the same channels
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 4) atomicAdd(&red[48 * 32 + 8 * lane + k], v);
    }
  }
  __syncthreads();
  for (int i = tid; i < 48 * 32; i += 256) {
    const int kidx = i >> 5, co = i & 31;
    const int a = kidx >> 4, s = (kidx >> 2) & 3, c = kidx & 3;
    if (s < 3 && c < 3) atomicAdd(dw + ((co * 3 + a) * 3 + s) * 4 + c, red[i]);
  }
  if (tid < 32) atomicAdd(dbias + tid, red[48 * 32 + tid]);
}

}  // namespace

cudaError_t launch_small_conv1(const void* x, int F, int Hi, int Wi, int Ho, int Wo, const void* wpk, const float* scale,
                               const float* bias, void* y, cudaStream_t stream) {
  const long long M = (long long)F * Ho * Wo;
  if (M <= 0 || M > 0x7fffffffll - 16 || (long long)F * Hi * Wi > 0x7fffffffll) return cudaErrorInvalidValue;
  long long blocks = ((M + 15) / 16 + 7) / 8;  // 8 warps of 16 pixels per block, grid-stride over the rest
  if (blocks > 148 * 8) blocks = 148 * 8;
  small_conv1_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(x), F, Hi, Wi, Ho, Wo, static_cast<const __nv_bfloat16*>(wpk), scale, bias,
      static_cast<__nv_bfloat16*>(y));
  return cudaGetLastError();
}

}  // namespace pvr

extern "C" int pvr_small_conv1_wgrad(const float* dy, const void* y_bf16, int y_pitch, const void* x_nhwc4_bf16, int F,
                                     int Hi, int Wi, int Ho, int Wo, float* dw, float* dbias, void* stream) {
  if (!dy || !y_bf16 || !x_nhwc4_bf16 || !dw || !dbias || F <= 0 || Hi <= 0 || Wi <= 0 || Ho != (Hi - 1) / 2 + 1 ||
      Wo != (Wi - 1) / 2 + 1 || y_pitch < 32 || y_pitch % 8 ||
      ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y_bf16)) & 15) ||
      (reinterpret_cast<uintptr_t>(x_nhwc4_bf16) & 7)) {
    pvr_set_error("pvr_small_conv1_wgrad: invalid argument");
    return PVR_ERR_ARG;
  }
  const long long M = (long long)F * Ho * Wo;
  if (M > 0x7fffffffll - 128 || (long long)F * Hi * Wi > 0x7fffffffll) {
    pvr_set_error("pvr_small_conv1_wgrad: more than 2^31 pixels");
    return PVR_ERR_ARG;
  }
  long long blocks = (M + 127) / 128;
  if (blocks > 148 * 2) blocks = 148 * 2;
  static bool attr_set = false;  // benign race: the attribute is idempotent
  if (!attr_set) {
    const cudaError_t ea = cudaFuncSetAttribute(pvr::small_conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                pvr::WG_SMEM);
    if (ea != cudaSuccess) {
      pvr_set_error("pvr_small_conv1_wgrad: %s", cudaGetErrorString(ea));
      return PVR_ERR_CUDA;
    }
    attr_set = true;
  }
  pvr::small_conv1_wgrad_kernel<<<(unsigned)blocks, 256, pvr::WG_SMEM, static_cast<cudaStream_t>(stream)>>>(
      dy, static_cast<const __nv_bfloat16*>(y_bf16), y_pitch, static_cast<const __nv_bfloat16*>(x_nhwc4_bf16), F, Hi,
      Wi, Ho, Wo, dw, dbias);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pvr_set_error("pvr_small_conv1_wgrad: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}
