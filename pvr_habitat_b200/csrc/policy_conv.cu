// End-to-end finetuning pieces (PolicyNetWithConv, src/models.py:96-197; main_bc_finetune.py): the 5-layer conv trunk
// runs forward on the tcgen05 conv kernel (program.add_small_conv); its backward is expressed as GEMMs on the same
// tensor-core kernel around these layout kernels:
//   feature gather / scatter   reference feature order `cat([conv(frame_f^T)], -1).view(TB, -1)` (src/models.py:169-170)
//   ELU backward               dZ = dY * (y > 0 ? 1 : y + 1), bf16, 64-column rows for the GEMMs
//   im2col^T                   colT[(r,s,ci)][m] for the weight gradient  dW = dZ^T col      (3x3, stride 2, pad 1)
//   col2im                     dA[pixel] = sum over taps of dcol (input gradient of a layer from dcol = dZ W)
//   BatchNorm1d input gradient dx = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat))
#include "pvr_b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

extern void pvr_set_error(const char* fmt, ...);

namespace {

#define PVR_CHECK_LAUNCH(name)                               \
  do {                                                       \
    cudaError_t e_ = cudaGetLastError();                     \
    if (e_ != cudaSuccess) {                                 \
      pvr_set_error("%s: %s", name, cudaGetErrorString(e_)); \
      return PVR_ERR_CUDA;                                   \
    }                                                        \
  } while (0)

// feat[tb][c*(w*h*N) + x*(h*N) + f*h + y] = Y[(tb*N + f)][y][x][c]   (Y: NHWC bf16 with `pitch` channels per pixel)
__global__ void __launch_bounds__(256) feat_gather_kernel(const __nv_bfloat16* __restrict__ y, int pitch, int TB, int N,
                                                           int h, int w, int C, float* __restrict__ feat) {
  const long long D = (long long)C * w * h * N;
  const long long total = (long long)TB * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tb = i / D;
    long long r = i - tb * D;
    const int c = (int)(r / ((long long)w * h * N));
    r -= (long long)c * w * h * N;
    const int x = (int)(r / (h * N));
    r -= (long long)x * h * N;
    const int f = (int)(r / h), yy = (int)(r - (long long)f * h);
    feat[i] = __bfloat162float(y[(((tb * N + f) * h + yy) * w + x) * pitch + c]);
  }
}

// inverse of the gather for gradients: dY[(tb*N+f)][y][x][c] (fp32, C channels per pixel) = dfeat[tb][...]
__global__ void __launch_bounds__(256) feat_scatter_kernel(const float* __restrict__ dfeat, long long ld, int TB, int N,
                                                            int h, int w, int C, float* __restrict__ dy) {
  const long long per = (long long)N * h * w * C;
  const long long total = (long long)TB * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long tb = i / per;
    long long r = i - tb * per;
    const int f = (int)(r / ((long long)h * w * C));
    r -= (long long)f * h * w * C;
    const int yy = (int)(r / (w * C));
    r -= (long long)yy * w * C;
    const int x = (int)(r / C), c = (int)(r - (long long)x * C);
    dy[i] = dfeat[tb * ld + (long long)c * w * h * N + (long long)x * h * N + (long long)f * h + yy];
  }
}

// dZ[m][c] = dY[m][c] * ELU'(z) with ELU'(z) = 1 (y > 0) or y + 1 (y <= 0); bf16 rows of 64 (columns >= C are zero)
__global__ void __launch_bounds__(256) elu_backward_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                                            int pitch, long long M, int C, __nv_bfloat16* __restrict__ dz) {
  const long long total = M * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i >> 6;
    const int c = (int)(i & 63);
    float v = 0.f;
    if (c < C) {
      const float yy = __bfloat162float(y[m * pitch + c]);
      v = dy[m * C + c] * (yy > 0.f ? 1.f : yy + 1.f);
    }
    dz[i] = __float2bfloat16_rn(v);
  }
}

// The same with everything its consumers need, in one pass over dY (round 2: the three separate kernels — ELU backward,
// pvr_colsum_bf16 for the bias gradient, pvr_transpose_bf16 for the weight-gradient GEMM's K-major operand — took
// 3.2 ms of the 11.6 ms finetuning step, the column sum alone 1.9 ms on the 3.3 M-row first layer):
//   dz  (M, 64) bf16 row-major          operand of the input-gradient GEMM  dcol = dZ W
//   dzt (64, Mp) bf16 = dz^T            operand of the weight-gradient GEMM dW = dZ^T col  (columns >= M untouched)
//   colsum[c] += sum_m dz[m][c]         bias gradient (fp32 atomics, one per column and block)
// One block works on tiles of 64 rows: thread (r = tid / 4, part = tid % 4) computes 16 columns of row r, the tile is
// transposed through shared memory, thread (c = tid / 4, part) then owns 16 rows of column c.
__global__ void __launch_bounds__(256) elu_backward_fused_kernel(const float* __restrict__ dy,
                                                                  const __nv_bfloat16* __restrict__ y, int pitch,
                                                                  long long M, int C, __nv_bfloat16* __restrict__ dz,
                                                                  __nv_bfloat16* __restrict__ dzt, long long Mp,
                                                                  float* __restrict__ colsum) {
  __shared__ __nv_bfloat16 sh[64][72];  // [column][row], rows padded to 72 (144 B: 16-byte aligned, conflict-light)
  const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
  const long long tiles = (M + 63) >> 6;
  float acc = 0.f;  // running sum of column r (the thread's column in the second half), over this block's tiles
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const long long m = t * 64 + r;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
    if (m < M && part * 16 < C) {
      const float4* dp = reinterpret_cast<const float4*>(dy + m * C + part * 16);
      const uint4* yp = reinterpret_cast<const uint4*>(y + m * pitch + part * 16);
      const uint4 y0 = yp[0], y1 = yp[1];
      const uint32_t yw[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 d = dp[q4];
        const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * q4 + e;
          const uint32_t w = yw[j >> 1];
          const float yy = __uint_as_float((j & 1) ? (w & 0xFFFF0000u) : (w << 16));
          v[j] = dd[e] * (yy > 0.f ? 1.f : yy + 1.f);
        }
      }
    }
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 t2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      pk[j] = *reinterpret_cast<const uint32_t*>(&t2);
    }
    if (m < M) {
      uint4* out = reinterpret_cast<uint4*>(dz + m * 64 + part * 16);
      out[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      out[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    __syncthreads();  // the previous tile's readers are done with `sh`
    unsigned short* shs = reinterpret_cast<unsigned short*>(&sh[0][0]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      shs[(part * 16 + 2 * j) * 72 + r] = (unsigned short)(pk[j] & 0xFFFFu);
      shs[(part * 16 + 2 * j + 1) * 72 + r] = (unsigned short)(pk[j] >> 16);
    }
    __syncthreads();
    // second half: thread (c = r, part) owns rows part*16 .. +16 of column c
    const uint4 t0 = *reinterpret_cast<const uint4*>(&sh[r][part * 16]);
    const uint4 t1 = *reinterpret_cast<const uint4*>(&sh[r][part * 16 + 8]);
    const long long m0 = t * 64 + part * 16;
    if (m0 + 16 <= Mp) {
      uint4* o = reinterpret_cast<uint4*>(dzt + (long long)r * Mp + m0);
      o[0] = t0;
      o[1] = t1;
    }
    const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __uint_as_float(tw[j] << 16) + __uint_as_float(tw[j] & 0xFFFF0000u);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (colsum && part == 0 && r < C) atomicAdd(colsum + r, acc);
}

// colT[(tap*Ci + ci)][m] = A[f][2p-1+r][2q-1+s][ci] (0 outside), m = (f*Ho + p)*Wo + q, tap = r*3+s.
// One thread = one (tap, m): reads Ci contiguous channels, writes Ci rows (coalesced along m across the warp).
template <int CI>
__global__ void __launch_bounds__(256) im2col_t_kernel(const __nv_bfloat16* __restrict__ a, int pitch, int F, int Hi,
                                                        int Wi, int Ho, int Wo, long long Mp,
                                                        __nv_bfloat16* __restrict__ colT) {
  const long long M = (long long)F * Ho * Wo;
  const long long total = 9 * M;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i / M);
    const long long m = i - (long long)tap * M;
    const int q = (int)(m % Wo);
    const int p = (int)((m / Wo) % Ho);
    const long long f = m / ((long long)Wo * Ho);
    const int yy = 2 * p - 1 + tap / 3, xx = 2 * q - 1 + tap % 3;
    __nv_bfloat16 v[CI];
    if (yy >= 0 && yy < Hi && xx >= 0 && xx < Wi) {
      const __nv_bfloat16* src = a + ((f * Hi + yy) * Wi + xx) * pitch;
#pragma unroll
      for (int c = 0; c < CI; ++c) v[c] = src[c];
    } else {
#pragma unroll
      for (int c = 0; c < CI; ++c) v[c] = __float2bfloat16_rn(0.f);
    }
#pragma unroll
    for (int c = 0; c < CI; ++c) colT[(long long)(tap * CI + c) * Mp + m] = v[c];
  }
}

// CI = 32, vectorised: one thread = one tap and FOUR consecutive m: 16-byte channel loads, 8-byte stores per row.
__global__ void __launch_bounds__(256) im2col_t32_v4_kernel(const __nv_bfloat16* __restrict__ a, int pitch, int F, int Hi,
                                                             int Wi, int Ho, int Wo, long long Mp,
                                                             __nv_bfloat16* __restrict__ colT) {
  const long long M = (long long)F * Ho * Wo;
  const long long M4 = (M + 3) >> 2;
  const long long total = 9 * M4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i / M4);
    const long long m0 = (i - (long long)tap * M4) * 4;
    uint32_t v[4][16];  // [pixel][channel pair]
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long m = m0 + k;
      bool ok = m < M;
      const uint4* src = nullptr;
      if (ok) {
        const int q = (int)(m % Wo);
        const int p = (int)((m / Wo) % Ho);
        const long long f = m / ((long long)Wo * Ho);
        const int yy = 2 * p - 1 + tap / 3, xx = 2 * q - 1 + tap % 3;
        ok = yy >= 0 && yy < Hi && xx >= 0 && xx < Wi;
        src = reinterpret_cast<const uint4*>(a + ((f * Hi + yy) * Wi + xx) * pitch);
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 t = ok ? src[g] : make_uint4(0u, 0u, 0u, 0u);
        v[k][4 * g] = t.x; v[k][4 * g + 1] = t.y; v[k][4 * g + 2] = t.z; v[k][4 * g + 3] = t.w;
      }
    }
#pragma unroll
    for (int cp = 0; cp < 16; ++cp) {  // channels 2cp (low halves) and 2cp + 1 (high halves)
      uint2 lo, hi;
      lo.x = (v[0][cp] & 0xFFFFu) | (v[1][cp] << 16);
      lo.y = (v[2][cp] & 0xFFFFu) | (v[3][cp] << 16);
      hi.x = (v[0][cp] >> 16) | (v[1][cp] & 0xFFFF0000u);
      hi.y = (v[2][cp] >> 16) | (v[3][cp] & 0xFFFF0000u);
      *reinterpret_cast<uint2*>(colT + (long long)(tap * 32 + 2 * cp) * Mp + m0) = lo;
      *reinterpret_cast<uint2*>(colT + (long long)(tap * 32 + 2 * cp + 1) * Mp + m0) = hi;
    }
  }
}

// dA[f][y][x][ci] = sum over taps (r,s) with y = 2p-1+r, x = 2q-1+s of dcol[(f,p,q)][(r*3+s)*Ci + ci]   (fp32 out)
template <int CI>
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dcol, int Kp, int F, int Hi,
                                                      int Wi, int Ho, int Wo, float* __restrict__ dA) {
  const long long total = (long long)F * Hi * Wi;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wi);
    const int y = (int)((i / Wi) % Hi);
    const long long f = i / ((long long)Wi * Hi);
    float acc[CI];
#pragma unroll
    for (int c = 0; c < CI; ++c) acc[c] = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int t = y + 1 - r;
      if (t < 0 || (t & 1)) continue;
      const int p = t >> 1;
      if (p >= Ho) continue;
      for (int s = 0; s < 3; ++s) {
        const int u = x + 1 - s;
        if (u < 0 || (u & 1)) continue;
        const int q = u >> 1;
        if (q >= Wo) continue;
        const uint4* src = reinterpret_cast<const uint4*>(dcol + ((f * Ho + p) * Wo + q) * Kp + (r * 3 + s) * CI);
#pragma unroll
        for (int g = 0; g < CI / 8; ++g) {  // 16-byte loads: 8 bf16 each
          const uint4 t = src[g];
          const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[8 * g + 2 * e] += __uint_as_float(w[e] << 16);
            acc[8 * g + 2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
          }
        }
      }
    }
    float4* out = reinterpret_cast<float4*>(dA + i * CI);
#pragma unroll
    for (int g = 0; g < CI / 4; ++g) out[g] = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
  }
}

// dx = gamma * rstd * (dy - sum_dy / count - xhat * sum_dy_xhat / count)
__global__ void __launch_bounds__(256) bn_dx_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                     const float* __restrict__ x, long long ldx, long long M, int D,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ sum_dy_xhat,
                                                     const float* __restrict__ sum_dy, float inv_count,
                                                     float* __restrict__ dx, long long lddx) {
  const long long total = M * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / D;
    const int j = (int)(i - m * D);
    const float xh = (x[m * ldx + j] - mean[j]) * rstd[j];
    const float g = __bfloat162float(dy[m * lddy + j]);
    dx[m * lddx + j] = gamma[j] * rstd[j] * (g - sum_dy[j] * inv_count - xh * sum_dy_xhat[j] * inv_count);
  }
}

// bf16 (M, ld) -> fp32 (M, D) (input gradient without BatchNorm)
__global__ void __launch_bounds__(256) bf16_rows_to_f32_kernel(const __nv_bfloat16* __restrict__ s, long long lds,
                                                                long long M, int D, float* __restrict__ d, long long ldd) {
  const long long total = M * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / D;
    const int j = (int)(i - m * D);
    d[m * ldd + j] = __bfloat162float(s[m * lds + j]);
  }
}

inline unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  return (unsigned)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

}  // namespace

extern "C" int pvr_convfeat_gather(const void* y_bf16, int pitch, int TB, int N, int h, int w, int C, float* feat,
                                   void* stream) {
  if (!y_bf16 || !feat || TB <= 0 || N <= 0 || h <= 0 || w <= 0 || C <= 0 || pitch < C) {
    pvr_set_error("pvr_convfeat_gather: invalid argument");
    return PVR_ERR_ARG;
  }
  feat_gather_kernel<<<grid_for((long long)TB * C * w * h * N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(y_bf16), pitch, TB, N, h, w, C, feat);
  PVR_CHECK_LAUNCH("pvr_convfeat_gather");
  return PVR_OK;
}

extern "C" int pvr_convfeat_scatter(const float* dfeat, int64_t ld, int TB, int N, int h, int w, int C, float* dy,
                                    void* stream) {
  if (!dfeat || !dy || TB <= 0 || N <= 0 || h <= 0 || w <= 0 || C <= 0) {
    pvr_set_error("pvr_convfeat_scatter: invalid argument");
    return PVR_ERR_ARG;
  }
  feat_scatter_kernel<<<grid_for((long long)TB * C * w * h * N), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dfeat, ld, TB, N, h, w, C, dy);
  PVR_CHECK_LAUNCH("pvr_convfeat_scatter");
  return PVR_OK;
}

extern "C" int pvr_elu_backward(const float* dy, const void* y_bf16, int pitch, int64_t M, int C, void* dz_bf16,
                                void* stream) {
  if (!dy || !y_bf16 || !dz_bf16 || M <= 0 || C <= 0 || C > 64 || pitch < C) {
    pvr_set_error("pvr_elu_backward: invalid argument");
    return PVR_ERR_ARG;
  }
  elu_backward_kernel<<<grid_for(M * 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, static_cast<const __nv_bfloat16*>(y_bf16), pitch, M, C, static_cast<__nv_bfloat16*>(dz_bf16));
  PVR_CHECK_LAUNCH("pvr_elu_backward");
  return PVR_OK;
}

extern "C" int pvr_elu_backward_fused(const float* dy, const void* y_bf16, int pitch, int64_t M, int C, void* dz_bf16,
                                      void* dzt_bf16, int64_t Mp, float* colsum, void* stream) {
  if (!dy || !y_bf16 || !dz_bf16 || !dzt_bf16 || M <= 0 || (C != 32 && C != 16 && C != 48 && C != 64) || pitch < C ||
      pitch % 8 || Mp < M || Mp % 16) {
    pvr_set_error("pvr_elu_backward_fused: invalid argument (C multiple of 16, pitch multiple of 8, Mp multiple of 16)");
    return PVR_ERR_ARG;
  }
  const long long tiles = (M + 63) / 64;
  const unsigned grid = (unsigned)(tiles > 148 * 8 ? 148 * 8 : tiles);
  elu_backward_fused_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dy, static_cast<const __nv_bfloat16*>(y_bf16), pitch, M, C, static_cast<__nv_bfloat16*>(dz_bf16),
      static_cast<__nv_bfloat16*>(dzt_bf16), Mp, colsum);
  PVR_CHECK_LAUNCH("pvr_elu_backward_fused");
  return PVR_OK;
}

extern "C" int pvr_im2col_t(const void* a_bf16, int pitch, int F, int Hi, int Wi, int Ci, int Ho, int Wo, int64_t Mp,
                            void* colT_bf16, void* stream) {
  if (!a_bf16 || !colT_bf16 || F <= 0 || (Ci != 4 && Ci != 32) || pitch < Ci || Mp < (int64_t)F * Ho * Wo) {
    pvr_set_error("pvr_im2col_t: invalid argument (Ci must be 4 or 32)");
    return PVR_ERR_ARG;
  }
  const long long total = 9ll * F * Ho * Wo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Ci == 4)
    im2col_t_kernel<4><<<grid_for(total), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a_bf16), pitch, F, Hi, Wi, Ho,
                                                        Wo, Mp, static_cast<__nv_bfloat16*>(colT_bf16));
  else if (pitch % 8 == 0 && Mp % 4 == 0 && (reinterpret_cast<uintptr_t>(a_bf16) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(colT_bf16) & 7) == 0)
    im2col_t32_v4_kernel<<<grid_for((total + 3) / 4), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a_bf16), pitch, F, Hi,
                                                                    Wi, Ho, Wo, Mp, static_cast<__nv_bfloat16*>(colT_bf16));
  else
    im2col_t_kernel<32><<<grid_for(total), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(a_bf16), pitch, F, Hi, Wi, Ho,
                                                         Wo, Mp, static_cast<__nv_bfloat16*>(colT_bf16));
  PVR_CHECK_LAUNCH("pvr_im2col_t");
  return PVR_OK;
}

extern "C" int pvr_col2im(const void* dcol_bf16, int Kp, int F, int Hi, int Wi, int Ci, int Ho, int Wo, float* dA,
                          void* stream) {
  if (!dcol_bf16 || !dA || F <= 0 || Ci != 32 || Kp < 9 * Ci) {
    pvr_set_error("pvr_col2im: invalid argument (Ci must be 32)");
    return PVR_ERR_ARG;
  }
  col2im_kernel<32><<<grid_for((long long)F * Hi * Wi), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dcol_bf16), Kp, F, Hi, Wi, Ho, Wo, dA);
  PVR_CHECK_LAUNCH("pvr_col2im");
  return PVR_OK;
}

extern "C" int pvr_bn1d_backward_dx(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, int64_t M, int D,
                                    const float* mean, const float* rstd, const float* gamma,
                                    const float* sum_dy_xhat, const float* sum_dy, double count, float* dx,
                                    int64_t lddx, void* stream) {
  if (!dy_bf16 || !x || !mean || !rstd || !gamma || !sum_dy_xhat || !sum_dy || !dx || M <= 0 || D <= 0 || count <= 0) {
    pvr_set_error("pvr_bn1d_backward_dx: invalid argument");
    return PVR_ERR_ARG;
  }
  bn_dx_kernel<<<grid_for(M * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), lddy, x, ldx, M, D, mean, rstd, gamma, sum_dy_xhat, sum_dy,
      (float)(1.0 / count), dx, lddx);
  PVR_CHECK_LAUNCH("pvr_bn1d_backward_dx");
  return PVR_OK;
}

extern "C" int pvr_bf16_rows_to_f32(const void* src_bf16, int64_t lds, int64_t M, int D, float* dst, int64_t ldd,
                                    void* stream) {
  if (!src_bf16 || !dst || M <= 0 || D <= 0) {
    pvr_set_error("pvr_bf16_rows_to_f32: invalid argument");
    return PVR_ERR_ARG;
  }
  bf16_rows_to_f32_kernel<<<grid_for(M * D), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src_bf16), lds, M, D, dst, ldd);
  PVR_CHECK_LAUNCH("pvr_bf16_rows_to_f32");
  return PVR_OK;
}
