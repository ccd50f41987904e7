// BC policy kernels around the tcgen05 GEMMs: BatchNorm1d (train mode), LSTM cell forward / backward, policy +
// baseline heads, softmax cross-entropy, bias / weight-gradient reductions, layout helpers and the fused
// clip + RMSprop / Adam optimizer.
//
// Reference semantics: src/models.py:22-89 (PolicyNet: [BatchNorm1d] -> Linear+ReLU x2 -> 2-layer LSTM stepped one
// timestep at a time with done masking -> policy / baseline heads), main_bc_2.py:211-227 (mean NLL of log_softmax,
// grad-norm statistic, clip_grad_norm_(40), RMSprop(alpha .99, eps 1e-5; eps added outside the sqrt) with LambdaLR).
#include "pvr_b200.h"

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern void pvr_set_error(const char* fmt, ...);

namespace {

#define PVR_LAUNCH_CHECK(name)                                       \
  do {                                                               \
    cudaError_t e_ = cudaGetLastError();                             \
    if (e_ != cudaSuccess) {                                         \
      pvr_set_error("%s: %s", name, cudaGetErrorString(e_));         \
      return PVR_ERR_CUDA;                                           \
    }                                                                \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------------------------------------- BatchNorm1d
// sums[0..d) += sum_m x[m][j], sums[d..2d) += sum_m x[m][j]^2 (double): block = 32 columns x 8 row lanes.
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, long long ldx, int m, int d,
                                                        int rows_per_block, double* __restrict__ sums) {
  __shared__ double sh[2][8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(m, r0 + rows_per_block);
  double s = 0.0, ss = 0.0;
  if (col < d) {
    for (int r = r0 + rl; r < r1; r += 8) {
      const double v = (double)x[(long long)r * ldx + col];
      s += v;
      ss += v * v;
    }
  }
  sh[0][rl][threadIdx.x & 31] = s;
  sh[1][rl][threadIdx.x & 31] = ss;
  __syncthreads();
  if (rl == 0 && col < d) {
    for (int k = 1; k < 8; ++k) {
      s += sh[0][k][threadIdx.x & 31];
      ss += sh[1][k][threadIdx.x & 31];
    }
    atomicAdd(&sums[col], s);
    atomicAdd(&sums[d + col], ss);
  }
}

// mean / rstd from the (all-reduced) sums; running stats with momentum (unbiased variance), like nn.BatchNorm1d.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int d, double count, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  const double mu = sums[j] / count;
  double var = sums[d + j] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[j] = (float)mu;
  rstd[j] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * (float)mu;
    running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
  }
}

__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, int d,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= d) return;
  mean[j] = rm[j];
  rstd[j] = 1.f / sqrtf(rv[j] + eps);
}

// y = (x - mean) * rstd * gamma + beta -> bf16 ; 4 columns per thread.
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, long long ldx, int m, int d,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ y, long long ldy) {
  const int d4 = d >> 2;
  const long long total = (long long)m * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d4), c = (int)(i - (long long)r * d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * ldx + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    __nv_bfloat162 o0 = __floats2bfloat162_rn((v.x - mu.x) * rs.x * g.x + b.x, (v.y - mu.y) * rs.y * g.y + b.y);
    __nv_bfloat162 o1 = __floats2bfloat162_rn((v.z - mu.z) * rs.z * g.z + b.z, (v.w - mu.w) * rs.w * g.w + b.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&o0);
    o.y = *reinterpret_cast<uint32_t*>(&o1);
    *reinterpret_cast<uint2*>(y + (long long)r * ldy + c) = o;
  }
}

// scalar variants for feature counts / pitches that are not multiples of 4
__global__ void __launch_bounds__(256) bn_apply_scalar_kernel(const float* __restrict__ x, long long ldx, int m, int d,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta,
                                                               __nv_bfloat16* __restrict__ y, long long ldy) {
  const long long total = (long long)m * d;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d), c = (int)(i - (long long)r * d);
    const float v = x[(long long)r * ldx + c];
    const float o = gamma ? (v - mean[c]) * rstd[c] * gamma[c] + beta[c] : v;
    y[(long long)r * ldy + c] = __float2bfloat16_rn(o);
  }
}

// float -> bf16 cast of a (m x d) matrix (policy input without BatchNorm).
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ x, long long ldx, int m, int d,
                                                         __nv_bfloat16* __restrict__ y, long long ldy) {
  const int d4 = d >> 2;
  const long long total = (long long)m * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / d4), c = (int)(i - (long long)r * d4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (long long)r * ldx + c);
    __nv_bfloat162 o0 = __floats2bfloat162_rn(v.x, v.y), o1 = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&o0);
    o.y = *reinterpret_cast<uint32_t*>(&o1);
    *reinterpret_cast<uint2*>(y + (long long)r * ldy + c) = o;
  }
}

// dgamma[j] += sum_m dy[m][j] * xhat[m][j], dbeta[j] += sum_m dy[m][j]  (dy bf16 = gradient w.r.t. the BN output)
__global__ void __launch_bounds__(256) bn_backward_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                           const float* __restrict__ x, long long ldx, int m, int d,
                                                           int rows_per_block, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
  __shared__ float sh[2][8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(m, r0 + rows_per_block);
  float sg = 0.f, sb = 0.f;
  if (col < d) {
    const float mu = mean[col], rs = rstd[col];
    for (int r = r0 + rl; r < r1; r += 8) {
      const float g = __bfloat162float(dy[(long long)r * lddy + col]);
      sg += g * (x[(long long)r * ldx + col] - mu) * rs;
      sb += g;
    }
  }
  sh[0][rl][threadIdx.x & 31] = sg;
  sh[1][rl][threadIdx.x & 31] = sb;
  __syncthreads();
  if (rl == 0 && col < d) {
    for (int k = 1; k < 8; ++k) {
      sg += sh[0][k][threadIdx.x & 31];
      sb += sh[1][k][threadIdx.x & 31];
    }
    atomicAdd(&dgamma[col], sg);
    atomicAdd(&dbeta[col], sb);
  }
}

// ---------------------------------------------------------------------------------------------- LSTM cell
// One thread = one (sequence b, hidden unit j). Gate order i, f, g, o (PyTorch).
// pre = G[b][gate*H + j] + XP[b][gate*H + j]; c = sig(f) * (nd * c_prev) + sig(i) * tanh(g); h = sig(o) * tanh(c).
// Programmatic dependent launch inside the recurrence (GEMM -> cell -> GEMM -> ...): every kernel lets its successor
// start its prologue at once and waits for its predecessor's results before touching memory (see ptx.cuh).
__device__ __forceinline__ void pdl_launch_then_wait() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(
    const float* __restrict__ G, const float* __restrict__ XP, const float* __restrict__ c_prev,
    const float* __restrict__ nd, const float* __restrict__ nd_next, int B, int H, float* __restrict__ gates,
    float* __restrict__ c_out, float* __restrict__ h_out_f32, __nv_bfloat16* __restrict__ h_out,
    __nv_bfloat16* __restrict__ hm_next) {
  pdl_launch_then_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const long long g0 = (long long)b * 4 * H + j;
  float pi = G[g0], pf = G[g0 + H], pg = G[g0 + 2 * H], po = G[g0 + 3 * H];
  if (XP) {  // (null when the recurrent GEMM accumulated straight into the input projection)
    pi += XP[g0]; pf += XP[g0 + H]; pg += XP[g0 + 2 * H]; po += XP[g0 + 3 * H];
  }
  const float i = sigmoidf_(pi), f = sigmoidf_(pf), g = tanhf(pg), o = sigmoidf_(po);
  const float c = f * (nd[b] * c_prev[idx]) + i * g;
  const float h = o * tanhf(c);
  gates[g0] = i;
  gates[g0 + H] = f;
  gates[g0 + 2 * H] = g;
  gates[g0 + 3 * H] = o;
  c_out[idx] = c;
  h_out_f32[idx] = h;
  h_out[idx] = __float2bfloat16_rn(h);
  if (hm_next) hm_next[idx] = __float2bfloat16_rn(h * nd_next[b]);  // state entering step t+1 is masked by done[t+1]
}

// Backward of one step. dh = dh_out[b][j] (from above) + nd_next[b] * dh_rec[b][j] (from step t+1 through W_hh);
// dh_rec is zeroed after the read so the next split-K GEMM can accumulate into it.
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(
    const float* __restrict__ dh_out, float* __restrict__ dh_rec, float* __restrict__ dc_rec,
    const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_cur,
    const float* __restrict__ nd, const float* __restrict__ nd_next, int B, int H,
    __nv_bfloat16* __restrict__ dG) {
  pdl_launch_then_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, j = idx - b * H;
  const long long g0 = (long long)b * 4 * H + j;
  const float i = gates[g0], f = gates[g0 + H], g = gates[g0 + 2 * H], o = gates[g0 + 3 * H];
  float dh = dh_out ? dh_out[idx] : 0.f;
  if (nd_next) dh += nd_next[b] * dh_rec[idx];
  else dh += dh_rec[idx];  // last step: gradient of the returned state (zero in BC training)
  dh_rec[idx] = 0.f;
  const float tc = tanhf(c_cur[idx]);
  const float dc = dc_rec[idx] + dh * o * (1.f - tc * tc);
  const float cpm = nd[b] * c_prev[idx];
  dG[g0] = __float2bfloat16_rn(dc * g * i * (1.f - i));
  dG[g0 + H] = __float2bfloat16_rn(dc * cpm * f * (1.f - f));
  dG[g0 + 2 * H] = __float2bfloat16_rn(dc * i * (1.f - g * g));
  dG[g0 + 3 * H] = __float2bfloat16_rn(dh * tc * o * (1.f - o));
  dc_rec[idx] = dc * f * nd[b];
}

// hm0[b][j] = nd0[b] * h0[b][j] -> bf16 (state entering step 0, src/models.py:69-70)
__global__ void __launch_bounds__(256) mask_state_kernel(const float* __restrict__ h0, const float* __restrict__ nd0,
                                                          int B, int H, __nv_bfloat16* __restrict__ hm0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * H) hm0[idx] = __float2bfloat16_rn(h0[idx] * nd0[idx / H]);
}

// ---------------------------------------------------------------------------------------------- heads + loss
// One warp per row: logits[m][a] = h[m] . Wp[a] + bp[a] (a < A), baseline[m] = h[m] . Wb + bb.  K = hidden size.
template <int MAXA>
__global__ void __launch_bounds__(256) heads_fwd_kernel(const __nv_bfloat16* __restrict__ h, int m, int K,
                                                         const float* __restrict__ Wp, const float* __restrict__ bp,
                                                         const float* __restrict__ Wb, const float* __restrict__ bb,
                                                         int A, float* __restrict__ logits,
                                                         float* __restrict__ baseline) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m) return;
  float acc[MAXA + 1];
#pragma unroll
  for (int a = 0; a <= MAXA; ++a) acc[a] = 0.f;
  for (int k = lane * 2; k < K; k += 64) {
    const __nv_bfloat162 hv = *reinterpret_cast<const __nv_bfloat162*>(h + (long long)row * K + k);
    const float h0 = __low2float(hv), h1 = __high2float(hv);
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
      if (a < A) acc[a] += h0 * Wp[a * K + k] + h1 * Wp[a * K + k + 1];
    acc[MAXA] += h0 * Wb[k] + h1 * Wb[k + 1];
  }
#pragma unroll
  for (int a = 0; a <= MAXA; ++a) acc[a] = warp_sum(acc[a]);
  if (lane == 0) {
    for (int a = 0; a < A; ++a) logits[(long long)row * A + a] = acc[a] + bp[a];
    baseline[row] = acc[MAXA] + bb[0];
  }
}

// Same product with the (A + 1) weight rows staged in shared memory and 16-byte loads of h: a warp walks rows with a
// grid stride, a lane owns 8 consecutive k per 256. K % 8 == 0.
template <int MAXA>
__global__ void __launch_bounds__(256) heads_fwd_staged_kernel(const __nv_bfloat16* __restrict__ h, int m, int K,
                                                                const float* __restrict__ Wp,
                                                                const float* __restrict__ bp,
                                                                const float* __restrict__ Wb,
                                                                const float* __restrict__ bb, int A,
                                                                float* __restrict__ logits,
                                                                float* __restrict__ baseline) {
  extern __shared__ float sw[];  // [(A + 1)][K]: policy rows, then the baseline row
  for (int i = threadIdx.x; i < (A + 1) * K; i += 256) sw[i] = i < A * K ? Wp[i] : Wb[i - A * K];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < m; row += gridDim.x * 8) {
    float acc[MAXA + 1];
#pragma unroll
    for (int a = 0; a <= MAXA; ++a) acc[a] = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 raw = *reinterpret_cast<const uint4*>(h + (long long)row * K + k);
      float hv[8];
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hv[2 * j] = __uint_as_float(w4[j] << 16);
        hv[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u);
      }
#pragma unroll
      for (int a = 0; a <= MAXA; ++a) {
        const int ra = a < MAXA ? a : A;  // slot MAXA = the baseline row
        if (a < MAXA && a >= A) continue;
        const float4 w0 = *reinterpret_cast<const float4*>(sw + ra * K + k);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + ra * K + k + 4);
        acc[a] += hv[0] * w0.x + hv[1] * w0.y + hv[2] * w0.z + hv[3] * w0.w + hv[4] * w1.x + hv[5] * w1.y +
                  hv[6] * w1.z + hv[7] * w1.w;
      }
    }
#pragma unroll
    for (int a = 0; a <= MAXA; ++a) acc[a] = warp_sum(acc[a]);
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < A) logits[(long long)row * A + a] = acc[a] + bp[a];
      baseline[row] = acc[MAXA] + bb[0];
    }
  }
}

// Mean softmax cross-entropy over rows (main_bc_2.py:211-214) and its gradient. One thread per row for the softmax
// over A actions, warp-shuffle + one atomic per warp for the loss sum.
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ tgt,
                                                       int m, int A, float inv_count, float* __restrict__ loss,
                                                       float* __restrict__ dlogits) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float nll = 0.f;
  if (row < m) {
    const float* l = logits + (long long)row * A;
    float mx = l[0];
    for (int a = 1; a < A; ++a) mx = fmaxf(mx, l[a]);
    float se = 0.f;
    for (int a = 0; a < A; ++a) se += expf(l[a] - mx);
    const float lse = mx + logf(se);
    const int t = (int)tgt[row];
    nll = lse - l[t];
    for (int a = 0; a < A; ++a)
      dlogits[(long long)row * A + a] = (expf(l[a] - lse) - (a == t ? 1.f : 0.f)) * inv_count;
  }
  nll = warp_sum(nll);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, nll * inv_count);
}

// dh[m][k] = scale * sum_a dl[m][a] * Wp[a][k]   (fp32, one thread per (row, 4 k))
__global__ void __launch_bounds__(256) heads_bwd_dh_kernel(const float* __restrict__ dl, const float* __restrict__ Wp,
                                                            int m, int K, int A, float scale, float* __restrict__ dh) {
  const int k4 = K >> 2;
  const long long total = (long long)m * k4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / k4), k = (int)(i - (long long)r * k4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < A; ++a) {
      const float g = dl[(long long)r * A + a] * scale;
      const float4 w = *reinterpret_cast<const float4*>(Wp + (long long)a * K + k);
      acc.x += g * w.x; acc.y += g * w.y; acc.z += g * w.z; acc.w += g * w.w;
    }
    *reinterpret_cast<float4*>(dh + (long long)r * K + k) = acc;
  }
}

// dWp[a][k] += scale * sum_m dl[m][a] * h[m][k]; dbp[a] += scale * sum_m dl[m][a].  grid (K/32, row chunks).
template <int MAXA>
__global__ void __launch_bounds__(256) heads_bwd_dw_kernel(const float* __restrict__ dl,
                                                            const __nv_bfloat16* __restrict__ h, int m, int K, int A,
                                                            int rows_per_block, float scale, float* __restrict__ dWp,
                                                            float* __restrict__ dbp) {
  __shared__ float sh[8][MAXA + 1][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
  float acc[MAXA + 1];
#pragma unroll
  for (int a = 0; a <= MAXA; ++a) acc[a] = 0.f;
  for (int r = r0 + rl; r < r1; r += 8) {
    const float hv = k < K ? __bfloat162float(h[(long long)r * K + k]) : 0.f;
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
      if (a < A) {
        const float g = dl[(long long)r * A + a];
        acc[a] += g * hv;
        if (blockIdx.x == 0 && lane == a) acc[MAXA] += g;  // bias gradient, once
      }
  }
#pragma unroll
  for (int a = 0; a <= MAXA; ++a) sh[rl][a][lane] = acc[a];
  __syncthreads();
  if (rl == 0) {
#pragma unroll
    for (int a = 0; a <= MAXA; ++a)
      for (int q = 1; q < 8; ++q) acc[a] += sh[q][a][lane];
    if (k < K)
      for (int a = 0; a < A; ++a) atomicAdd(&dWp[(long long)a * K + k], acc[a] * scale);
    if (blockIdx.x == 0 && lane < A) atomicAdd(&dbp[lane], acc[MAXA] * scale);
  }
}

// Same sums with 16-byte loads of h: a thread owns 8 consecutive k, a block 128 such threads (1024 k) x 2 row lanes.
// grid (ceil(K / 1024), row chunks). K % 8 == 0.
template <int MAXA>
__global__ void __launch_bounds__(256) heads_bwd_dw8_kernel(const float* __restrict__ dl,
                                                             const __nv_bfloat16* __restrict__ h, int m, int K, int A,
                                                             int rows_per_block, float scale, float* __restrict__ dWp,
                                                             float* __restrict__ dbp) {
  __shared__ float sh[128][8 * MAXA + 1];
  const int kt = threadIdx.x & 127, rl = threadIdx.x >> 7;
  const int k = blockIdx.x * 1024 + kt * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
  float acc[MAXA][8];
  float bacc = 0.f;
#pragma unroll
  for (int a = 0; a < MAXA; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j] = 0.f;
  if (k < K) {
#pragma unroll 4
    for (int r = r0 + rl; r < r1; r += 2) {
      const uint4 raw = *reinterpret_cast<const uint4*>(h + (long long)r * K + k);
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
      float hv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hv[2 * j] = __uint_as_float(w4[j] << 16);
        hv[2 * j + 1] = __uint_as_float(w4[j] & 0xffff0000u);
      }
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < A) {
          const float g = __ldg(dl + (long long)r * A + a);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[a][j] += g * hv[j];
          if (blockIdx.x == 0 && kt == a) bacc += g;  // bias gradient, once
        }
    }
  }
  if (rl == 1) {
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
#pragma unroll
      for (int j = 0; j < 8; ++j) sh[kt][a * 8 + j] = acc[a][j];
    sh[kt][8 * MAXA] = bacc;
  }
  __syncthreads();
  if (rl == 0 && k < K) {
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
      if (a < A) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd(&dWp[(long long)a * K + k + j], (acc[a][j] + sh[kt][a * 8 + j]) * scale);
      }
    if (blockIdx.x == 0 && kt < A) atomicAdd(&dbp[kt], (bacc + sh[kt][8 * MAXA]) * scale);
  }
}

// ---------------------------------------------------------------------------------------------- reductions / layout
// out[n] += sum_m y[m][n]  (bias gradients), y bf16.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ y, long long ldy, int m,
                                                           int n, int rows_per_block, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
  float s = 0.f;
  if (col < n)
    for (int r = r0 + rl; r < r1; r += 8) s += __bfloat162float(y[(long long)r * ldy + col]);
  sh[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && col < n) {
    for (int q = 1; q < 8; ++q) s += sh[q][lane];
    atomicAdd(&out[col], s);
  }
}

// out (cols x rows) = in (rows x cols)^T, bf16, 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                              int rows, int cols, __nv_bfloat16* __restrict__ out,
                                                              long long ldo) {
  __shared__ __nv_bfloat16 t[32][34];
  const int c0 = blockIdx.y * 32, r0 = blockIdx.x * 32;  // rows (millions in the finetune backward) on grid.x
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows && c0 + tx < cols) t[i][tx] = in[(long long)(r0 + i) * ldi + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) out[(long long)(c0 + i) * ldo + r0 + tx] = t[tx][i];
}

// fp32 (rows x cols) -> bf16 copy and (optionally) bf16 transposed copy (cols x rows): weights after each update.
__global__ void __launch_bounds__(256) cast_weight_kernel(const float* __restrict__ w, int rows, int cols,
                                                           __nv_bfloat16* __restrict__ wb, long long ldb,
                                                           __nv_bfloat16* __restrict__ wt, long long ldt) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows && c0 + tx < cols) {
      const float v = w[(long long)(r0 + i) * cols + c0 + tx];
      t[i][tx] = v;
      if (wb) wb[(long long)(r0 + i) * ldb + c0 + tx] = __float2bfloat16_rn(v);
    }
  if (!wt) return;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) wt[(long long)(c0 + i) * ldt + r0 + tx] = __float2bfloat16_rn(t[tx][i]);
}

// Same copies on 64 x 64 tiles with 16-byte loads and 8-byte stores (rows, cols, ldb, ldt multiples of 4).
__global__ void __launch_bounds__(256) cast_weight64_kernel(const float* __restrict__ w, int rows, int cols,
                                                             __nv_bfloat16* __restrict__ wb, long long ldb,
                                                             __nv_bfloat16* __restrict__ wt, long long ldt) {
  __shared__ float t[64][65];
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int q = (threadIdx.x & 15) * 4, l = threadIdx.x >> 4;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int r = r0 + it * 16 + l, c = c0 + q;
    if (r < rows && c < cols) {
      const float4 v = *reinterpret_cast<const float4*>(w + (long long)r * cols + c);
      t[it * 16 + l][q] = v.x; t[it * 16 + l][q + 1] = v.y; t[it * 16 + l][q + 2] = v.z; t[it * 16 + l][q + 3] = v.w;
      if (wb) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&a);
        o.y = *reinterpret_cast<const uint32_t*>(&b);
        *reinterpret_cast<uint2*>(wb + (long long)r * ldb + c) = o;
      }
    }
  }
  if (!wt) return;
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int cc = it * 16 + l;  // column of w = row of the transposed copy
    if (c0 + cc < cols && r0 + q < rows) {
      const __nv_bfloat162 a = __floats2bfloat162_rn(t[q][cc], t[q + 1][cc]);
      const __nv_bfloat162 b = __floats2bfloat162_rn(t[q + 2][cc], t[q + 3][cc]);
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&b);
      *reinterpret_cast<uint2*>(wt + (long long)(c0 + cc) * ldt + r0 + q) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------- optimizer
struct TensorList {
  float* p[32];
  float* g[32];
  float* s1[32];  // RMSprop square_avg / Adam exp_avg
  float* s2[32];  // Adam exp_avg_sq
  long long n[32];
  int count;
};

// total[0] += sum over all tensors of g^2 (the global grad norm^2 of main_bc_2.py:220-224 / clip_grad_norm_)
__global__ void __launch_bounds__(256) sumsq_kernel(const TensorList tl, double* __restrict__ total) {
  const int t = blockIdx.y;
  const float* g = tl.g[t];
  const long long n = tl.n[t];
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if ((long long)blockIdx.x * blockDim.x * (vec ? 4 : 1) >= n) return;  // small tensors: no idle blocks / atomics
  double s = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    // 16-byte loads, squares and sums in double as in the scalar loop
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll 4
    for (long long i = i0; i < (n >> 2); i += stride) {
      const float4 v = g4[i];
      s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    for (long long i = (n & ~3ll) + i0; i < n; i += stride) s += (double)g[i] * g[i];
  } else {
    for (long long i = i0; i < n; i += stride) s += (double)g[i] * g[i];
  }
  s = warp_sum_d(s);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; ++q) s += sh[q];
    atomicAdd(total, s);
  }
}

// clip coefficient = min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_), then
// mode 0 RMSprop (torch.optim.RMSprop, momentum 0, not centered): v = a v + (1-a) g^2; p -= lr g / (sqrt(v) + eps)
// mode 1 Adam (torch.optim.Adam, no weight decay/amsgrad): bias-corrected, eps added outside the sqrt as well.
__global__ void __launch_bounds__(256) optim_step_kernel(const TensorList tl, const double* __restrict__ sumsq,
                                                          float grad_scale, float max_norm, int mode, float lr,
                                                          float alpha_or_beta1, float beta2, float eps, float bc1,
                                                          float bc2, float* __restrict__ norm_out,
                                                          const float* __restrict__ lr_dev) {
  const int t = blockIdx.y;
  if (lr_dev) lr = __ldg(lr_dev);  // learning rate in device memory: the step can live inside a CUDA graph
  const double norm = sqrt(sumsq[0]) * (double)grad_scale;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / ((float)norm + 1e-6f));
  coef *= grad_scale;
  if (norm_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) norm_out[0] = (float)norm;
  float* p = tl.p[t];
  float* g = tl.g[t];
  float* s1 = tl.s1[t];
  float* s2 = tl.s2[t];
  const long long n = tl.n[t];
  // float4 path for the large, 16-byte aligned tensors (weights): 4x fewer instructions, more bytes in flight
  if (mode == 0 && (n & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1)) & 15) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* v4 = reinterpret_cast<float4*>(s1);
    const float oma = 1.f - alpha_or_beta1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2);
         i += (long long)gridDim.x * blockDim.x) {
      float4 gr = g4[i], v = v4[i], w = p4[i];
      gr.x *= coef; gr.y *= coef; gr.z *= coef; gr.w *= coef;
      v.x = alpha_or_beta1 * v.x + oma * gr.x * gr.x;
      v.y = alpha_or_beta1 * v.y + oma * gr.y * gr.y;
      v.z = alpha_or_beta1 * v.z + oma * gr.z * gr.z;
      v.w = alpha_or_beta1 * v.w + oma * gr.w * gr.w;
      w.x -= lr * gr.x / (sqrtf(v.x) + eps);
      w.y -= lr * gr.y / (sqrtf(v.y) + eps);
      w.z -= lr * gr.z / (sqrtf(v.z) + eps);
      w.w -= lr * gr.w / (sqrtf(v.w) + eps);
      g4[i] = gr;
      v4[i] = v;
      p4[i] = w;
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gr = g[i] * coef;
    g[i] = gr;  // the clipped gradient stays visible in .grad, like clip_grad_norm_
    if (mode == 0) {
      const float v = alpha_or_beta1 * s1[i] + (1.f - alpha_or_beta1) * gr * gr;
      s1[i] = v;
      p[i] -= lr * gr / (sqrtf(v) + eps);
    } else {
      const float m1 = alpha_or_beta1 * s1[i] + (1.f - alpha_or_beta1) * gr;
      const float v = beta2 * s2[i] + (1.f - beta2) * gr * gr;
      s1[i] = m1;
      s2[i] = v;
      p[i] -= (lr / bc1) * m1 / (sqrtf(v) / sqrtf(bc2) + eps);
    }
  }
}

int rows_per_block_for(int m) {
  int rpb = (m + 63) / 64;  // <= 64 row chunks
  return rpb < 8 ? 8 : rpb;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int pvr_bn1d_stats(const float* x, int64_t ldx, int m, int d, double* sums, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (!x || m <= 0 || d <= 0 || !sums) {
    pvr_set_error("pvr_bn1d_stats: invalid argument");
    return PVR_ERR_ARG;
  }
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d, st);
  const int rpb = rows_per_block_for(m);
  dim3 grid((d + 31) / 32, (m + rpb - 1) / rpb);
  bn_stats_kernel<<<grid, 256, 0, st>>>(x, ldx, m, d, rpb, sums);
  PVR_LAUNCH_CHECK("pvr_bn1d_stats");
  return PVR_OK;
}

extern "C" int pvr_bn1d_normalize(const float* x, int64_t ldx, int m, int d, const double* sums, double count,
                                  float eps, float momentum, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float* mean, float* rstd, void* y_bf16,
                                  int64_t ldy, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (!x || m <= 0 || d <= 0 || !sums || !gamma || !beta || !mean || !rstd || !y_bf16 || count <= 0) {
    pvr_set_error("pvr_bn1d_normalize: invalid argument");
    return PVR_ERR_ARG;
  }
  bn_finalize_kernel<<<(d + 255) / 256, 256, 0, st>>>(sums, d, count, eps, momentum, running_mean, running_var, mean,
                                                      rstd);
  PVR_LAUNCH_CHECK("pvr_bn1d_normalize(finalize)");
  if (d % 4 || ldx % 4 || ldy % 4)
    bn_apply_scalar_kernel<<<148 * 8, 256, 0, st>>>(x, ldx, m, d, mean, rstd, gamma, beta,
                                                    static_cast<__nv_bfloat16*>(y_bf16), ldy);
  else
    bn_apply_kernel<<<148 * 8, 256, 0, st>>>(x, ldx, m, d, mean, rstd, gamma, beta,
                                             static_cast<__nv_bfloat16*>(y_bf16), ldy);
  PVR_LAUNCH_CHECK("pvr_bn1d_normalize(apply)");
  return PVR_OK;
}

extern "C" int pvr_bn1d_eval(const float* x, int64_t ldx, int m, int d, const float* running_mean,
                             const float* running_var, float eps, const float* gamma, const float* beta, float* mean,
                             float* rstd, void* y_bf16, int64_t ldy, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (!x || m <= 0 || d <= 0 || !running_mean || !running_var || !gamma || !beta || !mean || !rstd || !y_bf16) {
    pvr_set_error("pvr_bn1d_eval: invalid argument");
    return PVR_ERR_ARG;
  }
  bn_eval_stats_kernel<<<(d + 255) / 256, 256, 0, st>>>(running_mean, running_var, eps, d, mean, rstd);
  PVR_LAUNCH_CHECK("pvr_bn1d_eval(stats)");
  if (d % 4 || ldx % 4 || ldy % 4)
    bn_apply_scalar_kernel<<<148 * 8, 256, 0, st>>>(x, ldx, m, d, mean, rstd, gamma, beta,
                                                    static_cast<__nv_bfloat16*>(y_bf16), ldy);
  else
    bn_apply_kernel<<<148 * 8, 256, 0, st>>>(x, ldx, m, d, mean, rstd, gamma, beta,
                                             static_cast<__nv_bfloat16*>(y_bf16), ldy);
  PVR_LAUNCH_CHECK("pvr_bn1d_eval(apply)");
  return PVR_OK;
}

extern "C" int pvr_cast_rows_bf16(const float* x, int64_t ldx, int m, int d, void* y_bf16, int64_t ldy,
                                  void* stream_) {
  if (!x || !y_bf16 || m <= 0 || d <= 0) {
    pvr_set_error("pvr_cast_rows_bf16: invalid argument");
    return PVR_ERR_ARG;
  }
  if (d % 4 || ldx % 4 || ldy % 4)
    bn_apply_scalar_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, ldx, m, d, nullptr, nullptr, nullptr, nullptr, static_cast<__nv_bfloat16*>(y_bf16), ldy);
  else
    cast_rows_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        x, ldx, m, d, static_cast<__nv_bfloat16*>(y_bf16), ldy);
  PVR_LAUNCH_CHECK("pvr_cast_rows_bf16");
  return PVR_OK;
}

extern "C" int pvr_bn1d_backward(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, int m, int d,
                                 const float* mean, const float* rstd, float* dgamma, float* dbeta, void* stream_) {
  if (!dy_bf16 || !x || !mean || !rstd || !dgamma || !dbeta || m <= 0 || d <= 0) {
    pvr_set_error("pvr_bn1d_backward: invalid argument");
    return PVR_ERR_ARG;
  }
  const int rpb = rows_per_block_for(m);
  dim3 grid((d + 31) / 32, (m + rpb - 1) / rpb);
  bn_backward_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), lddy, x, ldx, m, d, rpb, mean, rstd, dgamma, dbeta);
  PVR_LAUNCH_CHECK("pvr_bn1d_backward");
  return PVR_OK;
}

extern "C" int pvr_lstm_cell_forward(const float* G, const float* XP, const float* c_prev, const float* nd,
                                     const float* nd_next, int B, int H, float* gates, float* c_out, float* h_out_f32,
                                     void* h_out_bf16, void* hm_next_bf16, void* stream_) {
  if (!G || !XP || !c_prev || !nd || !gates || !c_out || !h_out_f32 || !h_out_bf16 || B <= 0 || H <= 0) {
    pvr_set_error("pvr_lstm_cell_forward: invalid argument");
    return PVR_ERR_ARG;
  }
  lstm_cell_fwd_kernel<<<(B * H + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      G, XP, c_prev, nd, nd_next, B, H, gates, c_out, h_out_f32, static_cast<__nv_bfloat16*>(h_out_bf16),
      static_cast<__nv_bfloat16*>(hm_next_bf16));
  PVR_LAUNCH_CHECK("pvr_lstm_cell_forward");
  return PVR_OK;
}

extern "C" int pvr_lstm_cell_backward(const float* dh_out, float* dh_rec, float* dc_rec, const float* gates,
                                      const float* c_prev, const float* c_cur, const float* nd, const float* nd_next,
                                      int B, int H, void* dG_bf16, void* stream_) {
  if (!dh_rec || !dc_rec || !gates || !c_prev || !c_cur || !nd || !dG_bf16 || B <= 0 || H <= 0) {
    pvr_set_error("pvr_lstm_cell_backward: invalid argument");
    return PVR_ERR_ARG;
  }
  lstm_cell_bwd_kernel<<<(B * H + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      dh_out, dh_rec, dc_rec, gates, c_prev, c_cur, nd, nd_next, B, H, static_cast<__nv_bfloat16*>(dG_bf16));
  PVR_LAUNCH_CHECK("pvr_lstm_cell_backward");
  return PVR_OK;
}

extern "C" int pvr_heads_forward(const void* h_bf16, int m, int K, const float* Wp, const float* bp, const float* Wb,
                                 const float* bb, int A, float* logits, float* baseline, void* stream_) {
  if (!h_bf16 || !Wp || !bp || !Wb || !bb || !logits || !baseline || m <= 0 || K <= 0 || K % 64 || A <= 0 || A > 8) {
    pvr_set_error("pvr_heads_forward: invalid argument (A <= 8, K %% 64 == 0)");
    return PVR_ERR_ARG;
  }
  const size_t stage_bytes = (size_t)(A + 1) * K * sizeof(float);
  if (stage_bytes <= 48 * 1024 && (reinterpret_cast<uintptr_t>(h_bf16) & 15) == 0) {
    const int blocks = (m + 7) / 8 < 296 ? (m + 7) / 8 : 296;
    heads_fwd_staged_kernel<8><<<blocks, 256, stage_bytes, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const __nv_bfloat16*>(h_bf16), m, K, Wp, bp, Wb, bb, A, logits, baseline);
  } else {
    heads_fwd_kernel<8><<<(m + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const __nv_bfloat16*>(h_bf16), m, K, Wp, bp, Wb, bb, A, logits, baseline);
  }
  PVR_LAUNCH_CHECK("pvr_heads_forward");
  return PVR_OK;
}

extern "C" int pvr_ce_loss(const float* logits, const int64_t* targets, int m, int A, float inv_count, float* loss,
                           float* dlogits, void* stream_) {
  if (!logits || !targets || !loss || !dlogits || m <= 0 || A <= 0) {
    pvr_set_error("pvr_ce_loss: invalid argument");
    return PVR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  cudaMemsetAsync(loss, 0, sizeof(float), st);
  ce_loss_kernel<<<(m + 255) / 256, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(targets), m, A,
                                                  inv_count, loss, dlogits);
  PVR_LAUNCH_CHECK("pvr_ce_loss");
  return PVR_OK;
}

extern "C" int pvr_heads_backward(const float* dlogits, const void* h_bf16, const float* Wp, int m, int K, int A,
                                  float scale, float* dh, float* dWp, float* dbp, void* stream_) {
  if (!dlogits || !h_bf16 || !Wp || !dh || !dWp || !dbp || m <= 0 || K <= 0 || K % 4 || A <= 0 || A > 8) {
    pvr_set_error("pvr_heads_backward: invalid argument");
    return PVR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  heads_bwd_dh_kernel<<<148 * 8, 256, 0, st>>>(dlogits, Wp, m, K, A, scale, dh);
  PVR_LAUNCH_CHECK("pvr_heads_backward(dh)");
  if (K % 8 == 0 && (reinterpret_cast<uintptr_t>(h_bf16) & 15) == 0) {
    const int kblocks = (K + 1023) / 1024;
    // few row chunks: every block ends with 8 A atomics per thread, which cost more than its share of the rows
    int chunks = 64 / kblocks > 0 ? 64 / kblocks : 1;
    if (chunks > (m + 15) / 16) chunks = (m + 15) / 16;
    const int rpb = (m + chunks - 1) / chunks;
    dim3 grid(kblocks, (m + rpb - 1) / rpb);
    heads_bwd_dw8_kernel<8><<<grid, 256, 0, st>>>(dlogits, static_cast<const __nv_bfloat16*>(h_bf16), m, K, A, rpb,
                                                  scale, dWp, dbp);
  } else {
    const int rpb = rows_per_block_for(m);
    dim3 grid((K + 31) / 32, (m + rpb - 1) / rpb);
    heads_bwd_dw_kernel<8><<<grid, 256, 0, st>>>(dlogits, static_cast<const __nv_bfloat16*>(h_bf16), m, K, A, rpb,
                                                 scale, dWp, dbp);
  }
  PVR_LAUNCH_CHECK("pvr_heads_backward(dW)");
  return PVR_OK;
}

extern "C" int pvr_colsum_bf16(const void* y_bf16, int64_t ldy, int m, int n, float* out, void* stream_) {
  if (!y_bf16 || !out || m <= 0 || n <= 0) {
    pvr_set_error("pvr_colsum_bf16: invalid argument");
    return PVR_ERR_ARG;
  }
  const int rpb = rows_per_block_for(m);
  dim3 grid((n + 31) / 32, (m + rpb - 1) / rpb);
  colsum_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __nv_bfloat16*>(y_bf16),
                                                                           ldy, m, n, rpb, out);
  PVR_LAUNCH_CHECK("pvr_colsum_bf16");
  return PVR_OK;
}

extern "C" int pvr_transpose_bf16(const void* in, int64_t ldi, int rows, int cols, void* out, int64_t ldo,
                                  void* stream_) {
  if (!in || !out || rows <= 0 || cols <= 0) {
    pvr_set_error("pvr_transpose_bf16: invalid argument");
    return PVR_ERR_ARG;
  }
  dim3 grid((rows + 31) / 32, (cols + 31) / 32);
  transpose_bf16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __nv_bfloat16*>(in), ldi, rows, cols, static_cast<__nv_bfloat16*>(out), ldo);
  PVR_LAUNCH_CHECK("pvr_transpose_bf16");
  return PVR_OK;
}

extern "C" int pvr_cast_weight(const float* w, int rows, int cols, void* w_bf16, int64_t ldb, void* wt_bf16,
                               int64_t ldt, void* stream_) {
  if (!w || rows <= 0 || cols <= 0 || (!w_bf16 && !wt_bf16)) {
    pvr_set_error("pvr_cast_weight: invalid argument");
    return PVR_ERR_ARG;
  }
  const bool vec = rows % 4 == 0 && cols % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (!w_bf16 || (ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(w_bf16) & 7) == 0)) &&
                   (!wt_bf16 || (ldt % 4 == 0 && (reinterpret_cast<uintptr_t>(wt_bf16) & 7) == 0));
  if (vec) {
    dim3 grid((cols + 63) / 64, (rows + 63) / 64);
    cast_weight64_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        w, rows, cols, static_cast<__nv_bfloat16*>(w_bf16), ldb, static_cast<__nv_bfloat16*>(wt_bf16), ldt);
  } else {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    cast_weight_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        w, rows, cols, static_cast<__nv_bfloat16*>(w_bf16), ldb, static_cast<__nv_bfloat16*>(wt_bf16), ldt);
  }
  PVR_LAUNCH_CHECK("pvr_cast_weight");
  return PVR_OK;
}

extern "C" int pvr_optim_sumsq(const float* const* grads, const int64_t* sizes, int count, double* sumsq,
                               void* stream_) {
  if (!grads || !sizes || !sumsq || count <= 0 || count > 32) {
    pvr_set_error("pvr_optim_sumsq: invalid argument (at most 32 tensors per call)");
    return PVR_ERR_ARG;
  }
  TensorList tl;
  tl.count = count;
  for (int i = 0; i < count; ++i) {
    tl.g[i] = const_cast<float*>(grads[i]);
    tl.n[i] = sizes[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  cudaMemsetAsync(sumsq, 0, sizeof(double), st);
  sumsq_kernel<<<dim3(148, count), 256, 0, st>>>(tl, sumsq);
  PVR_LAUNCH_CHECK("pvr_optim_sumsq");
  return PVR_OK;
}

static int optim_step_impl(const float* lr_dev, int mode, float* const* params, float* const* grads, float* const* state1,
                              float* const* state2, const int64_t* sizes, int count, const double* sumsq,
                              float grad_scale, float max_norm, float lr, float alpha_or_beta1, float beta2, float eps,
                              int step, float* norm_out, void* stream_) {
  if ((mode != PVR_OPT_RMSPROP && mode != PVR_OPT_ADAM) || !params || !grads || !state1 || !sizes || !sumsq ||
      count <= 0 || count > 32 || (mode == PVR_OPT_ADAM && !state2)) {
    pvr_set_error("pvr_optim_step: invalid argument (at most 32 tensors per call)");
    return PVR_ERR_ARG;
  }
  TensorList tl;
  tl.count = count;
  for (int i = 0; i < count; ++i) {
    tl.p[i] = params[i];
    tl.g[i] = grads[i];
    tl.s1[i] = state1[i];
    tl.s2[i] = state2 ? state2[i] : nullptr;
    tl.n[i] = sizes[i];
  }
  float bc1 = 1.f, bc2 = 1.f;
  if (mode == PVR_OPT_ADAM) {
    bc1 = 1.f - powf(alpha_or_beta1, (float)step);
    bc2 = 1.f - powf(beta2, (float)step);
  }
  optim_step_kernel<<<dim3(148, count), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      tl, sumsq, grad_scale, max_norm, mode, lr, alpha_or_beta1, beta2, eps, bc1, bc2, norm_out, lr_dev);
  PVR_LAUNCH_CHECK("pvr_optim_step");
  return PVR_OK;
}

extern "C" int pvr_optim_step(int mode, float* const* params, float* const* grads, float* const* state1,
                              float* const* state2, const int64_t* sizes, int count, const double* sumsq,
                              float grad_scale, float max_norm, float lr, float alpha_or_beta1, float beta2, float eps,
                              int step, float* norm_out, void* stream_) {
  return optim_step_impl(nullptr, mode, params, grads, state1, state2, sizes, count, sumsq, grad_scale, max_norm, lr,
                         alpha_or_beta1, beta2, eps, step, norm_out, stream_);
}

extern "C" int pvr_optim_step_dev(int mode, float* const* params, float* const* grads, float* const* state1,
                                  float* const* state2, const int64_t* sizes, int count, const double* sumsq,
                                  float grad_scale, float max_norm, const float* lr_dev, float alpha_or_beta1,
                                  float beta2, float eps, int step, float* norm_out, void* stream_) {
  if (!lr_dev || mode != PVR_OPT_RMSPROP) {
    pvr_set_error("pvr_optim_step_dev: needs a device learning rate and RMSprop (Adam's bias correction is host state)");
    return PVR_ERR_ARG;
  }
  return optim_step_impl(lr_dev, mode, params, grads, state1, state2, sizes, count, sumsq, grad_scale, max_norm, 0.f,
                         alpha_or_beta1, beta2, eps, step, norm_out, stream_);
}

// ---------------------------------------------------------------------------------------------- LSTM time loops
// The T sequential steps of one layer are issued from C. The launch sequence of a (layer, shape, buffers) tuple is
// captured once into a CUDA graph and replayed afterwards: 2T launches become one cudaGraphLaunch (the per-step
// kernels are a few microseconds each, so per-launch host cost would otherwise bound the step).
namespace {

struct GraphCacheEntry {
  unsigned char key[160];
  size_t key_len;
  int seen;  // number of eager executions so far (the first run stays eager: lazy one-time attribute setup)
  cudaGraphExec_t exec;
};
GraphCacheEntry g_graphs[128];
int g_num_graphs = 0;

GraphCacheEntry* graph_lookup(const void* key, size_t len) {
  for (int i = 0; i < g_num_graphs; ++i)
    if (g_graphs[i].key_len == len && memcmp(g_graphs[i].key, key, len) == 0) return &g_graphs[i];
  if (g_num_graphs == 128) {  // recycle everything (shapes changed many times)
    for (int i = 0; i < 128; ++i)
      if (g_graphs[i].exec) cudaGraphExecDestroy(g_graphs[i].exec);
    g_num_graphs = 0;
  }
  GraphCacheEntry* e = &g_graphs[g_num_graphs++];
  memset(e, 0, sizeof(*e));
  memcpy(e->key, key, len);
  e->key_len = len;
  return e;
}

int lstm_forward_issue(const pvr_lstm_fwd* L, cudaStream_t st) {
  const int T = L->T, B = L->B, H = L->H;
  const long long BH = (long long)B * H;
  __nv_bfloat16* hm = static_cast<__nv_bfloat16*>(L->hm);
  __nv_bfloat16* ho = static_cast<__nv_bfloat16*>(L->h_out);
  // Time chunks (PVR_LSTM_CONT_*): a chunk that continues an earlier one finds hm[0] written by that chunk's last
  // cell kernel; a chunk that is continued writes hm[T] / reads nd[T] (both arrays belong to the whole sequence).
  const bool cont_prev = (L->flags & PVR_LSTM_CONT_PREV) != 0, cont_next = (L->flags & PVR_LSTM_CONT_NEXT) != 0;
  if (!cont_prev) {
    mask_state_kernel<<<(int)((BH + 255) / 256), 256, 0, st>>>(L->h0, L->nd, B, H, hm);
    PVR_LAUNCH_CHECK("pvr_lstm_forward(mask)");
  }
  pvr_gemm_desc d;
  memset(&d, 0, sizeof(d));
  // The recurrent product is accumulated straight into the step's slice of the input projection (split-K slices,
  // fp32 atomics): twice as many CTAs stream W_hh (one wave of 128 instead of 64), no g_tmp round trip, and the cell
  // kernel reads one array instead of two. xp is consumed (overwritten) by the forward.
  const bool in_place = (H / 64) % 2 == 0 && getenv("PVR_LSTM_NO_INPLACE") == nullptr;
  d.b = L->w_hh; d.ldb = H; d.out = L->g_tmp; d.ldo = 4 * H; d.lda = H;
  d.m = B; d.n = 4 * H; d.n_pad = 4 * H; d.k = H; d.out_f32 = in_place ? 2 : 1; d.split_k = in_place ? 2 : 1;
  d.flags = PVR_GEMM_PDL;
  for (int t = 0; t < T; ++t) {
    d.a = hm + t * BH;
    float* xp_t = const_cast<float*>(L->xp) + (long long)t * B * 4 * H;
    if (in_place) d.out = xp_t;
    int rc = pvr_gemm(&d, st);
    if (rc != PVR_OK) return rc;
    launch_pdl(lstm_cell_fwd_kernel, (int)((BH + 255) / 256), 256, st, in_place ? xp_t : L->g_tmp,
               in_place ? (const float*)nullptr : (const float*)xp_t, L->c_all + t * BH, L->nd + (long long)t * B,
               (t + 1 < T || cont_next) ? L->nd + (long long)(t + 1) * B : nullptr, B, H,
               L->gates + (long long)t * B * 4 * H, L->c_all + (t + 1) * BH, L->h_last, ho + t * BH,
               (t + 1 < T || cont_next) ? hm + (t + 1) * BH : nullptr);
    PVR_LAUNCH_CHECK("pvr_lstm_forward(cell)");
  }
  return PVR_OK;
}

int lstm_backward_issue(const pvr_lstm_bwd* L, cudaStream_t st) {
  const int T = L->T, B = L->B, H = L->H;
  const long long BH = (long long)B * H;
  __nv_bfloat16* dG = static_cast<__nv_bfloat16*>(L->dG);
  pvr_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.b = L->w_hh_t; d.ldb = 4 * H; d.out = L->dh_rec; d.ldo = H; d.lda = 4 * H;
  d.m = B; d.n = H; d.n_pad = H; d.k = 4 * H; d.out_f32 = 2;
  d.split_k = (4 * H / 64) % 8 == 0 ? 8 : 1;
  d.flags = PVR_GEMM_PDL;
  // time chunks, processed last to first: dh_rec / dc_rec carry the recurrent gradient from one chunk to the one before
  const bool cont_prev = (L->flags & PVR_LSTM_CONT_PREV) != 0, cont_next = (L->flags & PVR_LSTM_CONT_NEXT) != 0;
  for (int t = T - 1; t >= 0; --t) {
    launch_pdl(lstm_cell_bwd_kernel, (int)((BH + 255) / 256), 256, st,
               L->dh_out ? L->dh_out + t * BH : nullptr, L->dh_rec, L->dc_rec, L->gates + (long long)t * B * 4 * H,
               L->c_all + t * BH, L->c_all + (t + 1) * BH, L->nd + (long long)t * B,
               (t + 1 < T || cont_next) ? L->nd + (long long)(t + 1) * B : nullptr, B, H,
               dG + (long long)t * B * 4 * H);
    PVR_LAUNCH_CHECK("pvr_lstm_backward(cell)");
    if (t > 0 || cont_prev) {
      d.a = dG + (long long)t * B * 4 * H;
      int rc = pvr_gemm(&d, st);
      if (rc != PVR_OK) return rc;
    }
  }
  return PVR_OK;
}

// Run `issue` through the graph cache: eager the first time a key is seen (and whenever the stream is already being
// captured by the caller), captured + instantiated the second time, replayed afterwards.
template <class Desc, class Fn>
int run_cached(const Desc* L, cudaStream_t st, Fn issue, const char* name) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return issue(L, st);
  unsigned char key[160];
  static_assert(sizeof(Desc) + sizeof(cudaStream_t) + sizeof(void*) <= sizeof(key), "graph key too small");
  memset(key, 0, sizeof(key));
  memcpy(key, L, sizeof(Desc));
  memcpy(key + sizeof(Desc), &st, sizeof(st));
  const void* tag = name;
  memcpy(key + sizeof(Desc) + sizeof(st), &tag, sizeof(tag));
  GraphCacheEntry* e = graph_lookup(key, sizeof(Desc) + sizeof(st) + sizeof(tag));
  if (e->exec) {
    cudaError_t err = cudaGraphLaunch(e->exec, st);
    if (err != cudaSuccess) {
      pvr_set_error("%s: cudaGraphLaunch: %s", name, cudaGetErrorString(err));
      return PVR_ERR_CUDA;
    }
    return PVR_OK;
  }
  if (e->seen++ == 0) return issue(L, st);
  // Capture on a private stream (the caller's may be the legacy default stream, which cannot be captured); the
  // instantiated graph is then launched on the caller's stream.
  static cudaStream_t cap = nullptr;
  if (!cap && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    return issue(L, st);
  }
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return issue(L, st);
  }
  const int rc = issue(L, cap);
  cudaError_t err = cudaStreamEndCapture(cap, &graph);
  if (rc != PVR_OK || err != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != PVR_OK) return rc;
    return issue(L, st);  // capture unavailable: stay eager
  }
  err = cudaGraphInstantiate(&e->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (err != cudaSuccess) {
    e->exec = nullptr;
    cudaGetLastError();
    return issue(L, st);
  }
  err = cudaGraphLaunch(e->exec, st);
  if (err != cudaSuccess) {
    pvr_set_error("%s: cudaGraphLaunch: %s", name, cudaGetErrorString(err));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}

}  // namespace

namespace pvr {  // lstm_persist.cu
int lstm_persist_supported(int T, int B, int H);
int lstm_persist_forward(const pvr_lstm_fwd* L, cudaStream_t st);
int lstm_persist_backward(const pvr_lstm_bwd* L, cudaStream_t st);
}  // namespace pvr

extern "C" int pvr_lstm_forward(const pvr_lstm_fwd* L, void* stream_) {
  if (!L || L->T <= 0 || L->B <= 0 || L->H <= 0 || L->H % 64 || !L->w_hh || !L->xp || !L->nd || !L->h0 || !L->c_all ||
      !L->hm || !L->h_out || !L->gates || !L->g_tmp || !L->h_last) {
    pvr_set_error("pvr_lstm_forward: invalid argument");
    return PVR_ERR_ARG;
  }
  // whole sequence in one call: the persistent kernel (lstm_persist.cu) when the shape fits the device
  if (L->flags == 0 && pvr::lstm_persist_supported(L->T, L->B, L->H))
    return pvr::lstm_persist_forward(L, static_cast<cudaStream_t>(stream_));
  return run_cached(L, static_cast<cudaStream_t>(stream_), lstm_forward_issue, "pvr_lstm_forward");
}

extern "C" int pvr_lstm_backward(const pvr_lstm_bwd* L, void* stream_) {
  if (!L || L->T <= 0 || L->B <= 0 || L->H <= 0 || L->H % 64 || !L->w_hh_t || !L->nd || !L->gates || !L->c_all ||
      !L->dh_rec || !L->dc_rec || !L->dG) {
    pvr_set_error("pvr_lstm_backward: invalid argument");
    return PVR_ERR_ARG;
  }
  if (L->flags == 0 && pvr::lstm_persist_supported(L->T, L->B, L->H))
    return pvr::lstm_persist_backward(L, static_cast<cudaStream_t>(stream_));  // accumulates dbias itself
  int rc = run_cached(L, static_cast<cudaStream_t>(stream_), lstm_backward_issue, "pvr_lstm_backward");
  if (rc == PVR_OK && L->dbias)  // per-step kernels: the bias gradient is a column sum over this call's dG rows
    rc = pvr_colsum_bf16(L->dG, 4 * L->H, L->T * L->B, 4 * L->H, L->dbias, stream_);
  return rc;
}
