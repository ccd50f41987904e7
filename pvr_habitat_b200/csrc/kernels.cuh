// Launchers of the non-GEMM kernels (pool_head.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvr {

cudaError_t launch_maxpool(const __nv_bfloat16* in, __nv_bfloat16* out, int n_img, int H, int W, int C, int P, int Q,
                           cudaStream_t stream);
cudaError_t launch_avgpool(const __nv_bfloat16* in, float* emb, long long emb_ld, int emb_off, int n_img, int HW,
                           int C, cudaStream_t stream);
cudaError_t launch_head_tail(const __nv_bfloat16* t, int pitch, const float* aux, float* emb, long long emb_ld,
                             int emb_off, int n_img, int H, int W, int c, cudaStream_t stream);
cudaError_t launch_head_tail_taps(const float* z, int zp, const float* aux, float* emb, long long emb_ld, int emb_off,
                                  int n_img, int H, int W, int c, cudaStream_t stream);

cudaError_t launch_flatten(const __nv_bfloat16* in, int pitch, float* emb, long long emb_ld, int emb_off, int n_img,
                           int HW, int C, cudaStream_t stream);

// ---- fp32 parity mode (conv_f32.cu)
struct ConvF32Params {
  const float* in;     // NHWC float32, pixel pitch in_pitch
  const float* w;      // (N, R, S, C) float32, dense
  const float* scale;  // (N) folded BN scale
  const float* bias;
  const float* res;    // optional residual (M, res_pitch), channel offset res_coff
  float* out;          // (M, out_pitch), channel offset out_coff
  long long M;         // n_images * P * Q
  int N, C, H, W, P, Q, R, S, stride_h, stride_w, lower_h, lower_w;
  int in_pitch, out_pitch, out_coff, res_pitch, res_coff, relu_n;
  int elu;             // activation after bias / residual: 0 none, 1 ELU, 2 QuickGELU, 3 erf GELU
};
cudaError_t launch_conv_f32(const ConvF32Params& p, cudaStream_t stream);
cudaError_t launch_maxpool_f32(const float* in, float* out, int n_img, int H, int W, int C, int P, int Q,
                               cudaStream_t stream);
cudaError_t launch_avgpool_f32(const float* in, float* emb, long long emb_ld, int emb_off, int n_img, int HW, int C,
                               cudaStream_t stream);
cudaError_t launch_flatten_f32(const float* in, int pitch, float* emb, long long emb_ld, int emb_off, int n_img, int HW,
                               int C, cudaStream_t stream);
cudaError_t launch_head_tail_f32(const float* t, int pitch, const float* aux, float* emb, long long emb_ld, int emb_off,
                                 int n_img, int H, int W, int c, cudaStream_t stream);

}  // namespace pvr
