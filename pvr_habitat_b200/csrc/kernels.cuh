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

}  // namespace pvr
