// Patch im2col for ViT patch sizes that the TMA-side patch embedding of clip_vit.ViTRunner cannot express: it views one
// row of a patch as p * 4 NHWC4 elements and needs that to be a whole number of 128-byte swizzle rows (p = 16 / 32).
// mae_huge (src/vision_models/mae.py:291-296) has p = 14: its 0.7 % of the encoder's FLOPs go through this gather
// (HBM bound: 0.4 MB read + 0.33 MB written per 224 x 224 frame) and one plain GEMM with K = 3 p^2 padded to 640.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

// one thread per pair of adjacent K indices (k_pad is even): a 4-byte (bf16) / 8-byte (f32) store
template <typename T>
__global__ void __launch_bounds__(256) vit_patchify_kernel(const T* __restrict__ x, long long pairs, int res, int p,
                                                           int grid, int k_pad, T* __restrict__ col) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= pairs) return;
  const int kh = k_pad >> 1;
  const long long row = i / kh;
  const int k0 = (int)(i - row * kh) * 2;
  const int gx = (int)(row % grid);
  const long long t = row / grid;
  const int gy = (int)(t % grid);
  const long long img = t / grid;
  T v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = k0 + e;
    T val = T(0.f);
    if (k < 3 * p * p) {
      const int c = k / (p * p), r = k - c * p * p;
      const int py = r / p, px = r - py * p;
      val = x[((img * res + gy * p + py) * res + gx * p + px) * 4 + c];
    }
    v[e] = val;
  }
  col[row * k_pad + k0] = v[0];
  col[row * k_pad + k0 + 1] = v[1];
}

}  // namespace
}  // namespace pvr

extern "C" int pvr_vit_patchify(const void* x_nhwc4, int n_img, int res, int patch, int k_pad, int f32, void* col,
                                void* stream) {
  if (!x_nhwc4 || !col || n_img <= 0 || res <= 0 || patch <= 0 || res % patch || k_pad < 3 * patch * patch ||
      (k_pad & 1)) {
    pvr_set_error("pvr_vit_patchify: invalid argument");
    return PVR_ERR_ARG;
  }
  const int grid = res / patch;
  const long long pairs = (long long)n_img * grid * grid * (k_pad / 2);
  const long long blocks = (pairs + 255) / 256;
  if (blocks > 0x7fffffffll) {
    pvr_set_error("pvr_vit_patchify: too many patches for one launch");
    return PVR_ERR_ARG;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (f32)
    pvr::vit_patchify_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(static_cast<const float*>(x_nhwc4), pairs, res,
                                                                       patch, grid, k_pad, static_cast<float*>(col));
  else
    pvr::vit_patchify_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(x_nhwc4), pairs, res, patch, grid, k_pad, static_cast<__nv_bfloat16*>(col));
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pvr_set_error("pvr_vit_patchify: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}
