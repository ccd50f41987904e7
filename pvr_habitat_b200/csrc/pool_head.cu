// Memory-bound companions of the conv kernel: 3x3/s2 max pool (torchvision/models/resnet.py:271 via
// src/vision_models/moco.py:11), global average pool (resnet.py:278-279) and the tail of the compression
// BasicBlock (moco.py:44-50, 88-94: conv2 -> bn2 -> += downsample(x) -> ReLU, flattened NCHW like
// src/embeddings.py:398 `out.view(-1, out_size)`).
#include "kernels.cuh"

namespace pvr {

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    f[2 * t + 0] = __uint_as_float(w[t] << 16);
    f[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
  }
}

// NHWC bf16, C % 8 == 0. One thread = 8 channels of one output pixel.
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in,
                                                            __nv_bfloat16* __restrict__ out, int n_img, int H, int W,
                                                            int C, int P, int Q) {
  const int cg = C >> 3;
  const long long total = (long long)n_img * P * Q * cg;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % cg);
    long long t = idx / cg;
    const int q = (int)(t % Q);
    t /= Q;
    const int pp = (int)(t % P);
    const int img = (int)(t / P);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = pp * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = q * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + (((long long)img * H + h) * W + w) * C) + g);
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
      }
    }
    uint4 o;
    __nv_bfloat162 a = __floats2bfloat162_rn(m[0], m[1]), b = __floats2bfloat162_rn(m[2], m[3]),
                   c = __floats2bfloat162_rn(m[4], m[5]), d = __floats2bfloat162_rn(m[6], m[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c);
    o.w = *reinterpret_cast<uint32_t*>(&d);
    reinterpret_cast<uint4*>(out + (((long long)img * P + pp) * Q + q) * C)[g] = o;
  }
}

// (n_img, HW, C) bf16 -> emb[img * emb_ld + emb_off + c] fp32 mean. One thread = 8 channels of one image.
__global__ void __launch_bounds__(256) avgpool_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ emb,
                                                       long long emb_ld, int emb_off, int n_img, int HW, int C) {
  const int cg = C >> 3;
  const long long total = (long long)n_img * cg;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % cg);
  const int img = (int)(idx / cg);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const uint4* src = reinterpret_cast<const uint4*>(in + (long long)img * HW * C) + g;
  for (int px = 0; px < HW; ++px) {
    float f[8];
    unpack8(__ldg(src + (long long)px * cg), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  float* dst = emb + (long long)img * emb_ld + emb_off + g * 8;
  const float inv = (float)HW;
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[j] = acc[j] / inv;
}

// Compression-head tail. `t` holds, per pixel (pitch elements): channels [0,c) = relu(bn1(conv1(x))) and
// [c,2c) = bn_d(conv_d(x) + b_d) from the merged tcgen05 GEMM. aux = w2[c][3][3][c] | scale2[c] | bias2[c] (fp32).
// out[img, co*H*W + y*W + x] = relu(scale2[co] * conv3x3(t[:, :c])[co] + bias2[co] + t[c + co]).
__global__ void __launch_bounds__(256) head_tail_kernel(const __nv_bfloat16* __restrict__ t, int pitch,
                                                         const float* __restrict__ aux, float* __restrict__ emb,
                                                         long long emb_ld, int emb_off, int H, int W, int c) {
  extern __shared__ float hs[];
  float* w2 = hs;                 // c*9*c
  float* sc = w2 + c * 9 * c;     // c
  float* bi = sc + c;             // c
  float* act = bi + c;            // H*W*2c
  const int img = blockIdx.x;
  const int HW = H * W;
  for (int i = threadIdx.x; i < c * 9 * c + 2 * c; i += blockDim.x) hs[i] = aux[i];
  for (int i = threadIdx.x; i < HW * 2 * c; i += blockDim.x) {
    const int px = i / (2 * c), ch = i - px * 2 * c;
    act[i] = __bfloat162float(t[((long long)img * HW + px) * pitch + ch]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < c * HW; o += blockDim.x) {
    const int co = o / HW, px = o - co * HW;
    const int y = px / W, x = px - y * W;
    float acc = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int yy = y - 1 + r;
      if (yy < 0 || yy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int xx = x - 1 + s;
        if (xx < 0 || xx >= W) continue;
        const float* a = act + (yy * W + xx) * 2 * c;
        const float* wv = w2 + ((co * 3 + r) * 3 + s) * c;
        for (int ci = 0; ci < c; ++ci) acc = fmaf(a[ci], wv[ci], acc);
      }
    }
    float v = fmaf(acc, sc[co], bi[co]) + act[px * 2 * c + c + co];
    emb[(long long)img * emb_ld + emb_off + o] = fmaxf(v, 0.f);
  }
}

// Compression-head tail over per-tap partial sums. `z` holds, per pixel (zp floats): 9 taps x 2c values
// z[q][tap*2c + j] = W_tap[j] . x[q] from the 1x1 tcgen05 GEMM (the 3x3 convolution's input is read once, not nine
// times). Step 1 gathers t[p][j] = scale1[j] * sum_tap z[p + tap - 1][tap*2c + j] + bias1[j] (ReLU on j < c) into
// shared memory, channel-major with an odd pixel pitch (conflict-free for both steps); step 2 is head_tail_kernel's
// 3x3 conv + BN + identity add + ReLU. aux = w2t[9][c][cp] (tap, input channel, output channel padded to cp = 4k) |
// scale2[c] | bias2[c] | scale1[2c] | bias1[2c], padded with zeros to a multiple of 4 floats (+4).
__global__ void __launch_bounds__(256) head_tail_taps_kernel(const float* __restrict__ z, int zp,
                                                              const float* __restrict__ aux, float* __restrict__ emb,
                                                              long long emb_ld, int emb_off, int n_img, int H, int W,
                                                              int c) {
  extern __shared__ __align__(16) float hs[];
  const int HW = H * W, HWp = HW | 1, c2 = 2 * c, cp = (c + 3) & ~3;
  float* w2t = hs;                 // [9][c][cp]: tap-major, output channel fastest (float4 of 4 output channels)
  float* sc2 = w2t + 9 * c * cp;   // c
  float* bi2 = sc2 + c;            // c
  float* sc1 = bi2 + c;            // 2c
  float* bi1 = sc1 + c2;           // 2c
  float* act = bi1 + c2 + ((4 - ((6 * c) & 3)) & 3);  // [2c][HWp]
  // weights + BN vectors once per (persistent) CTA: aux is already laid out [9][c][cp] | scale2 | bias2 | scale1 | bias1
  {
    const int n4 = (9 * c * cp + 6 * c) >> 2;  // 6c is even; the host pads the block to a multiple of 4 floats
    const float4* src = reinterpret_cast<const float4*>(aux);
    float4* dst = reinterpret_cast<float4*>(hs);
    for (int i = threadIdx.x; i < n4 + 1; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();
  for (int img = blockIdx.x; img < n_img; img += gridDim.x) {
  const float* zi = z + (long long)img * HW * zp;
#pragma unroll 2
  for (int i = threadIdx.x; i < HW * c2; i += blockDim.x) {
    const int px = i / c2, j = i - px * c2;  // adjacent threads read adjacent floats of one pixel's tap block
    const int y = px / W, x = px - y * W;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = y - 1 + r;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int xx = x - 1 + s;
        if (xx < 0 || xx >= W) continue;
        acc += __ldg(zi + (long long)(yy * W + xx) * zp + (r * 3 + s) * c2 + j);
      }
    }
    float v = fmaf(acc, __ldg(aux + 9 * c * cp + 2 * c + j), __ldg(aux + 9 * c * cp + 2 * c + c2 + j));
    if (j < c) v = fmaxf(v, 0.f);
    act[j * HWp + px] = v;
  }
  __syncthreads();
  // conv2: one thread = one pixel x four output channels (one activation load and one float4 weight load per 4 FMAs)
  const int groups = cp >> 2;
  for (int o = threadIdx.x; o < groups * HW; o += blockDim.x) {
    const int g = o / HW, px = o - g * HW;
    const int y = px / W, x = px - y * W;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int yy = y - 1 + r;
      if (yy < 0 || yy >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int xx = x - 1 + s;
        if (xx < 0 || xx >= W) continue;
        const float* a = act + yy * W + xx;
        const float4* wv = reinterpret_cast<const float4*>(w2t + (r * 3 + s) * c * cp) + g;
#pragma unroll 6
        for (int ci = 0; ci < c; ++ci) {
          const float av = a[ci * HWp];
          const float4 w4 = wv[ci * groups];
          acc0 = fmaf(av, w4.x, acc0);
          acc1 = fmaf(av, w4.y, acc1);
          acc2 = fmaf(av, w4.z, acc2);
          acc3 = fmaf(av, w4.w, acc3);
        }
      }
    }
    const float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int co = g * 4 + e;
      if (co < c) {
        const float v = fmaf(accs[e], sc2[co], bi2[co]) + act[(c + co) * HWp + px];
        emb[(long long)img * emb_ld + emb_off + co * HW + px] = fmaxf(v, 0.f);
      }
    }
  }
  __syncthreads();  // act is rewritten for the next image
  }
}

// emb[img][off + c*HW + px] = in[img][px][c]: NHWC bf16 -> NCHW-flattened fp32 (the small-conv PVR's output order).
__global__ void __launch_bounds__(256) flatten_kernel(const __nv_bfloat16* __restrict__ in, int pitch,
                                                       float* __restrict__ emb, long long emb_ld, int emb_off,
                                                       int n_img, int HW, int C) {
  const long long total = (long long)n_img * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(i % HW);
    const int c = (int)((i / HW) % C);
    const long long img = i / ((long long)HW * C);
    emb[img * emb_ld + emb_off + (long long)c * HW + px] = __bfloat162float(in[(img * HW + px) * pitch + c]);
  }
}

}  // namespace

cudaError_t launch_maxpool(const __nv_bfloat16* in, __nv_bfloat16* out, int n_img, int H, int W, int C, int P, int Q,
                           cudaStream_t stream) {
  const long long total = (long long)n_img * P * Q * (C >> 3);
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  maxpool3x3s2_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, n_img, H, W, C, P, Q);
  return cudaGetLastError();
}

cudaError_t launch_avgpool(const __nv_bfloat16* in, float* emb, long long emb_ld, int emb_off, int n_img, int HW,
                           int C, cudaStream_t stream) {
  const long long total = (long long)n_img * (C >> 3);
  avgpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, emb, emb_ld, emb_off, n_img, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_flatten(const __nv_bfloat16* in, int pitch, float* emb, long long emb_ld, int emb_off, int n_img,
                           int HW, int C, cudaStream_t stream) {
  const long long total = (long long)n_img * HW * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  flatten_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, pitch, emb, emb_ld, emb_off, n_img, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_head_tail(const __nv_bfloat16* t, int pitch, const float* aux, float* emb, long long emb_ld,
                             int emb_off, int n_img, int H, int W, int c, cudaStream_t stream) {
  const size_t smem = (size_t)(c * 9 * c + 2 * c + H * W * 2 * c) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  head_tail_kernel<<<n_img, 256, smem, stream>>>(t, pitch, aux, emb, emb_ld, emb_off, H, W, c);
  return cudaGetLastError();
}

cudaError_t launch_head_tail_taps(const float* z, int zp, const float* aux, float* emb, long long emb_ld, int emb_off,
                                  int n_img, int H, int W, int c, cudaStream_t stream) {
  const size_t smem = (size_t)(9 * c * ((c + 3) & ~3) + 6 * c + 4 + 2 * c * ((H * W) | 1)) * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(head_tail_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  // persistent: the weights (66 KB at c = 42) are staged once per CTA; as many CTAs per SM as shared memory allows
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
  const int grid = n_img < per_sm * 148 ? n_img : per_sm * 148;
  head_tail_taps_kernel<<<grid, 256, smem, stream>>>(z, zp, aux, emb, emb_ld, emb_off, n_img, H, W, c);
  return cudaGetLastError();
}

}  // namespace pvr
