// Host-visible declarations of the tcgen05 implicit-GEMM convolution / GEMM kernel.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvr {

// How the A operand (activations) reaches shared memory.
enum AMode : int {
  A_TILED = 0,     // plain (M x K) row-major matrix: 2-D tiled TMA, 128B swizzle (1x1 stride-1 convs, GEMMs)
  A_IM2COL64 = 1,  // NHWC tensor, C_in % 64 == 0: one im2col TMA (64 channels x 128 pixels) per K chunk, 128B swizzle
  A_IM2COL8 = 2,   // NHWC tensor with 8-element pixels: 8 im2col TMAs (8 ch x 128 px) per K chunk, no swizzle
  A_IM2COL32 = 3,  // NHWC tensor with 32-element pixels (W-expanded stem input): 2 im2col TMAs (32 ch x 128 px)
                   // per K chunk, 64B swizzle
};

struct ConvGemmParams {
  int M;             // GEMM rows = images * P * Q (output pixels)
  int P, Q;          // output height / width per image (im2col modes)
  int num_m_tiles;   // ceil(M / 128)
  int num_n_tiles;   // n_pad / BLOCK_N
  int num_k_chunks;  // K chunks of 64 per tile (= k_pad / 64 / split_k)
  int cin_chunks;    // A_IM2COL64: C_in / 64
  int S;             // filter taps per row
  int taps;          // R * S (real taps; chunks may be padded with repeats of tap 0 against zero weights)
  int stride_w, stride_h;
  int lower_w, lower_h;
  int n_valid;       // real output channels (columns >= n_valid are not stored)
  int relu_n;        // ReLU is applied to output columns < relu_n (0: none, >= n_valid: all)
  int has_res;       // residual present (EPI_TMA path reads it through tmap_res)
  int elu;           // ELU(alpha = 1) on every output channel (direct epilogue; small-conv PVR)
  int quick_gelu;    // bf16 output, after bias: 1 = x * sigmoid(1.702 x) (CLIP MLP), 2 = erf GELU (MAE / timm MLP)
  int res_mode;      // 0: out += res; 1: out = res > 0 ? out : 0 (ReLU backward with the saved activation)
  int reverse;       // 1: walk the (m, n) tiles last to first (zig-zag over consecutive layers, L2 reuse)
  int pdl;           // 1: launch with programmatic stream serialization (prologue overlaps the previous kernel)
  int n_tiles_shift; // log2(num_n_tiles) when it is a power of two, else -1 (set by launch_conv_gemm)
  int kc_split;      // > 0: K chunks >= kc_split come from a second input (tmap_a2), A_TILED only: projection shortcut
  int a2_im2col;     // second input is read with im2col-mode 1x1 taps at stride a2_stride (else a plain 2-D matrix)
  int a2_stride;
  int mn;            // 1: MN-major operands (A is (K, M), W is (K, N) row-major): out = A^T W, fp32 output
  int cta2;          // 1: CTA-pair kernel (cta_group::2): num_m_tiles counts 256-row tiles, W map box = BLOCK_N / 2 rows
  int split_k;       // >= 1; K is cut into split_k slices of num_k_chunks chunks, each its own tile (fp32 atomics)
  int out_is_f32;    // fp32 output: TMA-staged (tmap_out is an fp32 map) or, with !epi_tma, atomically accumulated
  float* out_f32;    // direct fp32 accumulate target (split-K)
  int out_coff, res_coff;  // channel offsets of the output / residual tile inside their pixel rows (EPI_TMA)
  long long ldo, ldr;  // output / residual row pitch in elements
  __nv_bfloat16* out;
  const __nv_bfloat16* res;  // may be null
  const float* scale;        // (n_pad)
  const float* bias;         // (n_pad)
  // 1: A-stationary tile order (plain 1x1 layers whose K chunks exactly fill the stage ring, e.g. 256 -> 1024 +
  // residual of layer3): CTA w walks m-tiles w, w + grid, ... and ALL n-tiles of one m-tile back to back, so K chunk
  // kc of the A tile is already in stage kc from the previous n-tile and only the W chunk is reloaded — 96 KB instead
  // of 160 KB taken in per 128 x 128 tile by a kernel that is bound by the SM's ingest rate. num_m_tiles is padded to
  // a multiple of the grid; rows past M are zero-filled by the TMA loads and clipped by the TMA stores.
  int astat;
};

// Stages of the TMA ring of the single-CTA kernel (host-side twin of Cfg<BLOCK_N, EPI_TMA>::STAGES).
int conv_gemm_stages(int block_n, bool epi_tma);

// Launch on `stream`; block_n in {32, 64, 128, 256}. Returns cudaError_t of the launch.
// epi_tma: stage the output through shared memory + TMA store (needs n_valid % 64 == 0, block_n >= 64); tmap_out /
// tmap_res are (pitch x M) bf16 maps with a (64 x 128) box over the output / residual pixel rows.
cudaError_t launch_conv_gemm(int block_n, int a_mode, bool epi_tma, const CUtensorMap& tmap_a,
                             const CUtensorMap& tmap_b, const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                             const ConvGemmParams& p, int num_sms, cudaStream_t stream, const CUtensorMap* tmap_a2 = nullptr);

// Tensor-map builders (driver entry points resolved through cudaGetDriverEntryPoint; no -lcuda needed).
// 2-D K-major bf16 matrix (rows x k), row pitch ld elements, box = (64 x box_rows), 128B swizzle.
bool make_tmap_2d(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint32_t box_rows,
                  const char** err);
// The same matrix as (64, rows, k / 64) with a (64, box_rows, chunks) box: `chunks` consecutive K chunks per instruction.
bool make_tmap_kchunks(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint32_t box_rows,
                       uint32_t chunks, const char** err);
// 2-D fp32 matrix (rows x cols), row pitch ld elements, box = (32 x box_rows) = 128-byte rows, 128B swizzle.
bool make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t ld,
                      uint32_t box_rows, const char** err);
// 2-D K-major bf16 matrix with a (32 x box_rows) box and 64B swizzle (stem tap matrices).
bool make_tmap_2d_sw64(CUtensorMap* out, const void* base, uint64_t k, uint64_t rows, uint64_t ld, uint32_t box_rows,
                       const char** err);
// Tiled 4-D map over an NHWC bf16 tensor (N, H, W, pitch) exposing `c` channels: box = (64 ch, box_w, box_h, 1),
// 128B swizzle (patch-resident 3x3 convolution, conv3x3_patch.cu).
bool make_tmap_4d(CUtensorMap* out, const void* base, int c, int pitch, int w, int h, int n, int box_w, int box_h,
                  const char** err, int stride_h = 1);
// Stem input in the compact padded layout (PVR_FMT_STEM_PAD_BF16): overlapping 64-byte windows, 16 bytes apart.
bool make_tmap_stem_compact(CUtensorMap* out, const void* base, int w_out, int h_in, int n, int box_w, int box_h,
                            int stride_h, const char** err);
// 3x3 / stride 1 / pad 1 convolution, C_in = C_out = 64, W % 8 == 0, with the input patch resident in shared memory
// (three column-shifted copies; the nine taps are UMMA descriptor offsets) and the weights resident for the whole
// kernel. Same epilogue contract as conv_gemm (scale/bias/ReLU, bf16 NHWC out).
struct Conv3x3PatchParams {
  int n_img, P, Q;         // images, height, width (output == input size)
  int tiles_p, tiles_q;    // ceil(P / 16), Q / 8
  int relu;
  int reverse;             // 1: tiles walked last to first
  int pdl;                 // 1: launch with programmatic stream serialization
  int stem;                // 1: 7x7/s2 stem over the W-expanded input (P, Q = output size), 0: 3x3/s1 64->64
  const float* scale;      // (64)
  const float* bias;       // (64)
  // stem only: fused 3x3/s2/pad-1 max pool. pool_out != nullptr: tiles_p = ceil(pool_P / 7), tiles_q = ceil(pool_Q / 3)
  // and the pooled NHWC (n_img, pool_P, pool_Q, 64) tensor is the only output.
  __nv_bfloat16* pool_out;
  int pool_P, pool_Q;
};
cudaError_t launch_conv3x3_patch(const CUtensorMap& tmap_in, const CUtensorMap& tmap_w, const CUtensorMap& tmap_out,
                                 const Conv3x3PatchParams& p, int num_sms, cudaStream_t stream);
// Back-to-back fusion of `out = relu(bn3(conv3 t2) + x)` (1x1, 64 -> 256) with the next block's
// `t1 = relu(bn1(conv1 out))` (1x1, 256 -> n2), conv_b2b.cu.
struct ConvB2BParams {
  int M, num_m_tiles;  // pixels, ceil(M / 128)
  int n2;              // 64 or 128
  int streamed;        // 1: layer2 variant (128 -> 512 + residual, then 512 -> 128), weights streamed through rings
  int k1_chunks;       // 1: conv3 over t2 (+ residual); 2: [t2 | x] projection-shortcut GEMM, no residual
  int reverse, pdl;
  const float* scale1; // (256) conv3's folded BN
  const float* bias1;
  const float* scale2; // (n2) the next conv1's folded BN
  const float* bias2;
};
cudaError_t launch_conv_b2b(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tw3,
                            const CUtensorMap& tres, const CUtensorMap& tout, const CUtensorMap& tw1,
                            const CUtensorMap& tout2, const ConvB2BParams& p, int num_sms, cudaStream_t stream);
// im2col map over an NHWC bf16 tensor (N, H, W, pitch) exposing `c` channels per pixel.
bool make_tmap_im2col(CUtensorMap* out, const void* base, int c, int pitch, int w, int h, int n, int lower_w,
                      int lower_h, int upper_w, int upper_h, int stride_w, int stride_h, int channels_per_pixel,
                      int pixels_per_column, int swizzle_bytes, const char** err);

// First layer of the small-conv trunk (3 -> 32 channels, 3x3 stride 2, ELU) with register-resident mma.sync fragments (small_conv.cu). `x` is
// (F, Hi, Wi, 4) bf16, `wpk` the packed weight of program.pack_first_small_conv, `y` (F, Ho, Wo, 32) bf16.
cudaError_t launch_small_conv1(const void* x, int F, int Hi, int Wi, int Ho, int Wo, const void* wpk, const float* scale,
                               const float* bias, void* y, cudaStream_t stream);

// 2x2 / stride 2 average pool NHWC -> NHWC (clip_rn.cu); f32 != 0: float32 tensors (parity mode), else bf16.
cudaError_t launch_avgpool2(const void* in, void* out, int n_img, int H, int W, int C, int f32, cudaStream_t stream);

}  // namespace pvr
