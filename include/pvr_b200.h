/*
 * pvr_b200 — C ABI of the B200-native PVR-embed -> BC-train hot path.
 *
 * The reference (sparisi/pvr_habitat) is pure Python: its "plugin interface" for this path is the Python class API
 * (EmbeddingNet / PolicyNet, src/embeddings.py:339-402, src/models.py:13-89). The Python mirror of those classes in
 * pvr_habitat_b200/ binds exactly the entry points below through ctypes.
 *
 * Conventions: every function returns 0 on success and a negative code on failure (pvr_last_error() gives the
 * message, thread-local); nothing throws across the ABI. All buffers are caller-owned DEVICE pointers unless a
 * parameter says "host". `stream` is a cudaStream_t passed as void*. No function synchronises the device and none
 * has a CPU fallback: without a CUDA device every compute entry point fails with PVR_ERR_CUDA.
 */
#ifndef PVR_B200_H
#define PVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVR_OK 0
#define PVR_ERR_ARG (-1)
#define PVR_ERR_CUDA (-2)
#define PVR_ERR_STATE (-3)

/* Message of the last failure on this thread ("" if none). */
const char* pvr_last_error(void);
/* ABI version of the loaded library (bumped on every signature change). */
int pvr_abi_version(void);

/* ------------------------------------------------------------------------------------------------------------
 * K1 — fused uint8 HWC decode + Resize + CenterCrop + /255 + Normalize + frame split.
 * Replaces src/embeddings.py:80-85 (transforms) + :391-394 (EmbeddingNet.forward prologue) and the host-side
 * np.split/np.concatenate frame split of main_bc_1.py:134 / behavioral_cloning/save_embedded_obs.py:153.
 *
 *   in   : (N, H, W, 3*n_frames) uint8, frame f = channels [3f, 3f+3)
 *   out  : image index f*N + i (frame-major, the reference's intermediate order) or, with sample_major != 0,
 *          i*n_frames + f (so that an (N, n_frames*O) embedding row is contiguous):
 *          PVR_FMT_NCHW_F32       (n_frames*N, 3, crop, crop) float32 — bit-exact with the reference transforms
 *          PVR_FMT_NHWC4_BF16     (n_frames*N, crop, crop, 4) bf16, channel 3 = 0 — round-to-nearest of the above
 *          PVR_FMT_STEM_BF16      (n_frames*N, crop, crop/2, 8, 4) bf16: for output column q of the 7x7/2 stem the
 *                                 8 input columns 2q-3..2q+4 (zeros outside the image) x 4 channels = 64 B, so the
 *                                 stem is a 7x1-tap implicit GEMM fed by 64-byte TMA rows
 *   rh,rw: size after Resize (short side -> 256: rh = 256, rw = int(256*W/H) for H <= W), bilinear,
 *          align_corners=False, no antialias, rounded half-to-even back to uint8 (torchvision semantics)
 *   top,left: CenterCrop offsets inside the resized image; crop = 224
 *   mean,stdv: host float[3]
 */
#define PVR_FMT_NCHW_F32 0
#define PVR_FMT_NHWC4_BF16 1
#define PVR_FMT_STEM_BF16 2
#define PVR_FMT_NHWC4_F32 3 /* (n*N, crop, crop, 4) float32, RGB + zero pad: input of the fp32 parity mode */
/* (n*N, crop, crop + 8, 4) bf16: NHWC4 rows with 3 zero pixels before column 0 and 5 after the last one (row pitch
 * (crop + 8) * 8 bytes). The 7x7/2 stem reads it through a tensor map whose pixel dimension advances by TWO pixels
 * (16 bytes) while each "pixel" spans 8 columns x 4 channels (64 bytes): the W-expansion of PVR_FMT_STEM_BF16 happens
 * in the TMA addressing instead of in HBM — 0.42 MB per frame instead of 1.6 MB written here and read by the stem. */
#define PVR_FMT_STEM_PAD_BF16 4
/* OR-ed into out_fmt: Resize is bicubic (A = -0.75, no antialias, clamped to [0, 255] before the rounding cast) instead
 * of bilinear — T.Resize(256, interpolation=3) of the MAE encoders, src/embeddings.py:81. */
#define PVR_RESIZE_BICUBIC 0x100
/* OR-ed into out_fmt (bilinear only) — the 'maskrcnn_l3' transforms of src/embeddings.py:283-294, which resize a FLOAT
 * image: the bilinear interpolant is not rounded back to uint8 and not divided by 255; out = (v - mean) / stdv with
 * mean = (103.530, 116.280, 123.675), stdv = 1 on values in [0, 255]. */
#define PVR_RESIZE_FLOAT 0x200
/* OR-ed into out_fmt: rows 0 and 2 of every image trade places before the resize. That is what the reference's
 * `_rgb_to_bgr` does (src/embeddings.py:285-288): `x[:,:,[0,1,2]] = x[:,:,[2,1,0]]` indexes dimension 2 of the NCHW
 * tensor — the rows — so the channels stay RGB and three rows are permuted; reproduced as is for parity (with
 * CenterCrop(224) of a 256-row resize the rows only reach the output for frames of fewer than 54 rows). */
#define PVR_SWAP_ROWS_0_2 0x400
int pvr_preprocess_u8(const uint8_t* in, int N, int H, int W, int n_frames, int rh, int rw, int top, int left,
                      int crop, const float* mean, const float* stdv, void* out, int out_fmt, int sample_major,
                      void* stream);

/* Same contract with an ANTIALIASED bicubic Resize (a = -0.5, separable, float32 intermediate): CLIP's transforms,
 * T.Resize(res, BICUBIC, antialias=True) -> CenterCrop(res), src/embeddings.py:309-314 (mean / stdv = CLIP's). Output
 * formats PVR_FMT_NCHW_F32, PVR_FMT_NHWC4_BF16, PVR_FMT_NHWC4_F32. The first call for an (input size, output size)
 * pair allocates and fills two small weight tables (not capturable into a CUDA graph); down-scaling up to 7.5x.
 * Round-1 status: arithmetic core verified on CPU against the oracle, kernel not yet run on a GPU (experimental). */
int pvr_preprocess_u8_aa(const uint8_t* in, int N, int H, int W, int n_frames, int rh, int rw, int top, int left,
                         int crop, const float* mean, const float* stdv, void* out, int out_fmt, int sample_major,
                         void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Encoder "program": the frozen PVR network (ResNet-50 conv5 / l4 / l3 compressed / uber, src/vision_models/
 * moco.py:6-113, src/embeddings.py:44-57) flattened by the Python host into a list of fused ops over NHWC bf16
 * activation slots. Replaces EmbeddingNet._forward (src/embeddings.py:374-384) for those names.
 */
#define PVR_OP_CONV 1     /* implicit-GEMM conv (tcgen05) + folded-BN scale/bias + optional residual + optional ReLU */
#define PVR_OP_MAXPOOL 2  /* 3x3 stride-2 pad-1 max pool, NHWC bf16 */
#define PVR_OP_AVGPOOL 3  /* global average pool -> float32 rows of the embedding */
#define PVR_OP_HEAD 4     /* compression-head tail: 3x3 conv c->c + BN + identity add + ReLU -> NCHW-flatten f32.
                             act == 0: input = bf16 (2c) per pixel [relu(bn1(conv1 x)) | bn_d(conv_d x)];
                             act == 1: input = float32 per-tap partial sums, 9 x 2c per pixel (pitch in floats): the op
                             first forms sum_tap Z[p + tap - 1][tap], applies scale1/bias1 (+ ReLU on the first c) */
#define PVR_OP_AVGPOOL2 6 /* 2x2 stride-2 average pool NHWC -> NHWC (nn.AvgPool2d(2) of CLIP's ModifiedResNet); bf16 or,
                           * with PVR_OP_FP32 in flags, float32 */
#define PVR_OP_FLATTEN 5  /* NHWC bf16 slot -> NCHW-flattened float32 embedding columns (`out.view(-1, out_size)`) */

/* CONV flag: the output slot holds float32 values (out_pitch floats per pixel; the slot is sized as 2*out_pitch bf16
 * elements per pixel). 1x1 / stride-1 convs without residual, c_out % 32 == 0. Used by the compression heads, whose
 * 3x3 convolution over C = 1024 / 2048 channels runs as ONE 1x1 GEMM producing the 9 per-tap partial sums of every
 * pixel (the input is read once instead of nine times); the HEAD op adds the shifted partial sums in fp32. */
#define PVR_CONV_OUT_F32 1
/* Any op: fp32 parity mode (north star: embeddings within 1e-5 of the reference). Activations, weights, residuals are
 * float32 (pitches count floats, slots are sized as 2 bf16 elements per value), weights are dense (c_out, r, s, c_in)
 * float32 with k_pad = r*s*c_in; the op runs on the CUDA cores (conv_f32_kernel, fp32 FMA accumulation) instead of the
 * bf16 tensor-core kernels. No fusion flags (in2, OUT_F32) in this mode. */
#define PVR_OP_FP32 2

typedef struct pvr_op {
  int32_t kind;
  int32_t in_slot, out_slot, res_slot; /* activation slots; res_slot < 0: no residual; out_slot < 0: embedding */
  int32_t c_in, h_in, w_in;            /* logical input tensor per image, as the op's loader sees it */
  int32_t in_pitch;                    /* channel pitch (elements) of one input pixel, >= c_in */
  int32_t c_out, h_out, w_out;
  int32_t out_pitch, res_pitch;        /* channel pitch of output / residual pixels */
  int32_t out_coff, res_coff;          /* channel offset inside the output / residual pixel */
  int32_t r, s;                        /* filter taps (rows, cols) as seen by the loader */
  int32_t stride_h, stride_w;
  int32_t lower_h, lower_w;            /* coordinate of tap (0,0) for output pixel (0,0), i.e. -padding */
  int32_t relu_n;                      /* ReLU on output channels < relu_n (0: none; c_out: all) */
  int32_t block_n;                     /* N tile hint: 0 = auto, else 32/64/128/256 */
  int32_t k_pad;                       /* packed K extent of `weight` (multiple of 64) */
  int32_t n_pad;                       /* packed row count of `weight` (multiple of the N tile) */
  int32_t emb_offset;                  /* AVGPOOL/HEAD/FLATTEN: first column inside the embedding row */
  int32_t act;                         /* CONV: 0 = ReLU mask of relu_n only, 3 = ELU on every channel (small-conv PVR) */
  /* Optional second input of a 1x1 / stride-1 CONV (in2_c > 0): a 1x1 convolution with stride in2_stride over slot
   * in2_slot accumulated into the same output, K ordered [c_in | in2_c], k_pad = c_in + in2_c. This is how the
   * projection shortcut of torchvision's Bottleneck (resnet.py:143-166: out = bn3(conv3(t)) + bn_d(conv_d(x))) runs
   * as ONE GEMM: both BN scales are folded into the packed weight rows, bias = b3 + b_d, scale = 1. */
  int32_t in2_slot, in2_c, in2_h, in2_w, in2_pitch, in2_stride;
  int32_t flags;                       /* CONV: PVR_CONV_OUT_F32 */
  const void* weight;                  /* bf16 (n_pad, k_pad), K-major, K ordered (tap_row, tap_col, channel) */
  const float* scale;                  /* (n_pad) folded BN scale */
  const float* bias;                   /* (n_pad) folded BN bias (+ conv bias) */
  const void* aux;                     /* HEAD: float32 weights (c, 3, 3, c) + scale/bias, see pack in heads.py */
} pvr_op;

typedef struct pvr_slot {
  int64_t elems_per_image; /* bf16 elements per image */
} pvr_slot;

typedef struct pvr_encoder pvr_encoder;

/* Build a program. `ops`/`slots` are host arrays (copied). `emb_width` = floats per image in the output row. */
int pvr_encoder_create(const pvr_op* ops, int n_ops, const pvr_slot* slots, int n_slots, int emb_width,
                       pvr_encoder** out);
/* Bytes of workspace needed for a batch of `n_images`. */
int64_t pvr_encoder_workspace_bytes(const pvr_encoder* enc, int n_images);
/* Bind batch size + workspace (device, 1024-B aligned): encodes all TMA descriptors. Slot 0 is the input slot:
 * the caller (or pvr_preprocess_u8) writes NHWC bf16 frames there; its address is returned in *slot0. */
int pvr_encoder_bind(pvr_encoder* enc, int n_images, void* workspace, int64_t workspace_bytes, void** slot0);
/* Run all ops for the bound batch. Embedding rows are written to emb[i*emb_ld + ...] (float32, device). */
int pvr_encoder_forward(pvr_encoder* enc, float* emb, int64_t emb_ld, void* stream);
/* Same, with a CUDA event between ops; synchronises at the end and writes the device time of each op (ms) to the
 * host array op_ms[n_ops]. Used by bench.py for the per-kernel roofline numbers, not on the product path. */
int pvr_encoder_forward_timed(pvr_encoder* enc, float* emb, int64_t emb_ld, void* stream, float* op_ms);
/* Device address of a slot after bind (tests / per-layer parity). */
void* pvr_encoder_slot_ptr(const pvr_encoder* enc, int slot);
/* Number of kernel launches one forward issues (for bench.py's gpu_launches). */
int pvr_encoder_launch_count(const pvr_encoder* enc);
void pvr_encoder_destroy(pvr_encoder* enc);

/* ------------------------------------------------------------------------------------------------------------
 * Plain GEMM on the same tcgen05 core (used by 1x1 convs, the policy MLP/LSTM input projections, ViT later):
 *   out[m, n] = act( scale[n] * sum_k a[m,k] * b[n,k] + bias[n] (+ res[m,n]) ),  a,b bf16 K-major, out bf16.
 * m arbitrary, k multiple of 64, b has n_pad rows (multiple of 64).
 */
int pvr_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, const float* scale,
                  const float* bias, const void* res, int64_t ldr, int m, int n, int n_pad, int k, int relu,
                  void* stream);

/* General form used by the policy network (nn.Linear / nn.LSTM call sites of src/models.py:22-44 and their
 * backward): out = epilogue(a (m x k) * b^T (n x k)), a and b bf16 row-major (K contiguous). */
typedef struct pvr_gemm_desc {
  const void* a;      /* bf16 (m, k), row pitch lda */
  const void* b;      /* bf16 (n_pad, k), row pitch ldb; rows >= n must be zero */
  void* out;          /* bf16 or fp32 (m, n), row pitch ldo */
  const float* scale; /* (n_pad) per-column scale or NULL (= 1) */
  const float* bias;  /* (n_pad) per-column bias or NULL (= 0) */
  const void* res;    /* optional residual (m, n), row pitch ldr: bf16, or fp32 when out_f32 == 1 (may alias out) */
  int64_t lda, ldb, ldo, ldr;
  int32_t m, n, n_pad, k;
  int32_t relu;       /* ReLU on the result */
  int32_t res_mode;   /* 0: out += res; 1: out = res > 0 ? out : 0 (ReLU backward against the saved activation) */
  int32_t out_f32;    /* 0: bf16 out; 1: fp32 out; 2: fp32 out, atomically accumulated (out must hold the addend) */
  int32_t split_k;    /* out_f32 == 2 only: number of K slices computed by separate CTAs */
  int32_t act;        /* 0: none (or `relu`), 1: ReLU, 2: QuickGELU x*sigmoid(1.702x), 3: erf GELU (bf16 output) */
  int32_t flags;      /* PVR_GEMM_PDL: launch with programmatic stream serialization (the kernel's prologue overlaps
                         the previous kernel of the stream; it waits for that kernel before touching a / res / out) */
} pvr_gemm_desc;
#define PVR_GEMM_PDL 1
/* Operands given transposed: a is a row-major (k, m) matrix (row pitch lda), b a row-major (k, n_pad) matrix; the result
 * is still out (m, n) = a^T b. fp32 output (out_f32 == 1), no split-K, no residual; k is arbitrary (rows beyond it read
 * as zero). This is dW = dY^T X of nn.Linear with dY, X as the (rows, features) activations: no transposed copies. */
#define PVR_GEMM_MN 2
int pvr_gemm(const pvr_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * CLIP-architecture ViT pieces (src/embeddings.py:303-304,375-376 -> openai/CLIP VisionTransformer.forward). The
 * patch embedding, QKV / projection / MLP matmuls go through pvr_gemm (QuickGELU and the fp32 residual stream are
 * GEMM epilogues); width must be 768 (ViT-B), head_dim 64.
 */
/* y (rows, width) bf16 = LayerNorm(x[r * row_step]) * gamma + beta; x fp32 (row_step = tokens picks the class token).
 * width 768 (ViT-B) or 1024 (ViT-L, mae_large). */
int pvr_layernorm(const float* x, int64_t row_step, int64_t rows, int width, const float* gamma, const float* beta,
                  float eps, void* y_bf16, void* stream);
/* Same with a float32 result (dense rows, ldy == width): the final `norm` of the MAE encoder, whose class-token row IS
 * the embedding (src/vision_models/mae.py:222, src/embeddings.py:378-379). */
int pvr_layernorm_f32(const float* x, int64_t row_step, int64_t rows, int width, const float* gamma, const float* beta,
                      float eps, float* y, int64_t ldy, void* stream);
/* x_out (n_img*tokens, width) fp32 = ln_pre([class_embedding | patches] + positional_embedding); gamma == beta == NULL:
 * no LayerNorm (MAE: cls_token + pos_embed[0] | patches + pos_embed[1:], src/vision_models/mae.py:206-217). */
int pvr_vit_embed(const void* patches_bf16, const float* cls, const float* pos, int n_img, int tokens, int width,
                  const float* gamma, const float* beta, float eps, float* x_out, void* stream);
/* out (n_img*tokens, width) bf16 = softmax(q k^T / sqrt(head_dim)) v per image and head; qkv (n_img*tokens, 3*width) bf16
 * laid out [q | k | v] with heads contiguous inside each (nn.MultiheadAttention in_proj order). tcgen05 kernel. */
int pvr_attention(const void* qkv_bf16, int n_img, int tokens, int width, int heads, void* out_bf16, void* stream);
/* Same contract for any head_dim in {64, 80, 96, 128} and any sequence whose K / V fit shared memory (mae_huge,
 * src/vision_models/mae.py:291-296: 257 tokens, 16 heads of 80): warp-level mma.sync kernel with online softmax
 * (csrc/attention_mma.cu). pvr_attention forwards here for the shapes its tcgen05 kernel does not cover. */
/* Token assembly of CLIP's AttentionPool2d (openai/CLIP clip/model.py; `clip.load("RN50")`, src/embeddings.py:305-306):
 * tokens (n_img, hw + 1, width) with tokens[i][0] = mean_j x[i][j] + pos[0], tokens[i][j + 1] = x[i][j] + pos[j + 1];
 * x (n_img, hw, width) is the NHWC output of layer4. f32 = 0: bf16 in / out; f32 = 1: float32 in / out. */
int pvr_attnpool_tokens(const void* x, int n_img, int hw, int width, const float* pos, int f32, void* tokens,
                        void* stream);
int pvr_attention_mma(const void* qkv_bf16, int n_img, int tokens, int width, int heads, void* out_bf16, void* stream);
/* im2col of non-overlapping p x p patches (timm PatchEmbed, Conv2d(3, width, p, stride p), for patch sizes whose rows
 * are not a whole number of 128-byte TMA rows: mae_huge, p = 14): col (n_img * grid * grid, k_pad) with
 * col[(img, gy, gx)][c * p * p + py * p + px] = x[img][gy * p + py][gx * p + px][c], zero beyond 3 p^2. x is NHWC4;
 * f32 = 0: bf16 in / bf16 out, f32 = 1: float32 in / float32 out (fp32 parity mode). The patch embedding is then one
 * GEMM against the (width, 3 p^2) weight in its natural torch layout, zero padded to k_pad. */
int pvr_vit_patchify(const void* x_nhwc4, int n_img, int res, int patch, int k_pad, int f32, void* col, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BC policy network pieces (src/models.py:13-89 PolicyNet, main_bc_2.py:206-227 loss / clip / RMSprop).
 * All matrices are row-major; "bf16" pointers are void*.
 */
/* BatchNorm1d, train mode (src/models.py:30-34): phase 1 accumulates per-feature sum / sum of squares in double
 * (sums: 2*d doubles, zeroed inside) — all-reduce `sums` across ranks for data-parallel training — phase 2 turns them
 * into mean / rstd (biased variance, eps), updates the running statistics (momentum, unbiased variance over `count`
 * rows) and writes y = (x - mean) * rstd * gamma + beta as bf16. */
int pvr_bn1d_stats(const float* x, int64_t ldx, int m, int d, double* sums, void* stream);
int pvr_bn1d_normalize(const float* x, int64_t ldx, int m, int d, const double* sums, double count, float eps,
                       float momentum, const float* gamma, const float* beta, float* running_mean, float* running_var,
                       float* mean, float* rstd, void* y_bf16, int64_t ldy, void* stream);
/* eval mode: running statistics. */
int pvr_bn1d_eval(const float* x, int64_t ldx, int m, int d, const float* running_mean, const float* running_var,
                  float eps, const float* gamma, const float* beta, float* mean, float* rstd, void* y_bf16,
                  int64_t ldy, void* stream);
/* dgamma += sum dy * xhat, dbeta += sum dy (dy bf16 = gradient w.r.t. the BN output; outputs must be zeroed). */
int pvr_bn1d_backward(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, int m, int d, const float* mean,
                      const float* rstd, float* dgamma, float* dbeta, void* stream);
/* fp32 rows -> bf16 rows (policy input without BatchNorm; `x.float()` of src/models.py:60). */
int pvr_cast_rows_bf16(const float* x, int64_t ldx, int m, int d, void* y_bf16, int64_t ldy, void* stream);

/* One LSTM step (gate order i,f,g,o; state masked by notdone before the step, src/models.py:67-72). */
int pvr_lstm_cell_forward(const float* G, const float* XP, const float* c_prev, const float* nd, const float* nd_next,
                          int B, int H, float* gates, float* c_out, float* h_out_f32, void* h_out_bf16,
                          void* hm_next_bf16, void* stream);
int pvr_lstm_cell_backward(const float* dh_out, float* dh_rec, float* dc_rec, const float* gates, const float* c_prev,
                           const float* c_cur, const float* nd, const float* nd_next, int B, int H, void* dG_bf16,
                           void* stream);

/* All T steps of one LSTM layer (recurrent GEMM on tcgen05 + fused cell kernel per step).
 * `flags` (both structs) describe a CHUNK of a longer sequence whose arrays the pointers are offsets into, so that the
 * two layers can run as a wavefront on two streams (layer 1 works on chunk c while layer 0 works on chunk c + 1):
 *   PVR_LSTM_CONT_PREV: an earlier chunk precedes this one. Forward: hm[0] was written by that chunk's last step (h0 is
 *                       not read). Backward: the recurrent gradient is also propagated out of step 0 (into dh_rec).
 *   PVR_LSTM_CONT_NEXT: a later chunk follows. Forward: the last step also writes hm[T] and reads nd[T]. Backward: the
 *                       last step masks the incoming dh_rec by nd[T] (dh_rec / dc_rec hold the next chunk's result).
 * The chunked sequence launches exactly the kernels of the unchunked one. */
#define PVR_LSTM_CONT_PREV 1
#define PVR_LSTM_CONT_NEXT 2
typedef struct pvr_lstm_fwd {
  int32_t T, B, H, flags;
  const void* w_hh;   /* bf16 (4H, H) */
  const float* xp;    /* (T*B, 4H) fp32: x_t W_ih^T + b_ih + b_hh for every step (one big GEMM) */
  const float* nd;    /* (T, B) fp32 notdone = |1 - done| */
  const float* h0;    /* (B, H) fp32 initial hidden state; the initial cell state is c_all[0] */
  float* c_all;       /* ((T+1)*B, H) fp32: c_all[0] = c0 on entry, c_all[t+1] = c_t on exit (saved for backward) */
  void* hm;           /* bf16 (T*B, H) workspace: hm[t] = nd[t] * h[t-1], the recurrent GEMM operand */
  void* h_out;        /* bf16 (T*B, H): layer output */
  float* gates;       /* (T*B, 4H) fp32 activated gates, saved for backward */
  float* g_tmp;       /* (B, 4H) fp32 scratch */
  float* h_last;      /* (B, H) fp32: h_{T-1} */
  /* Optional arrival counters of the persistent kernel: a device buffer of at least PVR_LSTM_COUNTER_BYTES(T, B) that
   * belongs to this sequence (zeroed by the call). NULL: a process-wide buffer is used, which serialises nothing by
   * itself — callers that run recurrences of DIFFERENT sequences concurrently on several streams must pass their own. */
  void* counters;
  int64_t counters_bytes;
} pvr_lstm_fwd;
#define PVR_LSTM_COUNTER_BYTES(T, B) ((int64_t)((T) + 1) * (((B) + 31) / 32) * 16 * 4)
int pvr_lstm_forward(const pvr_lstm_fwd* layer, void* stream);
/* 1 if the whole-sequence persistent kernels (csrc/lstm_persist.cu: W_hh resident in tensor / shared memory, one launch for all
 * T steps) serve this shape on the current device: H = 1024, B <= 128, all CTAs co-resident. pvr_lstm_forward /
 * pvr_lstm_backward use them when `flags == 0`; otherwise (and with PVR_LSTM_PERSIST=0) the per-step kernels run. */
int pvr_lstm_persist_supported(int T, int B, int H);
/* Development aid: per-phase clock64 stamps of the persistent kernels' first 64 steps into `buf` (device memory,
 * 128 x 64 x 16 int64); NULL (the default) switches it off. */
int pvr_lstm_persist_profile(void* buf);

typedef struct pvr_lstm_bwd {
  int32_t T, B, H, flags;
  const void* w_hh_t;   /* bf16 (H, 4H): W_hh transposed */
  const float* nd;      /* (T, B) */
  const float* gates;   /* saved by the forward */
  const float* c_all;   /* saved by the forward */
  const float* dh_out;  /* (T*B, H) fp32 gradient w.r.t. the layer output, or NULL */
  float* dh_rec;        /* (B, H) fp32: gradient of the final hidden state on entry (zeros in BC) */
  float* dc_rec;        /* (B, H) fp32: gradient of the final cell state on entry (zeros in BC) */
  void* dG;             /* bf16 (T*B, 4H) out: gradient w.r.t. the gate pre-activations */
  float* dbias;         /* optional (4H) fp32: += sum over t, b of dG — the gradient of bias_ih (= that of bias_hh) */
  void* counters;       /* optional, as in pvr_lstm_fwd */
  int64_t counters_bytes;
} pvr_lstm_bwd;
int pvr_lstm_backward(const pvr_lstm_bwd* layer, void* stream);

/* Heads (src/models.py:75-76): logits = h Wp^T + bp (A <= 8 actions), baseline = h Wb^T + bb; h bf16 (m, K). */
int pvr_heads_forward(const void* h_bf16, int m, int K, const float* Wp, const float* bp, const float* Wb,
                      const float* bb, int A, float* logits, float* baseline, void* stream);
/* dh = scale * dlogits Wp (fp32), dWp += scale * dlogits^T h, dbp += scale * sum dlogits (outputs zeroed by caller). */
int pvr_heads_backward(const float* dlogits, const void* h_bf16, const float* Wp, int m, int K, int A, float scale,
                       float* dh, float* dWp, float* dbp, void* stream);
/* loss = inv_count * sum_m nll(log_softmax(logits[m]), targets[m]) (main_bc_2.py:211-214; inv_count = 1/(T*B) of
 * the GLOBAL batch), dlogits = inv_count * (softmax - onehot). `loss` is a device float. */
int pvr_ce_loss(const float* logits, const int64_t* targets, int m, int A, float inv_count, float* loss, float* dlogits,
                void* stream);

/* out[n] += sum_m y[m][n] (bias gradients), y bf16. */
int pvr_colsum_bf16(const void* y_bf16, int64_t ldy, int m, int n, float* out, void* stream);
/* out (cols x rows) = in (rows x cols)^T, bf16. */
int pvr_transpose_bf16(const void* in, int64_t ldi, int rows, int cols, void* out, int64_t ldo, void* stream);
/* fp32 weight (rows x cols) -> bf16 copy and/or bf16 transposed copy (either may be NULL). */
int pvr_cast_weight(const float* w, int rows, int cols, void* w_bf16, int64_t ldb, void* wt_bf16, int64_t ldt,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 parity mode of the ViT encoders (csrc/vit_f32.cu; CLIP `encode_image`, src/embeddings.py:375-376, and the MAE
 * encoders, :377-379): float32 end to end on the CUDA cores — the north star's "<= 1e-5 in the fp32 mode". LayerNorm
 * is pvr_layernorm_f32 above. Not a performance path.
 */
/* out (M x N) = act(a (M x K) w^T (N x K) + bias (+ res)), float32, w dense; act 0 none / 2 QuickGELU / 3 erf GELU;
 * `res` may alias `out`. Blocked summation (16-deep slices summed on their own). */
int pvr_gemm_f32(const float* a, int64_t lda, const float* w, const float* bias, const float* res, int64_t ldr,
                 float* out, int64_t ldo, int64_t M, int N, int K, int act, void* stream);
/* x = [cls | patches] + pos, LayerNorm'ed when gamma / beta are given (CLIP ln_pre); patches float32 (n*(tokens-1), W). */
int pvr_vit_embed_f32(const float* patches, const float* cls, const float* pos, int n_img, int tokens, int width,
                      const float* gamma, const float* beta, float eps, float* x_out, void* stream);
/* softmax(q k^T / 8) v per (image, head), head_dim 64; qkv float32 (n*tokens, 3*width) -> out (n*tokens, width). */
int pvr_attention_f32(const float* qkv, int n_img, int tokens, int width, int heads, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data-parallel collectives (K12). The reference is single process (no DDP / NCCL
 * anywhere); data parallelism is new functionality whose parity target is the reference at the GLOBAL batch. One
 * communicator per process / GPU over NCCL (NVLink 5 / NVSwitch inside a node), bound at run time to the libnccl.so.2
 * of the torch installation. Every call is a plain stream-ordered launch on `stream`: it can be captured into a CUDA
 * graph with the kernels around it and forked onto a second stream to overlap with compute. `dtype`: PVR_COMM_*.
 * Reductions are sums (the BC loss is pre-scaled by 1 / global rows, pvr_ce_loss). */
#define PVR_COMM_F32 0
#define PVR_COMM_F64 1
#define PVR_COMM_BF16 2
#define PVR_COMM_I64 3
/* Load NCCL (optional explicit path of libnccl.so.2; NULL = the copy already loaded in the process). */
int pvr_comm_load(const char* libnccl_path);
/* NCCL version code (e.g. 22809), 0 if NCCL is not loaded. */
int pvr_comm_version(void);
/* 128-byte unique id, created on one rank and handed to the others by the host (torch.distributed store / file). */
int pvr_comm_unique_id(void* id128);
/* Collective over all ranks: creates this rank's communicator on the CURRENT CUDA device. */
int pvr_comm_init(int rank, int world, const void* id128, void** comm_out);
int pvr_comm_destroy(void* comm);
/* In-place sum over ranks of `count` elements. */
int pvr_comm_allreduce(void* comm, void* buf, int64_t count, int dtype, void* stream);
/* recv (recv_count elements) = this rank's slice of the sum over ranks of send (world * recv_count elements). */
int pvr_comm_reduce_scatter(void* comm, const void* send, void* recv, int64_t recv_count, int dtype, void* stream);
/* recv (world * send_count elements) = concatenation over ranks of send (send_count elements). */
int pvr_comm_allgather(void* comm, const void* send, void* recv, int64_t send_count, int dtype, void* stream);
/* In-place broadcast from `root`. */
int pvr_comm_broadcast(void* comm, void* buf, int64_t count, int dtype, int root, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * End-to-end finetuning (PolicyNetWithConv, src/models.py:96-197; main_bc_finetune.py): layout kernels around the
 * GEMM formulation of the conv trunk's backward (3x3, stride 2, padding 1, 32 channels).
 */
/* feat (TB, C*w*h*N) fp32 in the reference's order `cat([conv(frame_f^T) for f], -1).view(TB, -1)`: index
 * c*(w*h*N) + x*(h*N) + f*h + y  <-  y_bf16[(tb*N+f)][y][x][c] (NHWC, `pitch` channels per pixel). */
int pvr_convfeat_gather(const void* y_bf16, int pitch, int TB, int N, int h, int w, int C, float* feat, void* stream);
/* inverse mapping for the gradient: dy (TB*N, h, w, C) fp32 <- dfeat (TB, ld). */
int pvr_convfeat_scatter(const float* dfeat, int64_t ld, int TB, int N, int h, int w, int C, float* dy, void* stream);
/* dz (M, 64) bf16 = dy (M, C) * ELU'(z), ELU' from the saved output y (1 if y > 0 else y + 1); columns >= C zero. */
int pvr_elu_backward(const float* dy, const void* y_bf16, int pitch, int64_t M, int C, void* dz_bf16, void* stream);
/* The same in one pass with what its consumers need: dz (M, 64) bf16, its transpose dzt (64, Mp) bf16 (K-major operand
 * of the weight-gradient GEMM; columns >= M are left untouched) and colsum[c] += sum_m dz[m][c] (bias gradient; may be
 * NULL). C a multiple of 16, Mp a multiple of 16. */
int pvr_elu_backward_fused(const float* dy, const void* y_bf16, int pitch, int64_t M, int C, void* dz_bf16,
                           void* dzt_bf16, int64_t Mp, float* colsum, void* stream);
/* colT ((9*Ci), Mp) bf16: transposed im2col of A (F, Hi, Wi, pitch) for a 3x3/s2/p1 conv, Ci = 4 or 32. */
int pvr_im2col_t(const void* a_bf16, int pitch, int F, int Hi, int Wi, int Ci, int Ho, int Wo, int64_t Mp,
                 void* colT_bf16, void* stream);
/* Weight / bias gradient of the FIRST layer of the small-conv trunk (src/models.py:108: Conv2d(3, 32, 3x3, stride 2,
 * padding 1) + ELU) in one pass, no im2col and no GEMM: dz = bf16(dy * ELU'(y));
 * dw[co][a][b][c] += sum_px dz[px][co] x[2 oy - 1 + a][2 ox - 1 + b][c] (fp32, (32, 3, 3, 4), zeroed by the caller; the
 * 4th channel is the padding of the NHWC4 frames and stays 0), dbias[co] += sum_px dz[px][co].
 * dy fp32 (F*Ho*Wo, 32); y bf16 with `y_pitch` elements per pixel; x bf16 (F, Hi, Wi, 4). */
int pvr_small_conv1_wgrad(const float* dy, const void* y_bf16, int y_pitch, const void* x_nhwc4_bf16, int F, int Hi,
                          int Wi, int Ho, int Wo, float* dw, float* dbias, void* stream);
/* dA (F, Hi, Wi, Ci) fp32 = col2im of dcol (F*Ho*Wo, Kp) bf16 (input gradient of the layer), Ci = 32. */
int pvr_col2im(const void* dcol_bf16, int Kp, int F, int Hi, int Wi, int Ci, int Ho, int Wo, float* dA, void* stream);
/* BatchNorm1d input gradient from dy (bf16), the saved statistics and the (all-reduced) sums of pvr_bn1d_backward. */
int pvr_bn1d_backward_dx(const void* dy_bf16, int64_t lddy, const float* x, int64_t ldx, int64_t M, int D,
                         const float* mean, const float* rstd, const float* gamma, const float* sum_dy_xhat,
                         const float* sum_dy, double count, float* dx, int64_t lddx, void* stream);
int pvr_bf16_rows_to_f32(const void* src_bf16, int64_t lds, int64_t M, int D, float* dst, int64_t ldd, void* stream);

/* Fused optimizer (main_bc_2.py:220-227): sumsq = sum of squared gradients over <= 32 tensors (device double, zeroed
 * inside); step = clip by the global norm sqrt(sumsq) (max_norm <= 0: no clipping; coefficient
 * min(1, max_norm / (norm + 1e-6)) like torch.nn.utils.clip_grad_norm_) and RMSprop (torch semantics: eps outside the
 * sqrt) or Adam update. The pre-clip norm is written to norm_out (device float) — the reference's gradient_norm stat. */
#define PVR_OPT_RMSPROP 0
#define PVR_OPT_ADAM 1
int pvr_optim_sumsq(const float* const* grads, const int64_t* sizes, int count, double* sumsq, void* stream);
int pvr_optim_step(int mode, float* const* params, float* const* grads, float* const* state1, float* const* state2,
                   const int64_t* sizes, int count, const double* sumsq, float grad_scale, float max_norm, float lr,
                   float alpha_or_beta1, float beta2, float eps, int step, float* norm_out, void* stream);
/* Same (RMSprop only) with the learning rate read from device memory, so that a whole training step — including the
 * LambdaLR schedule of main_bc_2.py:90 — can be replayed from a CUDA graph. */
int pvr_optim_step_dev(int mode, float* const* params, float* const* grads, float* const* state1, float* const* state2,
                       const int64_t* sizes, int count, const double* sumsq, float grad_scale, float max_norm,
                       const float* lr_dev, float alpha_or_beta1, float beta2, float eps, int step, float* norm_out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVR_B200_H */
