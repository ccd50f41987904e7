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
int pvr_preprocess_u8(const uint8_t* in, int N, int H, int W, int n_frames, int rh, int rw, int top, int left,
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
#define PVR_OP_HEAD 4     /* compression-head tail: 3x3 conv c->c + BN + identity add + ReLU -> NCHW-flatten f32 */

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
  int32_t emb_offset;                  /* AVGPOOL/HEAD: first column inside the embedding row */
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

#ifdef __cplusplus
}
#endif
#endif /* PVR_B200_H */
