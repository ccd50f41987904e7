// C ABI (include/pvr_b200.h): error reporting, encoder program executor, plain GEMM entry point.
#include "pvr_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "conv_gemm.cuh"
#include "kernels.cuh"

namespace {
thread_local char g_err[512] = "";
}

void pvr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pvr_last_error(void) { return g_err; }
extern "C" int pvr_abi_version(void) { return 3; }

namespace {

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}

int pick_block_n(int n_pad, long long m_tiles, int sms, int hint, bool has_res) {
  if (hint) return hint;
  if (n_pad % 64) return 32;
  // Largest tile that still gives every SM work; wide tiles amortise the A-operand traffic. Residual layers stay at
  // N <= 128: their epilogue needs 3 in-flight buffers per group (residual prefetch) and the smem for it.
  const int cand[3] = {256, 128, 64};
  for (int bn : cand)
    if (n_pad % bn == 0 && m_tiles * (n_pad / bn) >= sms && !(has_res && bn == 256)) return bn;
  return 64;
}

struct BoundConv {
  CUtensorMap ta, tb, to, tr, ta2, to2, tw1;
  pvr::ConvGemmParams p;
  pvr::Conv3x3PatchParams pp;
  pvr::ConvB2BParams bp;
  bool b2b;    // conv3 + residual fused with the next block's conv1 (conv_b2b.cu); the next op is marked fused
  int block_n, a_mode;
  bool epi_tma;
  bool patch;  // patch-resident 3x3 kernel (conv3x3_patch.cu)
  bool direct;  // first layer of the small-conv trunk: warp-level mma.sync kernel (small_conv.cu)
};

}  // namespace

struct pvr_encoder {
  std::vector<pvr_op> ops;
  std::vector<pvr_slot> slots;
  int emb_width = 0;
  int n_images = 0;
  int sms = 0;
  std::vector<char*> slot_ptr;
  std::vector<BoundConv> bound;  // one per op (unused for non-conv ops)
  std::vector<char> fused;       // op is executed inside the epilogue of the op before it (stem + max pool)
  // Small-batch (rollout) mode: the launch sequence of one forward into a fixed output buffer is captured into a CUDA
  // graph on its second use and replayed afterwards (50-100 launches of a few microseconds each are otherwise bound
  // by the host's launch cost). Invalidated by bind.
  cudaGraphExec_t graph = nullptr;
  float* graph_emb = nullptr;
  int64_t graph_ld = 0;
  int graph_seen = 0;
  ~pvr_encoder() {
    if (graph) cudaGraphExecDestroy(graph);
  }
};

extern "C" int pvr_encoder_create(const pvr_op* ops, int n_ops, const pvr_slot* slots, int n_slots, int emb_width,
                                  pvr_encoder** out) {
  if (!ops || n_ops <= 0 || !slots || n_slots <= 0 || !out || emb_width <= 0) {
    pvr_set_error("pvr_encoder_create: invalid argument");
    return PVR_ERR_ARG;
  }
  for (int i = 0; i < n_ops; ++i) {
    const pvr_op& o = ops[i];
    const bool emb_out = (o.kind == PVR_OP_AVGPOOL || o.kind == PVR_OP_HEAD || o.kind == PVR_OP_FLATTEN);
    if (o.in_slot < 0 || o.in_slot >= n_slots || (!emb_out && (o.out_slot < 0 || o.out_slot >= n_slots)) ||
        o.res_slot >= n_slots) {
      pvr_set_error("pvr_encoder_create: op %d references a slot out of range", i);
      return PVR_ERR_ARG;
    }
    if (o.kind == PVR_OP_CONV && (o.flags & PVR_OP_FP32)) {
      if (!o.weight || !o.scale || !o.bias || o.c_in % 4 || o.in_pitch % 4 || o.k_pad != o.r * o.s * o.c_in ||
          o.in2_c != 0 || (o.flags & PVR_CONV_OUT_F32)) {
        pvr_set_error("pvr_encoder_create: op %d has an invalid fp32 conv description", i);
        return PVR_ERR_ARG;
      }
    } else if (o.kind == PVR_OP_CONV) {
      if (!o.weight || !o.scale || !o.bias || o.k_pad <= 0 || o.k_pad % 64 || o.n_pad <= 0 || o.n_pad % 32 ||
          o.c_out > o.n_pad || (o.in_pitch % 8) || (o.out_pitch % 8) || (o.out_coff % 8)) {
        pvr_set_error("pvr_encoder_create: op %d has an invalid conv description", i);
        return PVR_ERR_ARG;
      }
      if (!(o.c_in == 8 || o.c_in == 32 || o.c_in % 64 == 0)) {
        pvr_set_error("pvr_encoder_create: op %d: c_in must be 8, 32 or a multiple of 64 (got %d)", i, o.c_in);
        return PVR_ERR_ARG;
      }
    } else if (o.kind == PVR_OP_AVGPOOL2) {
      if (o.h_in < 2 || o.w_in < 2 || o.h_out != o.h_in / 2 || o.w_out != o.w_in / 2 || o.in_pitch != o.c_in ||
          o.out_pitch != o.c_in || o.c_in % 8) {
        pvr_set_error("pvr_encoder_create: op %d has an invalid 2x2 average-pool description", i);
        return PVR_ERR_ARG;
      }
    } else if (o.kind != PVR_OP_MAXPOOL && o.kind != PVR_OP_AVGPOOL && o.kind != PVR_OP_HEAD &&
               o.kind != PVR_OP_FLATTEN) {
      pvr_set_error("pvr_encoder_create: op %d has unknown kind %d", i, o.kind);
      return PVR_ERR_ARG;
    }
  }
  pvr_encoder* e = new (std::nothrow) pvr_encoder();
  if (!e) {
    pvr_set_error("pvr_encoder_create: out of memory");
    return PVR_ERR_STATE;
  }
  e->ops.assign(ops, ops + n_ops);
  e->slots.assign(slots, slots + n_slots);
  e->emb_width = emb_width;
  *out = e;
  return PVR_OK;
}

static int64_t slot_offset(const pvr_encoder* enc, int slot, int n_images) {
  int64_t off = 0;
  for (int s = 0; s < slot; ++s) {
    int64_t b = enc->slots[s].elems_per_image * 2 * (int64_t)n_images;
    off += (b + 1023) & ~int64_t(1023);
  }
  return off;
}

extern "C" int64_t pvr_encoder_workspace_bytes(const pvr_encoder* enc, int n_images) {
  if (!enc || n_images <= 0) return PVR_ERR_ARG;
  return slot_offset(enc, (int)enc->slots.size(), n_images);
}

extern "C" int pvr_encoder_bind(pvr_encoder* enc, int n_images, void* workspace, int64_t workspace_bytes,
                                void** slot0) {
  if (!enc || n_images <= 0 || !workspace || (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    pvr_set_error("pvr_encoder_bind: invalid argument (workspace must be 1024-byte aligned)");
    return PVR_ERR_ARG;
  }
  if (workspace_bytes < pvr_encoder_workspace_bytes(enc, n_images)) {
    pvr_set_error("pvr_encoder_bind: workspace too small");
    return PVR_ERR_ARG;
  }
  enc->sms = device_sm_count();
  if (enc->sms <= 0) {
    pvr_set_error("pvr_encoder_bind: no CUDA device (%s)", cudaGetErrorString(cudaGetLastError()));
    return PVR_ERR_CUDA;
  }
  enc->n_images = 0;
  if (enc->graph) cudaGraphExecDestroy(enc->graph);
  enc->graph = nullptr;
  enc->graph_seen = 0;
  enc->slot_ptr.resize(enc->slots.size());
  for (size_t s = 0; s < enc->slots.size(); ++s)
    enc->slot_ptr[s] = static_cast<char*>(workspace) + slot_offset(enc, (int)s, n_images);
  enc->bound.assign(enc->ops.size(), BoundConv());
  enc->fused.assign(enc->ops.size(), 0);
  // Zig-zag tile order: a conv walks its tiles in the direction opposite to the one its input was written in, so it
  // starts on the rows that are still in L2. slot_rev[s] = 1 when slot s was last written last-to-first.
  std::vector<int> slot_rev(enc->slots.size(), 0);
  const bool zigzag = getenv("PVR_NO_ZIGZAG") == nullptr;
  const int pdl = getenv("PVR_NO_PDL") == nullptr;
  const bool b2b_on = getenv("PVR_NO_B2B") == nullptr;
  // layer2 variant with streamed weights (256 KB of weights per tile go through two-slot rings): bit-identical, and
  // kernel by kernel as fast as the two launches it replaces (133 us against 86 + 51 us per 256 frames) — but it keeps
  // the 512-channel tensor from being re-read by the next conv1 (0.8 MB per frame and block), and on the power-capped
  // B200s of this pool fewer HBM bytes are worth +1.2 % on the whole default step (A/B in one job, twice:
  // 45.02 / 45.29 k -> 45.68 / 45.77 k frames/s). On by default since the end of round 2; PVR_B2B_STREAM=0 turns it off.
  const bool b2b_stream_on = b2b_on && !(getenv("PVR_B2B_STREAM") && atoi(getenv("PVR_B2B_STREAM")) == 0);
  const int pair_mode = getenv("PVR_CTA2") ? atoi(getenv("PVR_CTA2")) : 2;  // 0 off, 1: 256-wide pair tiles; 2 (default since round 2: the N = 128 3x3 convs of layer2 gain 8 %: a single CTA reads A + W at the shared-memory port limit there): + 128-wide (no residual); 3: + residual layers
  for (size_t i = 0; i < enc->ops.size(); ++i) {
    const pvr_op& o = enc->ops[i];
    if (o.kind != PVR_OP_CONV) {
      if (!enc->fused[i] && o.out_slot >= 0 && o.out_slot < (int)slot_rev.size()) slot_rev[o.out_slot] = 0;
      continue;
    }
    if (enc->fused[i]) continue;  // executed inside the previous op's kernel (conv_b2b.cu)
    if (o.flags & PVR_OP_FP32) continue;  // fp32 parity mode: plain pointers, nothing to bind
    BoundConv& b = enc->bound[i];
    pvr::ConvGemmParams& p = b.p;
    memset(&p, 0, sizeof(p));
    const int reverse = zigzag ? !slot_rev[o.in_slot] : 0;
    slot_rev[o.out_slot] = reverse;
    p.reverse = reverse;
    p.pdl = pdl;
    b.pp.pdl = pdl;
    const long long M = (long long)n_images * o.h_out * o.w_out;
    b.patch = false;
    b.direct = false;
    if (o.r == 3 && o.s == 3 && o.stride_h == 1 && o.stride_w == 1 && o.lower_h == -1 && o.lower_w == -1 &&
        o.c_in == 64 && o.c_out == 64 && o.n_pad == 64 && o.in_pitch == 64 && o.out_pitch == 64 && o.out_coff == 0 &&
        o.h_in == o.h_out && o.w_in == o.w_out && o.w_out % 8 == 0 && o.res_slot < 0 && o.act == 0 &&
        (o.relu_n == 0 || o.relu_n >= 64) && o.k_pad == 576 && o.block_n == 0) {
      // layer1-style 3x3 convolution: keep the input patch and the weights resident in shared memory
      const char* err = "";
      b.pp.n_img = n_images; b.pp.P = o.h_out; b.pp.Q = o.w_out;
      b.pp.tiles_p = (o.h_out + 15) / 16; b.pp.tiles_q = o.w_out / 8;
      b.pp.relu = o.relu_n > 0; b.pp.scale = o.scale; b.pp.bias = o.bias; b.pp.stem = 0; b.pp.reverse = reverse;
      if (!pvr::make_tmap_4d(&b.ta, enc->slot_ptr[o.in_slot], 64, 64, o.w_in, o.h_in, n_images, 8, 18, &err) ||
          !pvr::make_tmap_2d(&b.tb, o.weight, 576, 64, 576, 64, &err) ||
          !pvr::make_tmap_4d(&b.to, enc->slot_ptr[o.out_slot], 64, 64, o.w_out, o.h_out, n_images, 8, 16, &err)) {
        pvr_set_error("pvr_encoder_bind: op %zu: patch conv tensor maps: %s", i, err);
        return PVR_ERR_CUDA;
      }
      b.patch = true;
      continue;
    }
    if (o.c_in == 32 && o.r == 7 && o.s == 1 && o.stride_h == 2 && o.stride_w == 1 && o.lower_h == -3 &&
        o.lower_w == 0 && o.c_out == 64 && o.n_pad == 64 && (o.in_pitch == 32 || o.in_pitch == 8) && o.out_pitch == 64 &&
        o.out_coff == 0 &&
        o.w_in == o.w_out && o.h_in == 2 * o.h_out && o.w_out % 8 == 0 && o.res_slot < 0 && o.act == 0 &&
        (o.relu_n == 0 || o.relu_n >= 64) && o.k_pad == 256 && o.block_n == 0) {
      // ResNet stem over the W-expanded input: patch-resident row taps (conv3x3_patch.cu, STEM)
      const char* err = "";
      b.pp.n_img = n_images; b.pp.P = o.h_out; b.pp.Q = o.w_out;
      b.pp.tiles_p = (o.h_out + 15) / 16; b.pp.tiles_q = o.w_out / 8;
      b.pp.relu = o.relu_n > 0; b.pp.scale = o.scale; b.pp.bias = o.bias; b.pp.stem = 1; b.pp.reverse = reverse;
      // The 3x3/s2 max pool that follows the stem is fused into its epilogue when nothing else reads the stem output.
      if (i + 1 < enc->ops.size() && b.pp.relu && !getenv("PVR_NO_POOL_FUSION")) {
        const pvr_op& mp = enc->ops[i + 1];
        bool sole_reader = mp.kind == PVR_OP_MAXPOOL && mp.in_slot == o.out_slot && mp.c_in == 64 &&
                           mp.h_in == o.h_out && mp.w_in == o.w_out && mp.h_out == (o.h_out - 1) / 2 + 1 &&
                           mp.w_out == (o.w_out - 1) / 2 + 1;
        for (size_t j = i + 2; sole_reader && j < enc->ops.size(); ++j) {  // until the slot is written again
          if (enc->ops[j].in_slot == o.out_slot || enc->ops[j].res_slot == o.out_slot) sole_reader = false;
          if (enc->ops[j].out_slot == o.out_slot) break;
        }
        if (sole_reader) {
          b.pp.pool_out = reinterpret_cast<__nv_bfloat16*>(enc->slot_ptr[mp.out_slot]);
          b.pp.pool_P = mp.h_out; b.pp.pool_Q = mp.w_out;
          b.pp.tiles_p = (mp.h_out + 6) / 7; b.pp.tiles_q = (mp.w_out + 2) / 3;
          enc->fused[i + 1] = 1;
          slot_rev[mp.out_slot] = reverse;
        }
      }
      // in_pitch 32: W-expanded input (PVR_FMT_STEM_BF16); in_pitch 8: compact padded rows, expanded by the tensor map
      const bool a_ok = o.in_pitch == 8
          ? pvr::make_tmap_stem_compact(&b.ta, enc->slot_ptr[o.in_slot], o.w_in, o.h_in, n_images, 8, 37, 2, &err)
          : pvr::make_tmap_4d(&b.ta, enc->slot_ptr[o.in_slot], 32, 32, o.w_in, o.h_in, n_images, 8, 37, &err, 2);
      if (!a_ok ||
          !pvr::make_tmap_2d_sw64(&b.tb, o.weight, 256, 64, 256, 64, &err) ||
          !pvr::make_tmap_4d(&b.to, enc->slot_ptr[o.out_slot], 64, 64, o.w_out, o.h_out, n_images, 8, 16, &err)) {
        pvr_set_error("pvr_encoder_bind: op %zu: stem patch tensor maps: %s", i, err);
        return PVR_ERR_CUDA;
      }
      b.patch = true;
      continue;
    }
    if (M > 0x7fffffffll) {
      pvr_set_error("pvr_encoder_bind: batch too large");
      return PVR_ERR_ARG;
    }
    b.b2b = false;
    if (b2b_on && i + 1 < enc->ops.size()) {
      // layer1 tail: relu(bn3(conv3 t2) + x), 64 -> 256, directly followed by the next block's conv1 (256 -> 64 / 128)
      const pvr_op& q = enc->ops[i + 1];
      auto pw = [](const pvr_op& c) {
        return c.kind == PVR_OP_CONV && c.r == 1 && c.s == 1 && c.stride_h == 1 && c.stride_w == 1 && c.lower_h == 0 &&
               c.lower_w == 0 && c.h_in == c.h_out && c.w_in == c.w_out && c.act == 0 && c.flags == 0 &&
               c.out_coff == 0 && c.block_n == 0;
      };
      // first GEMM: identity block (K = 64, + residual) or projection-shortcut block (K = [t2 | x] = 128, no residual)
      const bool ident = o.in2_c == 0 && o.k_pad == 64 && o.res_slot >= 0 && o.res_pitch == 256 && o.res_coff == 0;
      const bool proj = o.in2_c == 64 && o.in2_stride == 1 && o.in2_pitch == 64 && o.k_pad == 128 && o.res_slot < 0 &&
                        o.in2_slot >= 0 && o.in2_slot < (int)enc->slots.size();
      if (pw(o) && pw(q) && q.in2_c == 0 && (ident || proj) && o.c_in == 64 && o.in_pitch == 64 && o.c_out == 256 &&
          o.n_pad == 256 && o.out_pitch == 256 && o.relu_n >= 256 && q.in_slot == o.out_slot && q.c_in == 256 &&
          q.in_pitch == 256 && q.k_pad == 256 && q.res_slot < 0 && (q.c_out == 64 || (q.c_out == 128 && ident)) &&
          q.n_pad == q.c_out && q.out_pitch == q.c_out && q.relu_n >= q.c_out && q.h_in == o.h_out &&
          q.w_in == o.w_out && q.out_slot != o.out_slot && q.out_slot != o.res_slot && q.out_slot != o.in_slot &&
          !(proj && q.out_slot == o.in2_slot)) {
        const char* err = "";
        pvr::ConvB2BParams& bp = b.bp;
        bp.M = (int)M; bp.num_m_tiles = (int)((M + 127) / 128); bp.n2 = q.c_out; bp.reverse = reverse; bp.pdl = pdl;
        bp.k1_chunks = proj ? 2 : 1;
        bp.scale1 = o.scale; bp.bias1 = o.bias; bp.scale2 = q.scale; bp.bias2 = q.bias;
        b.ta2 = b.ta;
        b.tr = b.ta;
        if (!pvr::make_tmap_2d(&b.ta, enc->slot_ptr[o.in_slot], 64, (uint64_t)M, 64, 128, &err) ||
            (proj && !pvr::make_tmap_2d(&b.ta2, enc->slot_ptr[o.in2_slot], 64, (uint64_t)M, 64, 128, &err)) ||
            !pvr::make_tmap_2d(&b.tb, o.weight, (uint64_t)o.k_pad, 256, (uint64_t)o.k_pad, 256, &err) ||
            (ident && !pvr::make_tmap_2d(&b.tr, enc->slot_ptr[o.res_slot], 256, (uint64_t)M, 256, 128, &err)) ||
            !pvr::make_tmap_2d(&b.to, enc->slot_ptr[o.out_slot], 256, (uint64_t)M, 256, 128, &err) ||
            !pvr::make_tmap_2d(&b.tw1, q.weight, 256, (uint64_t)q.n_pad, 256, (uint32_t)q.n_pad, &err) ||
            !pvr::make_tmap_2d(&b.to2, enc->slot_ptr[q.out_slot], (uint64_t)q.c_out, (uint64_t)M,
                               (uint64_t)q.out_pitch, 128, &err)) {
          pvr_set_error("pvr_encoder_bind: op %zu: back-to-back tensor maps: %s", i, err);
          return PVR_ERR_CUDA;
        }
        b.b2b = true;
        enc->fused[i + 1] = 1;
        slot_rev[q.out_slot] = reverse;
        continue;
      }
      // layer2 identity blocks: 128 -> 512 (+ residual), then 512 -> 128; weights streamed (conv_b2b_stream_kernel)
      if (b2b_stream_on && pw(o) && pw(q) && o.in2_c == 0 && q.in2_c == 0 && o.c_in == 128 && o.in_pitch == 128 &&
          o.k_pad == 128 && o.c_out == 512 && o.n_pad == 512 && o.out_pitch == 512 && o.res_slot >= 0 &&
          o.res_pitch == 512 && o.res_coff == 0 && o.relu_n >= 512 && q.in_slot == o.out_slot && q.c_in == 512 &&
          q.in_pitch == 512 && q.k_pad == 512 && q.res_slot < 0 && q.c_out == 128 && q.n_pad == 128 &&
          q.out_pitch == 128 && q.relu_n >= 128 && q.h_in == o.h_out && q.w_in == o.w_out &&
          q.out_slot != o.out_slot && q.out_slot != o.res_slot && q.out_slot != o.in_slot) {
        const char* err = "";
        pvr::ConvB2BParams& bp = b.bp;
        bp.M = (int)M; bp.num_m_tiles = (int)((M + 127) / 128); bp.n2 = 128; bp.reverse = reverse; bp.pdl = pdl;
        bp.k1_chunks = 2; bp.streamed = 1;
        bp.scale1 = o.scale; bp.bias1 = o.bias; bp.scale2 = q.scale; bp.bias2 = q.bias;
        b.ta2 = b.ta;
        if (!pvr::make_tmap_2d(&b.ta, enc->slot_ptr[o.in_slot], 128, (uint64_t)M, 128, 128, &err) ||
            !pvr::make_tmap_2d(&b.tb, o.weight, 128, 512, 128, 256, &err) ||
            !pvr::make_tmap_2d(&b.tr, enc->slot_ptr[o.res_slot], 512, (uint64_t)M, 512, 128, &err) ||
            !pvr::make_tmap_2d(&b.to, enc->slot_ptr[o.out_slot], 512, (uint64_t)M, 512, 128, &err) ||
            !pvr::make_tmap_2d(&b.tw1, q.weight, 512, 128, 512, 128, &err) ||
            !pvr::make_tmap_2d(&b.to2, enc->slot_ptr[q.out_slot], 128, (uint64_t)M, 128, 128, &err)) {
          pvr_set_error("pvr_encoder_bind: op %zu: back-to-back (streamed) tensor maps: %s", i, err);
          return PVR_ERR_CUDA;
        }
        b.ta2 = b.ta;
        b.b2b = true;
        enc->fused[i + 1] = 1;
        slot_rev[q.out_slot] = reverse;
        continue;
      }
    }
    p.M = (int)M;
    p.P = o.h_out;
    p.Q = o.w_out;
    p.num_m_tiles = (int)((M + 127) / 128);
    b.block_n = pick_block_n(o.n_pad, p.num_m_tiles, enc->sms, o.block_n, o.res_slot >= 0);
    // CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles): wide layers whose tiles are bound by the L2 -> SM operand
    // traffic. Each CTA of a pair stages half of the W tile. Residual layers stay single-CTA at N = 128 (measured:
    // their in-place residual epilogue is slower at N = 256).
    if (pair_mode && o.block_n == 0 && o.n_pad % 128 == 0 && o.c_out % 64 == 0 && o.c_in % 64 == 0 && o.act != 3 &&
        M % 256 == 0 && o.k_pad >= 512 && !(o.flags & PVR_CONV_OUT_F32)) {
      const int pbn = (o.n_pad % 256 == 0 && o.res_slot < 0) ? 256 : 128;
      const bool allow = pbn == 256 ? true : (pair_mode >= 2 && (o.res_slot < 0 || pair_mode >= 3));
      if (allow && (M / 256) * (o.n_pad / pbn) >= enc->sms / 2) {
        p.cta2 = 1;
        p.num_m_tiles = (int)(M / 256);
        b.block_n = pbn;
      }
    }
    if (o.n_pad % b.block_n) {
      pvr_set_error("pvr_encoder_bind: op %zu: n_pad %d not a multiple of the N tile %d", i, o.n_pad, b.block_n);
      return PVR_ERR_ARG;
    }
    p.num_n_tiles = o.n_pad / b.block_n;
    p.num_k_chunks = o.k_pad / 64;
    p.split_k = 1;
    p.S = o.s;
    p.taps = o.r * o.s;
    p.stride_w = o.stride_w;
    p.stride_h = o.stride_h;
    p.lower_w = o.lower_w;
    p.lower_h = o.lower_h;
    p.n_valid = o.c_out;
    p.relu_n = o.relu_n;
    p.elu = o.act == 3;
    p.ldo = o.out_pitch;
    p.out = reinterpret_cast<__nv_bfloat16*>(enc->slot_ptr[o.out_slot]) + o.out_coff;
    if (o.res_slot >= 0) {
      p.res = reinterpret_cast<const __nv_bfloat16*>(enc->slot_ptr[o.res_slot]) + o.res_coff;
      p.ldr = o.res_pitch;
    }
    p.scale = o.scale;
    p.bias = o.bias;
    const char* err = "";
    const void* in = enc->slot_ptr[o.in_slot];
    const bool pointwise = (o.r == 1 && o.s == 1 && o.stride_h == 1 && o.stride_w == 1 && o.lower_h == 0 &&
                            o.lower_w == 0 && o.h_in == o.h_out && o.w_in == o.w_out);
    bool ok;
    if (o.in2_c > 0 && !(pointwise && o.c_in % 64 == 0 && o.in2_c % 64 == 0 && o.in2_stride >= 1 &&
                         o.in2_slot >= 0 && o.in2_slot < (int)enc->slots.size() &&
                         (o.in2_h - 1) / o.in2_stride + 1 == o.h_out && (o.in2_w - 1) / o.in2_stride + 1 == o.w_out)) {
      pvr_set_error("pvr_encoder_bind: op %zu: a second input needs a 1x1/stride-1 conv with 64-multiple channels "
                    "and matching output size", i);
      return PVR_ERR_ARG;
    }
    b.ta2 = b.ta;
    if (pointwise && o.c_in % 64 == 0) {
      b.a_mode = pvr::A_TILED;
      if (o.k_pad != o.c_in + o.in2_c) {
        pvr_set_error("pvr_encoder_bind: op %zu: k_pad must equal c_in (+ in2_c) for 1x1 convs", i);
        return PVR_ERR_ARG;
      }
      ok = pvr::make_tmap_2d(&b.ta, in, (uint64_t)o.c_in, (uint64_t)M, (uint64_t)o.in_pitch, 128, &err);
      if (ok && o.in2_c > 0) {
        const void* in2 = enc->slot_ptr[o.in2_slot];
        p.kc_split = o.c_in / 64;
        p.a2_stride = o.in2_stride;
        if (o.in2_stride == 1) {
          ok = pvr::make_tmap_2d(&b.ta2, in2, (uint64_t)o.in2_c, (uint64_t)M, (uint64_t)o.in2_pitch, 128, &err);
        } else {
          p.a2_im2col = 1;
          const int up_w = (o.w_out - 1) * o.in2_stride - (o.in2_w - 1);
          const int up_h = (o.h_out - 1) * o.in2_stride - (o.in2_h - 1);
          ok = pvr::make_tmap_im2col(&b.ta2, in2, o.in2_c, o.in2_pitch, o.in2_w, o.in2_h, n_images, 0, 0, up_w, up_h,
                                     o.in2_stride, o.in2_stride, 64, 128, 128, &err);
        }
      }
    } else {
      const int upper_w = o.lower_w + (o.w_out - 1) * o.stride_w - (o.w_in - 1);
      const int upper_h = o.lower_h + (o.h_out - 1) * o.stride_h - (o.h_in - 1);
      // 3 -> 32 channels, 3x3 stride 2 over NHWC4 pixel pairs + ELU (program.add_small_conv, first layer): K = 27 is
      // not a tensor-core problem (see small_conv.cu); PVR_SMALL_CONV_GEMM=1 keeps it on the implicit-GEMM kernel
      static const bool direct_ok = !(getenv("PVR_SMALL_CONV_GEMM") && atoi(getenv("PVR_SMALL_CONV_GEMM")) != 0);
      b.direct = direct_ok && o.c_in == 8 && o.in_pitch == 8 && o.r == 3 && o.s == 2 && o.stride_h == 2 &&
                 o.stride_w == 1 && o.lower_h == -1 && o.lower_w == -1 && o.k_pad == 64 && o.n_pad == 32 &&
                 o.c_out == 32 && o.out_pitch == 32 && o.out_coff == 0 && o.act == 3 && o.res_slot < 0 &&
                 o.relu_n == 0 && !(o.flags & PVR_CONV_OUT_F32) && o.h_out == (o.h_in - 1) / 2 + 1 &&
                 o.w_out == (2 * o.w_in - 1) / 2 + 1;
      if (o.c_in == 8) {
        b.a_mode = pvr::A_IM2COL8;
        if (o.k_pad < o.r * o.s * 8) {
          pvr_set_error("pvr_encoder_bind: op %zu: k_pad too small", i);
          return PVR_ERR_ARG;
        }
        ok = pvr::make_tmap_im2col(&b.ta, in, 8, o.in_pitch, o.w_in, o.h_in, n_images, o.lower_w, o.lower_h, upper_w,
                                   upper_h, o.stride_w, o.stride_h, 8, 128, 0, &err);
      } else if (o.c_in == 32) {
        b.a_mode = pvr::A_IM2COL32;
        if (o.k_pad < o.r * o.s * 32) {
          pvr_set_error("pvr_encoder_bind: op %zu: k_pad too small", i);
          return PVR_ERR_ARG;
        }
        ok = pvr::make_tmap_im2col(&b.ta, in, 32, o.in_pitch, o.w_in, o.h_in, n_images, o.lower_w, o.lower_h, upper_w,
                                   upper_h, o.stride_w, o.stride_h, 32, 128, 64, &err);
      } else {
        b.a_mode = pvr::A_IM2COL64;
        p.cin_chunks = o.c_in / 64;
        if (o.k_pad != o.r * o.s * o.c_in) {
          pvr_set_error("pvr_encoder_bind: op %zu: k_pad must equal r*s*c_in", i);
          return PVR_ERR_ARG;
        }
        ok = pvr::make_tmap_im2col(&b.ta, in, o.c_in, o.in_pitch, o.w_in, o.h_in, n_images, o.lower_w, o.lower_h,
                                   upper_w, upper_h, o.stride_w, o.stride_h, 64, 128, 128, &err);
      }
    }
    if (!ok) {
      pvr_set_error("pvr_encoder_bind: op %zu: activation tensor map: %s", i, err);
      return PVR_ERR_CUDA;
    }
    if (!pvr::make_tmap_2d(&b.tb, o.weight, (uint64_t)o.k_pad, (uint64_t)o.n_pad, (uint64_t)o.k_pad,
                           (uint32_t)(p.cta2 ? b.block_n / 2 : b.block_n), &err)) {
      pvr_set_error("pvr_encoder_bind: op %zu: weight tensor map: %s", i, err);
      return PVR_ERR_CUDA;
    }
    // Output / residual staged through shared memory + TMA when the channel count allows full 64-column sub-tiles.
    b.epi_tma = (b.block_n >= 64 && o.c_out % 64 == 0 && o.act != 3);
    b.to = b.ta;
    b.tr = b.ta;
    if (o.flags & PVR_CONV_OUT_F32) {
      if (!pointwise || o.res_slot >= 0 || o.c_out % 32 || b.block_n < 64 || o.out_coff != 0 || o.act != 0 ||
          o.out_pitch % 4) {
        pvr_set_error("pvr_encoder_bind: op %zu: float32 output needs a 1x1 conv, no residual, c_out %% 32 == 0", i);
        return PVR_ERR_ARG;
      }
      b.epi_tma = true;
      p.out_is_f32 = 1;
      p.out_f32 = reinterpret_cast<float*>(enc->slot_ptr[o.out_slot]);
      if (!pvr::make_tmap_2d_f32(&b.to, enc->slot_ptr[o.out_slot], (uint64_t)o.c_out, (uint64_t)M,
                                 (uint64_t)o.out_pitch, 128, &err)) {
        pvr_set_error("pvr_encoder_bind: op %zu: float32 output tensor map: %s", i, err);
        return PVR_ERR_CUDA;
      }
    } else if (b.epi_tma) {
      p.has_res = o.res_slot >= 0;
      p.out_coff = o.out_coff;
      p.res_coff = o.res_coff;
      if (!pvr::make_tmap_2d(&b.to, enc->slot_ptr[o.out_slot], (uint64_t)o.out_pitch, (uint64_t)M,
                             (uint64_t)o.out_pitch, 128, &err) ||
          (p.has_res && !pvr::make_tmap_2d(&b.tr, enc->slot_ptr[o.res_slot], (uint64_t)o.res_pitch, (uint64_t)M,
                                           (uint64_t)o.res_pitch, 128, &err))) {
        pvr_set_error("pvr_encoder_bind: op %zu: output/residual tensor map: %s", i, err);
        return PVR_ERR_CUDA;
      }
      // A-stationary tile order (ConvGemmParams::astat): plain 1x1 layers whose K chunks exactly fill the stage ring
      // and that have several N tiles per M tile — layer3's 256 -> 1024 + residual (K = 4 chunks, 8 N tiles), which
      // takes in 160 KB per 128 x 128 tile at ~32 B/clk and is bound by neither HBM nor the tensor pipe. M tiles are
      // padded to a multiple of the grid (<= 4 % more tiles; the padding tiles read zeros and store nothing).
      // PVR_ASTAT: 0 = off, 1 (default) = when the padding costs <= 4 %, 2 = always (tests: small batches)
      const int astat_mode = getenv("PVR_ASTAT") ? atoi(getenv("PVR_ASTAT")) : 1;
      if (astat_mode && b.a_mode == pvr::A_TILED && !p.cta2 && p.kc_split == 0 && p.split_k == 1 &&
          p.num_k_chunks == pvr::conv_gemm_stages(b.block_n, true) && p.num_n_tiles >= 2 &&
          (p.num_n_tiles & (p.num_n_tiles - 1)) == 0) {
        const int padded = (p.num_m_tiles + enc->sms - 1) / enc->sms * enc->sms;
        if (astat_mode >= 2 || (long long)padded * 100 <= (long long)p.num_m_tiles * 104) {
          p.num_m_tiles = padded;
          p.astat = 1;
        }
      }
    }
  }
  enc->n_images = n_images;
  if (slot0) *slot0 = enc->slot_ptr[0];
  return PVR_OK;
}

static int encoder_run(pvr_encoder* enc, float* emb, int64_t emb_ld, cudaStream_t stream, cudaEvent_t* ev) {
  const int n = enc->n_images;
  for (size_t i = 0; i < enc->ops.size(); ++i) {
    const pvr_op& o = enc->ops[i];
    cudaError_t e = cudaSuccess;
    if (ev) cudaEventRecord(ev[i], stream);
    if (enc->fused[i]) continue;
    if (o.flags & PVR_OP_FP32) {  // fp32 parity mode (conv_f32.cu)
      const float* in = reinterpret_cast<const float*>(enc->slot_ptr[o.in_slot]);
      switch (o.kind) {
        case PVR_OP_CONV: {
          pvr::ConvF32Params q;
          memset(&q, 0, sizeof(q));
          q.in = in;
          q.w = static_cast<const float*>(o.weight);
          q.scale = o.scale;
          q.bias = o.bias;
          q.res = o.res_slot >= 0 ? reinterpret_cast<const float*>(enc->slot_ptr[o.res_slot]) : nullptr;
          q.out = reinterpret_cast<float*>(enc->slot_ptr[o.out_slot]);
          q.M = (long long)n * o.h_out * o.w_out;
          q.N = o.c_out; q.C = o.c_in; q.H = o.h_in; q.W = o.w_in; q.P = o.h_out; q.Q = o.w_out; q.R = o.r; q.S = o.s;
          q.stride_h = o.stride_h; q.stride_w = o.stride_w; q.lower_h = o.lower_h; q.lower_w = o.lower_w;
          q.in_pitch = o.in_pitch; q.out_pitch = o.out_pitch; q.out_coff = o.out_coff; q.res_pitch = o.res_pitch;
          q.res_coff = o.res_coff; q.relu_n = o.relu_n; q.elu = o.act == 3;
          e = pvr::launch_conv_f32(q, stream);
          break;
        }
        case PVR_OP_MAXPOOL:
          e = pvr::launch_maxpool_f32(in, reinterpret_cast<float*>(enc->slot_ptr[o.out_slot]), n, o.h_in, o.w_in, o.c_in,
                                      o.h_out, o.w_out, stream);
          break;
        case PVR_OP_AVGPOOL:
          e = pvr::launch_avgpool_f32(in, emb, emb_ld, o.emb_offset, n, o.h_in * o.w_in, o.c_in, stream);
          break;
        case PVR_OP_AVGPOOL2:
          e = pvr::launch_avgpool2(in, enc->slot_ptr[o.out_slot], n, o.h_in, o.w_in, o.c_in, 1, stream);
          break;
        case PVR_OP_FLATTEN:
          e = pvr::launch_flatten_f32(in, o.in_pitch, emb, emb_ld, o.emb_offset, n, o.h_in * o.w_in, o.c_in, stream);
          break;
        case PVR_OP_HEAD:
          e = pvr::launch_head_tail_f32(in, o.in_pitch, static_cast<const float*>(o.aux), emb, emb_ld, o.emb_offset, n,
                                        o.h_in, o.w_in, o.c_out, stream);
          break;
      }
      if (e != cudaSuccess) {
        pvr_set_error("pvr_encoder_forward: fp32 op %zu (kind %d): %s", i, o.kind, cudaGetErrorString(e));
        return PVR_ERR_CUDA;
      }
      continue;
    }
    switch (o.kind) {
      case PVR_OP_CONV: {
        const BoundConv& b = enc->bound[i];
        e = b.direct ? pvr::launch_small_conv1(enc->slot_ptr[o.in_slot], n, o.h_in, 2 * o.w_in, o.h_out, o.w_out,
                                               o.weight, o.scale, o.bias, enc->slot_ptr[o.out_slot], stream)
            : b.b2b   ? pvr::launch_conv_b2b(b.ta, b.ta2, b.tb, b.tr, b.to, b.tw1, b.to2, b.bp, enc->sms, stream)
            : b.patch ? pvr::launch_conv3x3_patch(b.ta, b.tb, b.to, b.pp, enc->sms, stream)
                      : pvr::launch_conv_gemm(b.block_n, b.a_mode, b.epi_tma, b.ta, b.tb, b.to, b.tr, b.p, enc->sms,
                                            stream, &b.ta2);
        break;
      }
      case PVR_OP_MAXPOOL:
        e = pvr::launch_maxpool(reinterpret_cast<const __nv_bfloat16*>(enc->slot_ptr[o.in_slot]),
                                reinterpret_cast<__nv_bfloat16*>(enc->slot_ptr[o.out_slot]), n, o.h_in, o.w_in,
                                o.c_in, o.h_out, o.w_out, stream);
        break;
      case PVR_OP_AVGPOOL:
        e = pvr::launch_avgpool(reinterpret_cast<const __nv_bfloat16*>(enc->slot_ptr[o.in_slot]), emb, emb_ld,
                                o.emb_offset, n, o.h_in * o.w_in, o.c_in, stream);
        break;
      case PVR_OP_AVGPOOL2:
        e = pvr::launch_avgpool2(enc->slot_ptr[o.in_slot], enc->slot_ptr[o.out_slot], n, o.h_in, o.w_in, o.c_in, 0,
                                 stream);
        break;
      case PVR_OP_FLATTEN:
        e = pvr::launch_flatten(reinterpret_cast<const __nv_bfloat16*>(enc->slot_ptr[o.in_slot]), o.in_pitch, emb,
                                emb_ld, o.emb_offset, n, o.h_in * o.w_in, o.c_in, stream);
        break;
      case PVR_OP_HEAD:
        if (o.act == 1) {
          e = pvr::launch_head_tail_taps(reinterpret_cast<const float*>(enc->slot_ptr[o.in_slot]), o.in_pitch,
                                         static_cast<const float*>(o.aux), emb, emb_ld, o.emb_offset, n, o.h_in,
                                         o.w_in, o.c_out, stream);
          break;
        }
        e = pvr::launch_head_tail(reinterpret_cast<const __nv_bfloat16*>(enc->slot_ptr[o.in_slot]), o.in_pitch,
                                  static_cast<const float*>(o.aux), emb, emb_ld, o.emb_offset, n, o.h_in, o.w_in,
                                  o.c_out, stream);
        break;
    }
    if (e != cudaSuccess) {
      pvr_set_error("pvr_encoder_forward: op %zu (kind %d): %s", i, o.kind, cudaGetErrorString(e));
      return PVR_ERR_CUDA;
    }
  }
  if (ev) cudaEventRecord(ev[enc->ops.size()], stream);
  return PVR_OK;
}

static int encoder_check(pvr_encoder* enc, float* emb, int64_t emb_ld) {
  if (!enc || enc->n_images <= 0) {
    pvr_set_error("pvr_encoder_forward: encoder is not bound");
    return PVR_ERR_STATE;
  }
  if (!emb || emb_ld < enc->emb_width) {
    pvr_set_error("pvr_encoder_forward: invalid embedding buffer");
    return PVR_ERR_ARG;
  }
  return PVR_OK;
}

extern "C" int pvr_encoder_forward(pvr_encoder* enc, float* emb, int64_t emb_ld, void* stream_) {
  int rc = encoder_check(enc, emb, emb_ld);
  if (rc != PVR_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  static const bool graphs_on = getenv("PVR_NO_ENC_GRAPH") == nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (!graphs_on || enc->n_images > 8 || cudaStreamIsCapturing(st, &cs) != cudaSuccess ||
      cs != cudaStreamCaptureStatusNone)
    return encoder_run(enc, emb, emb_ld, st, nullptr);
  if (enc->graph && enc->graph_emb == emb && enc->graph_ld == emb_ld) {
    cudaError_t e = cudaGraphLaunch(enc->graph, st);
    if (e != cudaSuccess) {
      pvr_set_error("pvr_encoder_forward: cudaGraphLaunch: %s", cudaGetErrorString(e));
      return PVR_ERR_CUDA;
    }
    return PVR_OK;
  }
  if (enc->graph_emb != emb || enc->graph_ld != emb_ld) {  // new output buffer: start over
    if (enc->graph) cudaGraphExecDestroy(enc->graph);
    enc->graph = nullptr;
    enc->graph_emb = emb;
    enc->graph_ld = emb_ld;
    enc->graph_seen = 0;
  }
  if (enc->graph_seen++ == 0) return encoder_run(enc, emb, emb_ld, st, nullptr);  // first use stays eager
  // capture on a private stream (the caller's may be the legacy default stream), replay on the caller's
  static cudaStream_t cap = nullptr;
  if (!cap && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    return encoder_run(enc, emb, emb_ld, st, nullptr);
  }
  cudaGraph_t graph = nullptr;
  if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return encoder_run(enc, emb, emb_ld, st, nullptr);
  }
  rc = encoder_run(enc, emb, emb_ld, cap, nullptr);
  cudaError_t e = cudaStreamEndCapture(cap, &graph);
  if (rc != PVR_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    if (rc != PVR_OK) return rc;
    return encoder_run(enc, emb, emb_ld, st, nullptr);
  }
  e = cudaGraphInstantiate(&enc->graph, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) {
    enc->graph = nullptr;
    cudaGetLastError();
    return encoder_run(enc, emb, emb_ld, st, nullptr);
  }
  e = cudaGraphLaunch(enc->graph, st);
  if (e != cudaSuccess) {
    pvr_set_error("pvr_encoder_forward: cudaGraphLaunch: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}

extern "C" int pvr_encoder_forward_timed(pvr_encoder* enc, float* emb, int64_t emb_ld, void* stream_, float* op_ms) {
  int rc = encoder_check(enc, emb, emb_ld);
  if (rc != PVR_OK) return rc;
  if (!op_ms) {
    pvr_set_error("pvr_encoder_forward_timed: op_ms is null");
    return PVR_ERR_ARG;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t n = enc->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) cudaEventCreate(&e);
  rc = encoder_run(enc, emb, emb_ld, stream, ev.data());
  if (rc == PVR_OK) {
    cudaError_t e = cudaEventSynchronize(ev[n]);
    if (e != cudaSuccess) {
      pvr_set_error("pvr_encoder_forward_timed: %s", cudaGetErrorString(e));
      rc = PVR_ERR_CUDA;
    } else {
      for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&op_ms[i], ev[i], ev[i + 1]);
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

extern "C" void* pvr_encoder_slot_ptr(const pvr_encoder* enc, int slot) {
  if (!enc || enc->n_images <= 0 || slot < 0 || slot >= (int)enc->slot_ptr.size()) return nullptr;
  return enc->slot_ptr[slot];
}

extern "C" int pvr_encoder_launch_count(const pvr_encoder* enc) {
  if (!enc) return 0;
  int n = (int)enc->ops.size();
  for (char f : enc->fused) n -= f ? 1 : 0;
  return n;
}

extern "C" void pvr_encoder_destroy(pvr_encoder* enc) { delete enc; }

namespace {
// Device constants for GEMMs called without scale / bias (ones / zeros), allocated once per process and device.
const float* unit_vector(bool ones) {
  static float* bufs[16][2] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!bufs[dev][0]) {
    const size_t n = 16384;
    float* z = nullptr;
    float* o = nullptr;
    if (cudaMalloc(&z, n * sizeof(float)) != cudaSuccess || cudaMalloc(&o, n * sizeof(float)) != cudaSuccess)
      return nullptr;
    std::vector<float> h(n, 1.0f);
    cudaMemset(z, 0, n * sizeof(float));
    cudaMemcpy(o, h.data(), n * sizeof(float), cudaMemcpyHostToDevice);
    bufs[dev][0] = z;
    bufs[dev][1] = o;
  }
  return bufs[dev][ones ? 1 : 0];
}
}  // namespace

extern "C" int pvr_gemm(const pvr_gemm_desc* d, void* stream) {
  const bool mn = d && (d->flags & PVR_GEMM_MN);
  if (mn && (d->out_f32 != 1 || d->res || d->split_k > 1 || d->n_pad % 64 || d->n % 32)) {
    pvr_set_error("pvr_gemm: PVR_GEMM_MN needs fp32 output, no residual, no split-K, n_pad %% 64 == 0");
    return PVR_ERR_ARG;
  }
  if (!d || !d->a || !d->b || !d->out || d->m <= 0 || d->n <= 0 || d->n > d->n_pad || d->n_pad % 32 || d->k <= 0 ||
      (!mn && d->k % 64) || d->lda % 8 || d->ldb % 8 || (d->res && !d->out_f32 && d->ldr % 8) || d->n_pad > 16384 || d->out_f32 < 0 ||
      d->out_f32 > 2 || (d->out_f32 ? d->ldo % 4 : d->ldo % 8)) {
    pvr_set_error("pvr_gemm: invalid argument");
    return PVR_ERR_ARG;
  }
  const int split_k = d->split_k > 1 ? d->split_k : 1;
  if ((split_k > 1 && d->out_f32 != 2) || (d->k / 64) % split_k || (d->out_f32 == 2 && d->res) ||
      (d->out_f32 == 1 && d->res && (d->ldr % 4 || d->res_mode != 0))) {
    pvr_set_error("pvr_gemm: split_k needs out_f32 == 2 and must divide k/64; fp32 residuals only with out_f32 == 1");
    return PVR_ERR_ARG;
  }
  const int sms = device_sm_count();
  const float* scale = d->scale ? d->scale : unit_vector(true);
  const float* bias = d->bias ? d->bias : unit_vector(false);
  if (sms <= 0 || !scale || !bias) {
    pvr_set_error("pvr_gemm: no CUDA device");
    return PVR_ERR_CUDA;
  }
  pvr::ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->m;
  p.pdl = (d->flags & PVR_GEMM_PDL) != 0;
  p.num_m_tiles = (d->m + 127) / 128;
  // (ResNet's residual layers are HBM-bound and stay at N <= 128; a long-K GEMM with a mask / residual operand — the
  // policy's dZ = (dG W^T) * (H > 0), K = 4096 — is tensor-bound and takes the 256-wide tile)
  static const int res256 = getenv("PVR_RES_N256") ? atoi(getenv("PVR_RES_N256")) : 1;
  const bool narrow_res = d->res != nullptr && !(res256 && d->k >= 2048 && d->out_f32 == 0);
  int block_n = pick_block_n(d->n_pad, (long long)p.num_m_tiles * split_k, sms, 0, narrow_res);
  if (d->out_f32 == 2 && block_n > 128) block_n = 128;
  // Weight-gradient GEMMs (MN-major operands, e.g. 4096 x 1024 outputs over K = T*B): 256-wide tiles even when they
  // leave a few SMs idle (128 tiles on 148 SMs) — at N = 128 a single CTA reads A + W at the shared-memory port limit
  // (128 B/clk), at N = 256 it does not.
  if (mn && block_n == 128 && d->n_pad % 256 == 0 && (long long)p.num_m_tiles * (d->n_pad / 256) * 5 >= (long long)sms * 4 &&
      !getenv("PVR_WGRAD_N128"))
    block_n = 256;
  // ... and 128-wide ones rather than 64 when they fill >= 80 % of the SMs in one wave (dW1: 8 x 16 tiles): a 64-wide
  // tile is bound by the shared-memory port at 2/3 of the tensor rate
  if (mn && block_n == 64 && d->n_pad % 128 == 0 && (long long)p.num_m_tiles * (d->n_pad / 128) * 5 >= (long long)sms * 4 &&
      !getenv("PVR_WGRAD_N128"))
    block_n = 128;
  // CTA pairs (cta_group::2, 256 x 256 tiles) for wide bf16-output GEMMs with enough K: ViT QKV / fc1
  static const bool pair_ok = !(getenv("PVR_CTA2") && atoi(getenv("PVR_CTA2")) == 0);
  static const int pair_f32 = getenv("PVR_PAIR_F32") ? atoi(getenv("PVR_PAIR_F32")) : 1;
  const bool pair_bf16 = d->out_f32 == 0 && !d->res && d->n % 64 == 0 && d->k >= 512;
  const bool pair_fp32 = pair_f32 && d->out_f32 == 1 && d->n % 32 == 0 && d->k >= 768;  // ViT proj / fc2 (+ residual)
  if (pair_ok && !mn && (pair_bf16 || pair_fp32) && split_k == 1 && d->n_pad % 256 == 0 &&
      ((long long)(d->m + 255) / 256) * (d->n_pad / 256) >= sms / 2) {
    p.cta2 = 1;
    p.num_m_tiles = (d->m + 255) / 256;
    block_n = 256;
  }
  p.num_n_tiles = d->n_pad / block_n;
  p.split_k = split_k;
  p.num_k_chunks = mn ? (d->k + 63) / 64 : d->k / 64 / split_k;
  p.mn = mn ? 1 : 0;
  p.n_valid = d->n;
  p.relu_n = (d->relu || d->act == 1) ? d->n : 0;
  p.ldo = d->ldo;
  p.ldr = d->ldr;
  p.res_mode = d->res_mode;
  p.quick_gelu = d->act == 2 ? 1 : (d->act == 3 ? 2 : 0);
  p.out_is_f32 = d->out_f32 != 0;
  p.out = static_cast<__nv_bfloat16*>(d->out);
  p.out_f32 = static_cast<float*>(d->out);
  p.res = static_cast<const __nv_bfloat16*>(d->res);
  p.scale = scale;
  p.bias = bias;
  CUtensorMap ta, tb, to, tr;
  const char* err = "";
  const bool maps_ok =
      mn ? (pvr::make_tmap_2d(&ta, d->a, (uint64_t)d->m, (uint64_t)d->k, (uint64_t)d->lda, 64, &err) &&
            pvr::make_tmap_2d(&tb, d->b, (uint64_t)d->n_pad, (uint64_t)d->k, (uint64_t)d->ldb, 64, &err))
         : (pvr::make_tmap_2d(&ta, d->a, (uint64_t)d->k, (uint64_t)d->m, (uint64_t)d->lda, 128, &err) &&
            pvr::make_tmap_2d(&tb, d->b, (uint64_t)d->k, (uint64_t)d->n_pad, (uint64_t)d->ldb,
                              (uint32_t)(p.cta2 ? block_n / 2 : block_n), &err));
  if (!maps_ok) {
    pvr_set_error("pvr_gemm: %s", err);
    return PVR_ERR_CUDA;
  }
  bool epi_tma;
  to = ta;
  tr = ta;
  if (d->out_f32 == 0) {
    epi_tma = (block_n >= 64 && d->n % 64 == 0);
    if (epi_tma) {
      p.has_res = d->res != nullptr;
      if (!pvr::make_tmap_2d(&to, d->out, (uint64_t)d->n, (uint64_t)d->m, (uint64_t)d->ldo, 128, &err) ||
          (d->res && !pvr::make_tmap_2d(&tr, d->res, (uint64_t)d->n, (uint64_t)d->m, (uint64_t)d->ldr, 128, &err))) {
        pvr_set_error("pvr_gemm: %s", err);
        return PVR_ERR_CUDA;
      }
    }
  } else if (d->out_f32 == 1) {
    epi_tma = true;
    if (block_n < 64 || d->n % 32) {
      pvr_set_error("pvr_gemm: fp32 output needs n %% 32 == 0 and n_pad %% 64 == 0");
      return PVR_ERR_ARG;
    }
    p.has_res = d->res != nullptr;  // fp32 residual (same layout as the output; in place allowed)
    if (!pvr::make_tmap_2d_f32(&to, d->out, (uint64_t)d->n, (uint64_t)d->m, (uint64_t)d->ldo, 128, &err) ||
        (d->res && !pvr::make_tmap_2d_f32(&tr, d->res, (uint64_t)d->n, (uint64_t)d->m, (uint64_t)d->ldr, 128, &err))) {
      pvr_set_error("pvr_gemm: %s", err);
      return PVR_ERR_CUDA;
    }
  } else {
    epi_tma = false;
  }
  cudaError_t e = pvr::launch_conv_gemm(block_n, pvr::A_TILED, epi_tma, ta, tb, to, tr, p, sms,
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    pvr_set_error("pvr_gemm: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}

extern "C" int pvr_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo,
                             const float* scale, const float* bias, const void* res, int64_t ldr, int m, int n,
                             int n_pad, int k, int relu, void* stream) {
  pvr_gemm_desc d;
  memset(&d, 0, sizeof(d));
  d.a = a; d.lda = lda; d.b = b; d.ldb = ldb; d.out = out; d.ldo = ldo;
  d.scale = scale; d.bias = bias; d.res = res; d.ldr = ldr;
  d.m = m; d.n = n; d.n_pad = n_pad; d.k = k; d.relu = relu; d.split_k = 1;
  if (!scale || !bias) {
    pvr_set_error("pvr_gemm_bf16: invalid argument");
    return PVR_ERR_ARG;
  }
  int rc = pvr_gemm(&d, stream);
  if (rc == PVR_ERR_ARG) pvr_set_error("pvr_gemm_bf16: invalid argument");
  return rc;
}
