// Back-to-back fusion of a layer1 Bottleneck tail with the head of the next block (torchvision resnet.py:143-166):
//
//   out = relu(bn3(conv3(t2)) + x)            1x1, 64 -> 256, residual     (written: the next block's identity)
//         (or, K1C = 2: relu(W'[t2 | x] + b), the projection-shortcut block as one GEMM, no residual)
//   t1' = relu(bn1'(conv1'(out)))             1x1, 256 -> N2 (64 or 128)   (written: input of the next 3x3)
//
// As two kernels the 256-channel tensor `out` (411 MB per 256 frames at 56x56) is written once and read twice (next
// conv1, next residual); both kernels are HBM bound. Here the staged bf16 output sub-tiles of the first GEMM — already
// in the 128-byte-swizzled K-major layout the TMA store wants — are at the same time the four K chunks of the A
// operand of the second GEMM, so `out` is never re-read for conv1'.
//
// One tile = 128 pixels. Both weight matrices stay resident in shared memory (W3: 256 x 64, W1': N2 x 256).
//   warp 0      TMA producer of the A1 tiles (128 x 64 of t2)                    a_full / a_empty
//   warp 1      tcgen05.mma #1: acc1[128 x 256] = A1 * W3^T (4 instructions)      acc1_full / acc1_empty
//   warps 2..9  two epilogue groups. Sub-tile (128 x 64) of acc1: + residual (prefetched by TMA into the staging
//               buffer), BN, ReLU, bf16, in place -> eb_ready. Every other tile a group also runs epilogue #2:
//               acc2[128 x N2] -> BN, ReLU, bf16 -> staging -> TMA store of t1'.
//   warp 10     manager: residual prefetch (eb_full), and per finished sub-tile the TMA store to `out` plus
//               tcgen05.mma #2: acc2 += subtile * W1'[:, chunk]^T (reads the same staging buffer); a buffer is
//               refilled once its store has drained (bulk groups) and its MMA has completed (eb_mma_done).
// TMEM: acc1 = columns [0, 256) (single stage: its 4 MMAs take 0.3 us of a ~2 us HBM-bound tile), acc2 = two stages
// of N2 columns from 256.
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace pvr {
namespace {

constexpr int B2B_THREADS = 352;

// K1C = K chunks (of 64) of the first GEMM: 1 = identity block (conv3 over t2, + residual x), 2 = projection-shortcut
// block (K = [t2 | x], both BN scales folded into the weights, no residual; program._bottleneck).
template <int N2, int K1C>
struct B2BCfg {
  static constexpr int A_STAGES = N2 == 64 ? 3 : 2;     // ring of 128 x 64 A chunks
  static constexpr int NB = K1C == 2 ? 3 : (N2 == 64 ? 5 : 4);  // staging buffers for the 128 x 64 sub-tiles of `out`
  static constexpr uint32_t W3_CHUNK = 256 * 128;        // 256 output channels x 64 K
  static constexpr uint32_t W3_BYTES = K1C * W3_CHUNK;
  static constexpr uint32_t W1_CHUNK = N2 * 128;         // N2 output channels x 64 K
  static constexpr uint32_t W1_BYTES = 4 * W1_CHUNK;
  static constexpr uint32_t A_BYTES = 16384, EB_BYTES = 16384;
  static constexpr uint32_t SB_BYTES = (512 + 2 * N2) * 4;
  static constexpr uint32_t SMEM = 1024 + W3_BYTES + W1_BYTES + A_STAGES * A_BYTES + NB * EB_BYTES + EB_BYTES +
                                   ((SB_BYTES + 1023) / 1024) * 1024 + 512;
  static_assert(SMEM <= 232448, "exceeds the 227 KiB shared memory of one CTA");
};

__device__ __forceinline__ uint32_t b2b_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t b2b_relu2(uint32_t v) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&v);
  x = __hmax2(x, __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&x);
}

template <int N2, int K1C, bool RES>
__global__ void __launch_bounds__(B2B_THREADS, 1)
conv_b2b_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                const __grid_constant__ CUtensorMap tmap_w3,
                const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_out2,
                const ConvB2BParams p) {
  using C = B2BCfg<N2, K1C>;
  constexpr int A_STAGES = C::A_STAGES, NB = C::NB;
  constexpr int D = NB - 2;  // residual prefetch runs D sub-tiles ahead of the stores
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW3 = smem;
  uint8_t* sW1 = sW3 + C::W3_BYTES;
  uint8_t* sA = sW1 + C::W1_BYTES;
  uint8_t* sEB = sA + A_STAGES * C::A_BYTES;
  uint8_t* sS2 = sEB + NB * C::EB_BYTES;
  float* sSB = reinterpret_cast<float*>(sS2 + C::EB_BYTES);  // scale1[256] | bias1[256] | scale2[N2] | bias2[N2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sSB) + ((C::SB_BYTES + 1023) / 1024) * 1024);
  uint64_t* w_bar = bars;
  uint64_t* a_full = w_bar + 1;
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* acc1_full = a_empty + A_STAGES;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* acc2_full = acc1_empty + 1;    // [2]
  uint64_t* acc2_empty = acc2_full + 2;    // [2]
  uint64_t* eb_full = acc2_empty + 2;      // [NB]
  uint64_t* eb_ready = eb_full + NB;       // [NB]
  uint64_t* eb_mma_done = eb_ready + NB;   // [NB]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(eb_mma_done + NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles;
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int n) {  // n-th tile of this CTA (zig-zag order over consecutive layers)
    const int t = (int)blockIdx.x + n * (int)gridDim.x;
    return p.reverse ? num_tiles - 1 - t : t;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    if (K1C == 2) prefetch_tmap(&tmap_a2);
    prefetch_tmap(&tmap_w3);
    if (RES) prefetch_tmap(&tmap_res);
    prefetch_tmap(&tmap_out);
    prefetch_tmap(&tmap_w1);
    prefetch_tmap(&tmap_out2);
    mbar_init(w_bar, 1);
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 256);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], 128);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&eb_full[s], 1);
      mbar_init(&eb_ready[s], 1);
      mbar_init(&eb_mma_done[s], 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 512 + 2 * N2; i += blockDim.x)
    sSB[i] = i < 256 ? p.scale1[i] : i < 512 ? p.bias1[i - 256] : i < 512 + N2 ? p.scale2[i - 512] : p.bias2[i - 512 - N2];
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  griddep_launch();  // programmatic dependent launch, see conv_gemm.cu
  griddep_wait();

  if (warp == 0) {
    // ===================================================== producer: weights once, then one A1 tile per tile
    if (elect_one()) {
      mbar_expect_tx(w_bar, C::W3_BYTES + C::W1_BYTES);
      for (int kc = 0; kc < K1C; ++kc) tma_load_2d(&tmap_w3, w_bar, sW3 + kc * C::W3_CHUNK, kc * 64, 0);
      for (int c = 0; c < 4; ++c) tma_load_2d(&tmap_w1, w_bar, sW1 + c * C::W1_CHUNK, c * 64, 0);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int n = 0; n < my_tiles; ++n) {
#pragma unroll
      for (int kc = 0; kc < K1C; ++kc) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[stage], C::A_BYTES);
          tma_load_2d(kc == 0 ? &tmap_a : &tmap_a2, &a_full[stage], sA + stage * C::A_BYTES, 0, tile_of(n) * 128);
        }
        __syncwarp();
        if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA #1: acc1 = A1 * W3^T
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256);
    mbar_wait(w_bar, 0);
    const uint64_t w3_desc = umma_desc_sw128(smem_u32(sW3));
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    uint32_t stage = 0, phase = 0;
    for (int n = 0; n < my_tiles; ++n) {
      mbar_wait(acc1_empty, (n & 1) ^ 1);  // both epilogue groups have read the previous tile's accumulator
#pragma unroll
      for (int kc = 0; kc < K1C; ++kc) {
        mbar_wait(&a_full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (C::A_BYTES >> 4));
          const uint64_t b_desc = w3_desc + (uint64_t)(kc * (C::W3_CHUNK >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc1, (kc | k) != 0);
          umma_commit(&a_empty[stage]);
          if (kc == K1C - 1) umma_commit(acc1_full);
        }
        __syncwarp();
        if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 10) {
    // ===================================================== manager (one lane: it owns the bulk groups of the stores)
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, N2);
      mbar_wait(w_bar, 0);
      const uint64_t w1_desc = umma_desc_sw128(smem_u32(sW1));
      const uint64_t eb_desc0 = umma_desc_sw128(smem_u32(sEB));
      const uint32_t total = (uint32_t)my_tiles * 4;
      for (uint32_t i = 0; i < total + D; ++i) {
        if (i < total) {
          const uint32_t s = i % NB;
          if (i >= (uint32_t)NB) {
            bulk_wait_group_read<1>();                                  // store of sub-tile i - NB has drained
            mbar_wait(&eb_mma_done[s], ((i / NB) - 1) & 1);             // ... and MMA #2 has read it
          }
          const int n = i >> 2, c = i & 3;
          if (RES) {
            mbar_expect_tx(&eb_full[s], C::EB_BYTES);
            tma_load_2d(&tmap_res, &eb_full[s], sEB + s * C::EB_BYTES, c * 64, tile_of(n) * 128);
          } else {
            mbar_arrive(&eb_full[s]);  // no residual: the buffer is simply free
          }
        }
        if (i >= (uint32_t)D) {
          const uint32_t qs = i - D;
          const uint32_t s = qs % NB, ph = (qs / NB) & 1;
          const int n = qs >> 2, c = qs & 3;
          const uint32_t abuf = n & 1;
          if (c == 0) {  // epilogue #2 of the tile that used this accumulator stage two tiles ago is done
            mbar_wait(&acc2_empty[abuf], ((n >> 1) & 1) ^ 1);
          }
          mbar_wait(&eb_ready[s], ph);
          tc_fence_after();
          tma_store_2d(&tmap_out, sEB + s * C::EB_BYTES, c * 64, tile_of(n) * 128);
          bulk_commit_group();
          const uint64_t a_desc = eb_desc0 + (uint64_t)(s * (C::EB_BYTES >> 4));
          const uint64_t b_desc = w1_desc + (uint64_t)(c * (C::W1_CHUNK >> 4));
          const uint32_t d_tmem = tmem_base + 256 + abuf * N2;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc2, (c | k) != 0);
          umma_commit(&eb_mma_done[s]);
          if (c == 3) umma_commit(&acc2_full[abuf]);
        }
      }
      bulk_wait_group<0>();
    }
  } else {
    // ===================================================== epilogue groups
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int gtid = (warp - 2 - group * 4) * 32 + lane;
    const bool leader = gtid == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t sb_addr = smem_u32(sSB);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    for (int n = 0; n < my_tiles; ++n) {
      mbar_wait(acc1_full, n & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = group; c < 4; c += 2) {  // sub-tiles c = group, group + 2
        const uint32_t q = (uint32_t)n * 4 + c;
        const uint32_t s = q % NB, ph = (q / NB) & 1;
        const uint32_t eb_row = smem_u32(sEB + s * C::EB_BYTES) + row * 128;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + lane_off + c * 64, v);
        mbar_wait(&eb_full[s], ph);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float f[32];
          tmem_wait_ld();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint4 s4 = ld_shared_v4(sb_addr + (c * 64 + h * 32 + jj * 4) * 4);
            const uint4 b4 = ld_shared_v4(sb_addr + 1024 + (c * 64 + h * 32 + jj * 4) * 4);
            f[4 * jj + 0] = fmaf(__uint_as_float(v[4 * jj + 0]), __uint_as_float(s4.x), __uint_as_float(b4.x));
            f[4 * jj + 1] = fmaf(__uint_as_float(v[4 * jj + 1]), __uint_as_float(s4.y), __uint_as_float(b4.y));
            f[4 * jj + 2] = fmaf(__uint_as_float(v[4 * jj + 2]), __uint_as_float(s4.z), __uint_as_float(b4.z));
            f[4 * jj + 3] = fmaf(__uint_as_float(v[4 * jj + 3]), __uint_as_float(s4.w), __uint_as_float(b4.w));
          }
          if (h == 0) tmem_ld_32x32b_x32(tmem_base + lane_off + c * 64 + 32, v);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t addr = eb_row + (((h * 4 + jj) ^ swz) << 4);
            const uint4 rv = RES ? ld_shared_v4(addr) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
            uint32_t o[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              o[t] = b2b_relu2(b2b_pack(f[8 * jj + 2 * t] + __uint_as_float(w[t] << 16),
                                        f[8 * jj + 2 * t + 1] + __uint_as_float(w[t] & 0xFFFF0000u)));
            st_shared_v4(addr, make_uint4(o[0], o[1], o[2], o[3]));
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + group, 128);
        if (leader) mbar_arrive(&eb_ready[s]);
      }
      tc_fence_before();
      mbar_arrive(acc1_empty);
      if ((n & 1) == group) {
        // ---- epilogue #2: t1' = relu(bn1'(acc2)) for this tile, N2 / 64 halves through one staging buffer
        const uint32_t abuf = n & 1;
        mbar_wait(&acc2_full[abuf], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t s2_row = smem_u32(sS2) + row * 128;
#pragma unroll 1
        for (int hh = 0; hh < N2 / 64; ++hh) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + abuf * N2 + hh * 64, v);
          if (leader) bulk_wait_group_read<0>();  // this group's previous t1' store has drained the staging buffer
          named_bar_sync(1 + group, 128);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[32];
            tmem_wait_ld();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const uint4 s4 = ld_shared_v4(sb_addr + 2048 + (hh * 64 + h * 32 + jj * 4) * 4);
              const uint4 b4 = ld_shared_v4(sb_addr + 2048 + N2 * 4 + (hh * 64 + h * 32 + jj * 4) * 4);
              f[4 * jj + 0] = fmaf(__uint_as_float(v[4 * jj + 0]), __uint_as_float(s4.x), __uint_as_float(b4.x));
              f[4 * jj + 1] = fmaf(__uint_as_float(v[4 * jj + 1]), __uint_as_float(s4.y), __uint_as_float(b4.y));
              f[4 * jj + 2] = fmaf(__uint_as_float(v[4 * jj + 2]), __uint_as_float(s4.z), __uint_as_float(b4.z));
              f[4 * jj + 3] = fmaf(__uint_as_float(v[4 * jj + 3]), __uint_as_float(s4.w), __uint_as_float(b4.w));
            }
            if (h == 0) tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + abuf * N2 + hh * 64 + 32, v);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              uint4 ov;
              ov.x = b2b_relu2(b2b_pack(f[8 * jj + 0], f[8 * jj + 1]));
              ov.y = b2b_relu2(b2b_pack(f[8 * jj + 2], f[8 * jj + 3]));
              ov.z = b2b_relu2(b2b_pack(f[8 * jj + 4], f[8 * jj + 5]));
              ov.w = b2b_relu2(b2b_pack(f[8 * jj + 6], f[8 * jj + 7]));
              st_shared_v4(s2_row + (((h * 4 + jj) ^ swz) << 4), ov);
            }
          }
          if (hh == N2 / 64 - 1) {
            tc_fence_before();
            mbar_arrive(&acc2_empty[abuf]);
          }
          fence_proxy_async();
          named_bar_sync(1 + group, 128);
          if (leader) {
            tma_store_2d(&tmap_out2, sS2, hh * 64, tile_of(n) * 128);
            bulk_commit_group();
          }
        }
      }
    }
    if (leader) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int N2, int K1C, bool RES>
cudaError_t launch_b2b(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tw3, const CUtensorMap& tres,
                       const CUtensorMap& tout, const CUtensorMap& tw1, const CUtensorMap& tout2,
                       const ConvB2BParams& p, int num_sms, cudaStream_t stream) {
  auto kern = conv_b2b_kernel<N2, K1C, RES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B2BCfg<N2, K1C>::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_m_tiles < num_sms ? p.num_m_tiles : num_sms);
  cfg.blockDim = dim3(B2B_THREADS);
  cfg.dynamicSmemBytes = B2BCfg<N2, K1C>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, ta, ta2, tw3, tres, tout, tw1, tout2, p);
}


// ------------------------------------------------------------------------------------------------------------------
// Layer2 variant: conv3 128 -> 512 (+ residual), next conv1 512 -> 128. The weights (128 KB + 128 KB) do not fit shared
// memory next to the staging buffers, so they are streamed from L2 through small rings: W3 in four 256 x 64 blocks per
// tile (two per 256-column half of `out`), W1' in eight 128 x 64 chunks (one per finished sub-tile). acc1 is reused
// for the two halves; 384 threads = the roles above + warp 11, the W1' producer.
constexpr int B2S_THREADS = 384;
struct B2SCfg {
  static constexpr int AS = 3, W3S = 2, W1S = 2, NB = 3;
  static constexpr uint32_t A_BYTES = 16384, W3_BYTES = 32768, W1_BYTES = 16384, EB_BYTES = 16384;
  static constexpr uint32_t SB_BYTES = (1024 + 256) * 4;  // scale1[512] | bias1[512] | scale2[128] | bias2[128]
  static constexpr uint32_t SMEM = 1024 + AS * A_BYTES + W3S * W3_BYTES + W1S * W1_BYTES + NB * EB_BYTES + EB_BYTES +
                                   ((SB_BYTES + 1023) / 1024) * 1024 + 512;
  static_assert(SMEM <= 232448, "exceeds the 227 KiB shared memory of one CTA");
};

__global__ void __launch_bounds__(B2S_THREADS, 1)
conv_b2b_stream_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w3,
                       const __grid_constant__ CUtensorMap tmap_res, const __grid_constant__ CUtensorMap tmap_out,
                       const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_out2,
                       const ConvB2BParams p) {
  using C = B2SCfg;
  constexpr int AS = C::AS, W3S = C::W3S, W1S = C::W1S, NB = C::NB, N2 = 128;
  constexpr int D = NB - 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW3 = sA + AS * C::A_BYTES;
  uint8_t* sW1 = sW3 + W3S * C::W3_BYTES;
  uint8_t* sEB = sW1 + W1S * C::W1_BYTES;
  uint8_t* sS2 = sEB + NB * C::EB_BYTES;
  float* sSB = reinterpret_cast<float*>(sS2 + C::EB_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sSB) + ((C::SB_BYTES + 1023) / 1024) * 1024);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + AS;
  uint64_t* w3_full = a_empty + AS;
  uint64_t* w3_empty = w3_full + W3S;
  uint64_t* w1_full = w3_empty + W3S;
  uint64_t* w1_empty = w1_full + W1S;
  uint64_t* acc1_full = w1_empty + W1S;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* acc2_full = acc1_empty + 1;   // [2]
  uint64_t* acc2_empty = acc2_full + 2;   // [2]
  uint64_t* eb_full = acc2_empty + 2;     // [NB]
  uint64_t* eb_ready = eb_full + NB;      // [NB]
  uint64_t* eb_mma_done = eb_ready + NB;  // [NB]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(eb_mma_done + NB);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles;
  const int my_tiles = (int)blockIdx.x < num_tiles ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto tile_of = [&](int n) {
    const int t = (int)blockIdx.x + n * (int)gridDim.x;
    return p.reverse ? num_tiles - 1 - t : t;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w3);
    prefetch_tmap(&tmap_res);
    prefetch_tmap(&tmap_out);
    prefetch_tmap(&tmap_w1);
    prefetch_tmap(&tmap_out2);
    for (int s = 0; s < AS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < W3S; ++s) { mbar_init(&w3_full[s], 1); mbar_init(&w3_empty[s], 1); }
    for (int s = 0; s < W1S; ++s) { mbar_init(&w1_full[s], 1); mbar_init(&w1_empty[s], 1); }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 256);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc2_full[s], 1); mbar_init(&acc2_empty[s], 128); }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&eb_full[s], 1);
      mbar_init(&eb_ready[s], 1);
      mbar_init(&eb_mma_done[s], 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 1024 + 256; i += blockDim.x)
    sSB[i] = i < 512 ? p.scale1[i] : i < 1024 ? p.bias1[i - 512] : i < 1152 ? p.scale2[i - 1024] : p.bias2[i - 1152];
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===================================================== producer: A chunks of the tile, then its four W3 blocks
    uint32_t as = 0, aph = 0, ws = 0, wph = 0;
    for (int n = 0; n < my_tiles; ++n) {
      const int m0 = tile_of(n) * 128;
      for (int kc = 0; kc < 2; ++kc) {
        mbar_wait(&a_empty[as], aph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&a_full[as], C::A_BYTES);
          tma_load_2d(&tmap_a, &a_full[as], sA + as * C::A_BYTES, kc * 64, m0);
        }
        __syncwarp();
        if (++as == AS) { as = 0; aph ^= 1; }
      }
      for (int b = 0; b < 4; ++b) {  // block b: output columns [256 * (b >> 1), +256), K chunk b & 1
        mbar_wait(&w3_empty[ws], wph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&w3_full[ws], C::W3_BYTES);
          tma_load_2d(&tmap_w3, &w3_full[ws], sW3 + ws * C::W3_BYTES, (b & 1) * 64, (b >> 1) * 256);
        }
        __syncwarp();
        if (++ws == W3S) { ws = 0; wph ^= 1; }
      }
    }
  } else if (warp == 11) {
    // ===================================================== W1' producer: one 128 x 64 chunk per sub-tile of `out`
    uint32_t s = 0, ph = 0;
    for (int n = 0; n < my_tiles; ++n)
      for (int j = 0; j < 8; ++j) {
        mbar_wait(&w1_empty[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&w1_full[s], C::W1_BYTES);
          tma_load_2d(&tmap_w1, &w1_full[s], sW1 + s * C::W1_BYTES, j * 64, 0);
        }
        __syncwarp();
        if (++s == W1S) { s = 0; ph ^= 1; }
      }
  } else if (warp == 1) {
    // ===================================================== MMA #1: acc1 = A (2 chunks) * W3[half]^T, twice per tile
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256);
    const uint64_t a_desc0 = umma_desc_sw128(smem_u32(sA));
    const uint64_t w_desc0 = umma_desc_sw128(smem_u32(sW3));
    uint32_t as = 0, aph = 0, ws = 0, wph = 0, hc = 0;
    for (int n = 0; n < my_tiles; ++n) {
      uint32_t slot[2];
      for (int h = 0; h < 2; ++h, ++hc) {
        mbar_wait(acc1_empty, (hc & 1) ^ 1);  // both epilogue groups have read the previous half
        for (int kc = 0; kc < 2; ++kc) {
          if (h == 0) {
            slot[kc] = as;
            mbar_wait(&a_full[as], aph);
            if (++as == AS) { as = 0; aph ^= 1; }
          }
          mbar_wait(&w3_full[ws], wph);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_desc = a_desc0 + (uint64_t)(slot[kc] * (C::A_BYTES >> 4));
            const uint64_t b_desc = w_desc0 + (uint64_t)(ws * (C::W3_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc1, (kc | k) != 0);
            umma_commit(&w3_empty[ws]);
            if (kc == 1) {
              umma_commit(acc1_full);
              if (h == 1) {  // the tile's A chunks are free once the second half has been multiplied
                umma_commit(&a_empty[slot[0]]);
                umma_commit(&a_empty[slot[1]]);
              }
            }
          }
          __syncwarp();
          if (++ws == W3S) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ===================================================== manager: residual prefetch, stores of `out`, MMA #2
    if (lane == 0) {
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, N2);
      const uint64_t w1_desc0 = umma_desc_sw128(smem_u32(sW1));
      const uint64_t eb_desc0 = umma_desc_sw128(smem_u32(sEB));
      const uint32_t total = (uint32_t)my_tiles * 8;
      uint32_t w1s = 0, w1ph = 0;
      for (uint32_t i = 0; i < total + D; ++i) {
        if (i < total) {
          const uint32_t s = i % NB;
          if (i >= (uint32_t)NB) {
            bulk_wait_group_read<1>();
            mbar_wait(&eb_mma_done[s], ((i / NB) - 1) & 1);
          }
          const int n = i >> 3, c8 = i & 7;
          mbar_expect_tx(&eb_full[s], C::EB_BYTES);
          tma_load_2d(&tmap_res, &eb_full[s], sEB + s * C::EB_BYTES, c8 * 64, tile_of(n) * 128);
        }
        if (i >= (uint32_t)D) {
          const uint32_t qs = i - D;
          const uint32_t s = qs % NB, ph = (qs / NB) & 1;
          const int n = qs >> 3, c8 = qs & 7;
          const uint32_t abuf = n & 1;
          if (c8 == 0) mbar_wait(&acc2_empty[abuf], ((n >> 1) & 1) ^ 1);
          mbar_wait(&eb_ready[s], ph);
          mbar_wait(&w1_full[w1s], w1ph);
          tc_fence_after();
          tma_store_2d(&tmap_out, sEB + s * C::EB_BYTES, c8 * 64, tile_of(n) * 128);
          bulk_commit_group();
          const uint64_t a_desc = eb_desc0 + (uint64_t)(s * (C::EB_BYTES >> 4));
          const uint64_t b_desc = w1_desc0 + (uint64_t)(w1s * (C::W1_BYTES >> 4));
          const uint32_t d_tmem = tmem_base + 256 + abuf * N2;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, a_desc + (k * 32 >> 4), b_desc + (k * 32 >> 4), idesc2, (c8 | k) != 0);
          umma_commit(&eb_mma_done[s]);
          umma_commit(&w1_empty[w1s]);
          if (c8 == 7) umma_commit(&acc2_full[abuf]);
          if (++w1s == W1S) { w1s = 0; w1ph ^= 1; }
        }
      }
      bulk_wait_group<0>();
    }
  } else {
    // ===================================================== epilogue groups
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int gtid = (warp - 2 - group * 4) * 32 + lane;
    const bool leader = gtid == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t sb_addr = smem_u32(sSB);
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    uint32_t hc = 0;
    for (int n = 0; n < my_tiles; ++n) {
      for (int h = 0; h < 2; ++h, ++hc) {
        mbar_wait(acc1_full, hc & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = group; c < 4; c += 2) {
          const uint32_t q = (uint32_t)n * 8 + h * 4 + c;
          const uint32_t s = q % NB, ph = (q / NB) & 1;
          const uint32_t eb_row = smem_u32(sEB + s * C::EB_BYTES) + row * 128;
          const int col0 = h * 256 + c * 64;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + c * 64, v);
          mbar_wait(&eb_full[s], ph);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float f[32];
            tmem_wait_ld();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const uint4 s4 = ld_shared_v4(sb_addr + (col0 + hh * 32 + jj * 4) * 4);
              const uint4 b4 = ld_shared_v4(sb_addr + 2048 + (col0 + hh * 32 + jj * 4) * 4);
              f[4 * jj + 0] = fmaf(__uint_as_float(v[4 * jj + 0]), __uint_as_float(s4.x), __uint_as_float(b4.x));
              f[4 * jj + 1] = fmaf(__uint_as_float(v[4 * jj + 1]), __uint_as_float(s4.y), __uint_as_float(b4.y));
              f[4 * jj + 2] = fmaf(__uint_as_float(v[4 * jj + 2]), __uint_as_float(s4.z), __uint_as_float(b4.z));
              f[4 * jj + 3] = fmaf(__uint_as_float(v[4 * jj + 3]), __uint_as_float(s4.w), __uint_as_float(b4.w));
            }
            if (hh == 0) tmem_ld_32x32b_x32(tmem_base + lane_off + c * 64 + 32, v);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint32_t addr = eb_row + (((hh * 4 + jj) ^ swz) << 4);
              const uint4 rv = ld_shared_v4(addr);
              const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
              uint32_t o[4];
#pragma unroll
              for (int t = 0; t < 4; ++t)
                o[t] = b2b_relu2(b2b_pack(f[8 * jj + 2 * t] + __uint_as_float(w[t] << 16),
                                          f[8 * jj + 2 * t + 1] + __uint_as_float(w[t] & 0xFFFF0000u)));
              st_shared_v4(addr, make_uint4(o[0], o[1], o[2], o[3]));
            }
          }
          fence_proxy_async();
          named_bar_sync(1 + group, 128);
          if (leader) mbar_arrive(&eb_ready[s]);
        }
        tc_fence_before();
        mbar_arrive(acc1_empty);
      }
      if ((n & 1) == group) {
        // ---- epilogue #2: t1' = relu(bn1'(acc2)), two 64-column halves through one staging buffer
        const uint32_t abuf = n & 1;
        mbar_wait(&acc2_full[abuf], (n >> 1) & 1);
        tc_fence_after();
        const uint32_t s2_row = smem_u32(sS2) + row * 128;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + abuf * N2 + hh * 64, v);
          if (leader) bulk_wait_group_read<0>();
          named_bar_sync(1 + group, 128);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float f[32];
            tmem_wait_ld();
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const uint4 s4 = ld_shared_v4(sb_addr + 4096 + (hh * 64 + h * 32 + jj * 4) * 4);
              const uint4 b4 = ld_shared_v4(sb_addr + 4096 + 512 + (hh * 64 + h * 32 + jj * 4) * 4);
              f[4 * jj + 0] = fmaf(__uint_as_float(v[4 * jj + 0]), __uint_as_float(s4.x), __uint_as_float(b4.x));
              f[4 * jj + 1] = fmaf(__uint_as_float(v[4 * jj + 1]), __uint_as_float(s4.y), __uint_as_float(b4.y));
              f[4 * jj + 2] = fmaf(__uint_as_float(v[4 * jj + 2]), __uint_as_float(s4.z), __uint_as_float(b4.z));
              f[4 * jj + 3] = fmaf(__uint_as_float(v[4 * jj + 3]), __uint_as_float(s4.w), __uint_as_float(b4.w));
            }
            if (h == 0) tmem_ld_32x32b_x32(tmem_base + lane_off + 256 + abuf * N2 + hh * 64 + 32, v);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              uint4 ov;
              ov.x = b2b_relu2(b2b_pack(f[8 * jj + 0], f[8 * jj + 1]));
              ov.y = b2b_relu2(b2b_pack(f[8 * jj + 2], f[8 * jj + 3]));
              ov.z = b2b_relu2(b2b_pack(f[8 * jj + 4], f[8 * jj + 5]));
              ov.w = b2b_relu2(b2b_pack(f[8 * jj + 6], f[8 * jj + 7]));
              st_shared_v4(s2_row + (((h * 4 + jj) ^ swz) << 4), ov);
            }
          }
          if (hh == 1) {
            tc_fence_before();
            mbar_arrive(&acc2_empty[abuf]);
          }
          fence_proxy_async();
          named_bar_sync(1 + group, 128);
          if (leader) {
            tma_store_2d(&tmap_out2, sS2, hh * 64, tile_of(n) * 128);
            bulk_commit_group();
          }
        }
      }
    }
    if (leader) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

cudaError_t launch_b2b_stream(const CUtensorMap& ta, const CUtensorMap& tw3, const CUtensorMap& tres,
                              const CUtensorMap& tout, const CUtensorMap& tw1, const CUtensorMap& tout2,
                              const ConvB2BParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_b2b_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         B2SCfg::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.num_m_tiles < num_sms ? p.num_m_tiles : num_sms);
  cfg.blockDim = dim3(B2S_THREADS);
  cfg.dynamicSmemBytes = B2SCfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv_b2b_stream_kernel, ta, tw3, tres, tout, tw1, tout2, p);
}

}  // namespace

cudaError_t launch_conv_b2b(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tw3,
                            const CUtensorMap& tres, const CUtensorMap& tout, const CUtensorMap& tw1,
                            const CUtensorMap& tout2, const ConvB2BParams& p, int num_sms, cudaStream_t stream) {
  if (p.streamed) return launch_b2b_stream(ta, tw3, tres, tout, tw1, tout2, p, num_sms, stream);
  if (p.k1_chunks == 1 && p.n2 == 64)
    return launch_b2b<64, 1, true>(ta, ta2, tw3, tres, tout, tw1, tout2, p, num_sms, stream);
  if (p.k1_chunks == 1 && p.n2 == 128)
    return launch_b2b<128, 1, true>(ta, ta2, tw3, tres, tout, tw1, tout2, p, num_sms, stream);
  if (p.k1_chunks == 2 && p.n2 == 64)
    return launch_b2b<64, 2, false>(ta, ta2, tw3, tres, tout, tw1, tout2, p, num_sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace pvr
