// 3x3 / stride-1 / pad-1 convolution with C_in = C_out = 64 (ResNet-50 layer1.*.conv2 at 56x56) as a patch-resident
// tcgen05 implicit GEMM.
//
// The im2col formulation (conv_gemm.cu, A_IM2COL64) re-reads every input pixel nine times from L2 into shared memory:
// at C = 64 the kernel is bound by L2 -> SM bytes (profiles/: 153 us per 256 frames against 45 us of tensor time).
// Here one tile = 16 output rows x 8 output columns (= 128 GEMM rows). The producer loads the 18 x 8 input patch
// three times, shifted by -1 / 0 / +1 columns (tiled 4-D TMA, out-of-range = zero = padding); each copy is
// [18 rows][8 px][128 B], i.e. one 1024-byte swizzle atom per patch row, so the A operand of tap (r, s) is simply
// copy s at byte offset r * 1024: nine taps, three loads. The 72 KiB of weights stay resident in shared memory for the
// whole (persistent) kernel. Epilogue: as conv_gemm (two groups, smem staging, TMA store of the 16 x 8 output patch).
#include "conv_gemm.cuh"
#include "ptx.cuh"

namespace pvr {
namespace {

// warp 0 TMA, warp 1 MMA, then GROUPS epilogue groups of 4 warps (one TMEM accumulator + one staging buffer each)

// STEM = false: 3x3/s1 conv, 64 -> 64 channels. Three column-shifted copies of the 18 x 8 patch, 128-byte pixels,
//               128B swizzle (one 1024-byte atom per patch row), nine 64 x 64 tap matrices resident.
// STEM = true : ResNet stem 7x7/s2 over the W-expanded input (PVR_FMT_STEM_BF16: 64-byte pixels holding the 8 input
//               columns x 4 channels of one output column). Only row taps remain; output row p reads input rows
//               2p-3+r, so the even taps (r = 0,2,4,6) read every second row starting at 2*p0-3 and the odd taps every
//               second row starting at 2*p0-2: two TMA loads with traversal stride 2 (19 rows each), 64B swizzle (one
//               512-byte atom per patch row); tap r = copy (r & 1) at byte offset (r >> 1) * 512. Seven 64 x 32 tap
//               matrices resident. im2col re-read 7x -> 2.3x.
// MODE 2      : the stem with the 3x3/s2/pad-1 max pool (torchvision resnet.py:271) fused into the epilogue. A tile of
//               16 x 8 stem pixels starts at (14*tp - 1, 6*tq - 1) and yields the 7 x 3 pooled pixels whose windows it
//               contains (tiles overlap by the pool halo: 1.5x stem math, which is free next to the HBM time); the
//               112 x 112 x 64 stem activation (411 MB per 256 frames, written and re-read) never reaches HBM.
//               Stem pixels outside the image are replaced by 0, which is neutral for a max over ReLU outputs.
template <int MODE>
struct PCfg {
  static constexpr bool STEM = MODE >= 1;
  static constexpr uint32_t ROW_BYTES = STEM ? 512 : 1024;          // 8 pixels of one patch row
  static constexpr uint32_t COPY_ROWS = STEM ? 19 : 18;
  static constexpr uint32_t COPY_BYTES = STEM ? 10240 : 18 * 1024;  // one copy of the patch (1024-byte aligned)
  static constexpr uint32_t COPY_TX = COPY_ROWS * ROW_BYTES;        // bytes one TMA load delivers
  static constexpr int COPIES = STEM ? 2 : 3;
  static constexpr uint32_t A_STAGE = COPIES * COPY_BYTES;
  static constexpr int TAPS = STEM ? 7 : 9;
  static constexpr uint32_t TAP_BYTES = STEM ? 64 * 64 : 64 * 128;  // one resident tap matrix
  static constexpr uint32_t W_BYTES = ((TAPS * TAP_BYTES + 1023) / 1024) * 1024;
  static constexpr int A_STAGES = STEM ? 4 : 2;
  // The stem epilogue (64 K-steps of math per 128 x 64 tile against 576 for the 3x3) is latency bound per group:
  // four groups keep four tiles in flight (measured: 2 groups = 0.72 us per tile and CTA with loads and MMAs removed).
  static constexpr int GROUPS = STEM ? 4 : 2;
  static constexpr int THREADS = 64 + 128 * GROUPS;
  static constexpr uint32_t SMEM = 1024 + W_BYTES + A_STAGES * A_STAGE + GROUPS * 16384 + 512 + 256;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

template <int MODE>
__global__ void __launch_bounds__(PCfg<MODE>::THREADS, 1)
conv3x3_patch_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                     const __grid_constant__ CUtensorMap tmap_out, const Conv3x3PatchParams p) {
  using C = PCfg<MODE>;
  constexpr bool STEM = MODE >= 1, POOL = MODE == 2;
  constexpr uint32_t COPY_BYTES = C::COPY_BYTES, A_STAGE = C::A_STAGE, W_BYTES = C::W_BYTES;
  constexpr int A_STAGES = C::A_STAGES, GROUPS = C::GROUPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sA = sW + W_BYTES;
  uint8_t* sOut = sA + A_STAGES * A_STAGE;
  float* sSB = reinterpret_cast<float*>(sOut + GROUPS * 16384);  // scale[64] | bias[64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sOut + GROUPS * 16384 + 512);
  uint64_t* empty_bar = full_bar + A_STAGES;
  uint64_t* w_bar = empty_bar + A_STAGES;
  uint64_t* tmem_full_bar = w_bar + 1;
  uint64_t* tmem_empty_bar = tmem_full_bar + GROUPS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + GROUPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_p * p.tiles_q;
  const int num_tiles = p.n_img * tiles_per_img;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_in);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_out);
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    for (int s = 0; s < GROUPS; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 128);  // only the group that owns the tile reads the accumulator
    }
    fence_barrier_init();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 192) sSB[threadIdx.x - 64] =
      (threadIdx.x - 64) < 64 ? p.scale[threadIdx.x - 64] : p.bias[threadIdx.x - 128];
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 64 * GROUPS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (broadcast from lane 0: the compiler then keeps the TMEM address in a uniform register for tcgen05.mma)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  griddep_launch();  // programmatic dependent launch: see conv_gemm.cu
  griddep_wait();

  if (warp == 0) {
    // producer: warp-uniform loop, one elected lane issues the TMA loads
    if (elect_one()) {
      // weights: the tap matrices, loaded once
      mbar_expect_tx(w_bar, C::TAPS * C::TAP_BYTES);
      for (int tap = 0; tap < C::TAPS; ++tap)
        tma_load_2d(&tmap_w, w_bar, sW + tap * C::TAP_BYTES, tap * (STEM ? 32 : 64), 0);
    }
    __syncwarp();
    uint32_t stage = 0, phase = 0;
    for (int it = blockIdx.x; it < num_tiles; it += gridDim.x) {
      const int tile = p.reverse ? num_tiles - 1 - it : it;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int tp = rem / p.tiles_q, tq = rem - tp * p.tiles_q;
      const int p0 = POOL ? 14 * tp - 1 : tp * 16, q0 = POOL ? 6 * tq - 1 : tq * 8;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full_bar[stage], C::COPIES * C::COPY_TX);
        if (STEM) {
#pragma unroll
          for (int cls = 0; cls < 2; ++cls)  // even / odd row taps: every second input row from 2*p0-3 / 2*p0-2
            tma_load_4d(&tmap_in, &full_bar[stage], sA + stage * A_STAGE + cls * COPY_BYTES, 0, q0, 2 * p0 - 3 + cls,
                        img);
        } else {
#pragma unroll
          for (int s = 0; s < 3; ++s)
            tma_load_4d(&tmap_in, &full_bar[stage], sA + stage * A_STAGE + s * COPY_BYTES, 0, q0 + s - 1, p0 - 1, img);
        }
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the tile loop (uniform loop state), one elected lane issues
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    mbar_wait(w_bar, 0);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    // descriptors are built once; taps / K-steps / stages only add to the 14-bit (address >> 4) field
    const uint64_t w_desc = STEM ? umma_desc_sw64(smem_u32(sW)) : umma_desc_sw128(smem_u32(sW));
    const uint64_t a_desc0 = STEM ? umma_desc_sw64(smem_u32(sA)) : umma_desc_sw128(smem_u32(sA));
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (A_STAGE >> 4));
        const uint32_t d_tmem = tmem_base + acc * 64;
        if (STEM) {
#pragma unroll
          for (int r = 0; r < 7; ++r)
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma_bf16(d_tmem, a_desc + (((r & 1) * COPY_BYTES + (r >> 1) * 512 + k * 32) >> 4),
                        w_desc + ((r * C::TAP_BYTES + k * 32) >> 4), idesc, (r | k) != 0);
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, a_desc + ((s * COPY_BYTES + r * 1024 + k * 32) >> 4),
                          w_desc + (((r * 3 + s) * C::TAP_BYTES + k * 32) >> 4), idesc, (r | s | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
      if (++acc == GROUPS) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // epilogue: group g takes tiles with (local tile index % GROUPS) == g, accumulator stage == g
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int gtid = (warp - 2 - group * 4) * 32 + lane;
    const bool leader = gtid == 0;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t out_row = smem_u32(sOut + group * 16384) + row * 128;
    const uint32_t sb_addr = smem_u32(sSB);
    uint32_t n = 0;
    for (int it = blockIdx.x; it < num_tiles; it += gridDim.x, ++n) {
      if (n % GROUPS != (uint32_t)group) continue;
      const uint32_t acc_phase = (n / GROUPS) & 1;
      const int tile = p.reverse ? num_tiles - 1 - it : it;
      const int img = tile / tiles_per_img;
      const int rem = tile - img * tiles_per_img;
      const int tp = rem / p.tiles_q, tq = rem - tp * p.tiles_q;
      const int p0 = POOL ? 14 * tp - 1 : tp * 16, q0 = POOL ? 6 * tq - 1 : tq * 8;
      mbar_wait(&tmem_full_bar[group], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + group * 64 + ((uint32_t)(quarter * 32) << 16);
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr, v);
      if (!POOL && leader) bulk_wait_group_read<0>();  // this group's previous store has drained the staging buffer
      named_bar_sync(1 + group, 128);  // (POOL: every thread of the group is done reading the previous tile)
      // POOL: stem pixels outside the image take no part in the max
      const bool inside = !POOL || ((unsigned)(p0 + (row >> 3)) < (unsigned)p.P && (unsigned)(q0 + (row & 7)) < (unsigned)p.Q);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float f[32];
        tmem_wait_ld();
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const uint4 s4 = ld_shared_v4(sb_addr + (h * 8 + jj) * 16);
          const uint4 b4 = ld_shared_v4(sb_addr + 256 + (h * 8 + jj) * 16);
          f[4 * jj + 0] = fmaf(__uint_as_float(v[4 * jj + 0]), __uint_as_float(s4.x), __uint_as_float(b4.x));
          f[4 * jj + 1] = fmaf(__uint_as_float(v[4 * jj + 1]), __uint_as_float(s4.y), __uint_as_float(b4.y));
          f[4 * jj + 2] = fmaf(__uint_as_float(v[4 * jj + 2]), __uint_as_float(s4.z), __uint_as_float(b4.z));
          f[4 * jj + 3] = fmaf(__uint_as_float(v[4 * jj + 3]), __uint_as_float(s4.w), __uint_as_float(b4.w));
        }
        if (h == 0) {
          tmem_ld_32x32b_x32(taddr + 32, v);  // second half of the columns, in flight during the math below
        } else {
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[group]);
        }
        if (p.relu) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) f[jj] = fmaxf(f[jj], 0.f);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          uint4 ov;
          ov.x = pack2(f[8 * jj + 0], f[8 * jj + 1]);
          ov.y = pack2(f[8 * jj + 2], f[8 * jj + 3]);
          ov.z = pack2(f[8 * jj + 4], f[8 * jj + 5]);
          ov.w = pack2(f[8 * jj + 6], f[8 * jj + 7]);
          if (POOL && !inside) ov = make_uint4(0u, 0u, 0u, 0u);
          st_shared_v4(out_row + (((h * 4 + jj) ^ swz) << 4), ov);
        }
      }
      if (POOL) {
        named_bar_sync(1 + group, 128);
        // 7 x 3 pooled pixels x 8 chunks of 8 channels: max over the 3 x 3 staged stem pixels, 16-byte global stores
        // (the 8 threads of a pixel write one full 128-byte line)
        const uint32_t buf = smem_u32(sOut + group * 16384);
        for (int item = gtid; item < 21 * 8; item += 128) {
          const int chunk = item & 7, px = item >> 3;
          const int pi = px / 3, pj = px - 3 * pi;
          const int pp = 7 * tp + pi, pq = 3 * tq + pj;
          if (pp >= p.pool_P || pq >= p.pool_Q) continue;
          __nv_bfloat162 m[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) m[e] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
          for (int dr = 0; dr < 3; ++dr)
#pragma unroll
            for (int dc = 0; dc < 3; ++dc) {
              const int r2 = (2 * pi + dr) * 8 + 2 * pj + dc;
              const uint4 v4 = ld_shared_v4(buf + r2 * 128 + ((chunk ^ (r2 & 7)) << 4));
              const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v4);
#pragma unroll
              for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], hv[e]);
            }
          __nv_bfloat16* dst = p.pool_out + ((((long long)img * p.pool_P + pp) * p.pool_Q + pq) * 64 + chunk * 8);
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(m);
        }
      } else {
        fence_proxy_async();
        named_bar_sync(1 + group, 128);
        if (leader) {
          tma_store_4d(&tmap_out, sOut + group * 16384, 0, q0, p0, img);  // rows >= P are clipped
          bulk_commit_group();
        }
      }
    }
    if (!POOL && leader) bulk_wait_group<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64 * GROUPS);
  }
}

}  // namespace

template <int MODE>
cudaError_t launch_patch(const CUtensorMap& tmap_in, const CUtensorMap& tmap_w, const CUtensorMap& tmap_out,
                         const Conv3x3PatchParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_patch_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         PCfg<MODE>::SMEM);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int tiles = p.n_img * p.tiles_p * p.tiles_q;
  const int grid = tiles < num_sms ? tiles : num_sms;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(PCfg<MODE>::THREADS);
  cfg.dynamicSmemBytes = PCfg<MODE>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv3x3_patch_kernel<MODE>, tmap_in, tmap_w, tmap_out, p);
}

cudaError_t launch_conv3x3_patch(const CUtensorMap& tmap_in, const CUtensorMap& tmap_w, const CUtensorMap& tmap_out,
                                 const Conv3x3PatchParams& p, int num_sms, cudaStream_t stream) {
  if (p.stem && p.pool_out) return launch_patch<2>(tmap_in, tmap_w, tmap_out, p, num_sms, stream);
  return p.stem ? launch_patch<1>(tmap_in, tmap_w, tmap_out, p, num_sms, stream)
                : launch_patch<0>(tmap_in, tmap_w, tmap_out, p, num_sms, stream);
}

}  // namespace pvr
