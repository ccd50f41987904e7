// Persistent LSTM recurrence (K9, src/models.py:68-72 called T times with seq_len 1; BPTT through the same steps).
//
// The per-step formulation (policy.cu: one tcgen05 GEMM launch + one cell launch per time step, 4 T launches per layer
// for forward + backward) is bound by launch / drain latency: 8.5 us per (128 x 1024) x (1024 x 4096) product, 9 % of
// the tensor peak, W_hh re-read from L2 at every step. Here ONE kernel runs all T steps of a layer:
//
//   * W_hh (8 MB bf16) is cut into 64 x 1024 slices that stay RESIDENT in shared memory (128 KB per CTA) for the whole
//     sequence: 64 slices x ceil(B / 64) batch tiles = 128 CTAs at B = 128, one per SM;
//   * a step's operand (the masked h_{t-1}, or dG_{t+1} in the backward pass; 64 rows x 1024) streams through a TMA
//     ring as 4 groups of 4 (64 x 64) K chunks — one 32 KB box per instruction: the SM's TMA pipe serves its
//     instructions one after the other at ~400 cycles + 2.3 cycles per 128-byte line, so 8 KB boxes cost 40 % more
//     time per step — each group gated by per-chunk arrival counters in global memory that the producing CTAs bump
//     when their part of the previous step is written: a dataflow barrier instead of a kernel boundary
//     (ld.acquire.gpu poll of one 64-byte line, fence.proxy.async, TMA). In the forward pass the four CTAs of a
//     cluster need the same rows: each loads one group and multicasts it (L2 requests / 4);
//   * the (64 x 64) fp32 accumulator lives in TMEM (tcgen05.mma M = 64, N = 64, K = 16 x 64 per step);
//   * clusters of 4 CTAs: CTA q of cluster (batch tile, hidden block J) owns gate q in the forward pass (rows
//     q * 1024 + 64 J.. of W_hh) and K slice q (gate q's rows of W_hh^T) in the backward pass. After the MMAs the four
//     CTAs exchange 16-column pieces through distributed shared memory (st.shared::cluster + remote mbarrier
//     arrive), so that CTA q holds all four gates (forward) / all four partial sums (backward) of hidden units
//     64 J + 16 q .. + 16 and runs the cell update for them with c_t (forward) / dc (backward) kept in registers;
//   * the cell phase writes h_t / masked h_t / gates / c_t (forward) or dG_t (backward) and bumps the counter of its
//     chunk for the next step.
//
// Arithmetic is the per-step path's (bf16 operands, fp32 accumulation, fp32 gates / cell state; same saved tensors, so
// forward and backward implementations can be mixed) except that the K = 1024 / 4096 sums are not split into atomic
// partial sums any more: the result is deterministic.
//
// Safety: every spin (mbarrier, arrival counter) is bounded and traps instead of hanging the GPU; the kernel needs all
// its CTAs co-resident (checked with cudaOccupancyMaxActiveClusters before launch, otherwise the caller falls back to
// the per-step path).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "pvr_b200.h"

extern void pvr_set_error(const char* fmt, ...);

namespace pvr {
namespace {

constexpr int HID = 1024;          // hidden size (PolicyNet: nn.LSTM(1024, 1024, 2))
constexpr int CHUNKS = 16;         // K chunks of 64 per step (K = 1024 per CTA in both passes)
constexpr int CPI = 4;             // K chunks per TMA instruction (one ring stage = CPI chunks)
constexpr int NG = CHUNKS / CPI;   // chunk groups per step
constexpr int NST = 2;             // TMA ring stages
constexpr int CHUNK_BYTES = 64 * 64 * 2;
constexpr int SMEM_W = CHUNKS * CHUNK_BYTES;      // 131072
constexpr int STAGE_BYTES = CPI * CHUNK_BYTES;
constexpr int SMEM_RING = NST * STAGE_BYTES;
constexpr int PIECE_BYTES = 64 * 16 * 4;          // one CTA's (64 rows x 16 columns) fp32 piece
constexpr int SMEM_RECV = 4 * PIECE_BYTES;        // 16384: [source CTA][row][16 fp32]
constexpr int SMEM_STAGE = 4 * PIECE_BYTES;       // 16384: [destination CTA][row][16 fp32], source of the bulk copies
constexpr int SMEM_BARS = 256;
constexpr int SMEM_TOTAL = SMEM_W + SMEM_RING + SMEM_RECV + SMEM_STAGE + SMEM_BARS + 1024;  // + alignment slack
static_assert(SMEM_TOTAL <= 232448, "shared memory budget");
constexpr int NTHREADS = 192;      // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: exchange + cell

struct PersistParams {
  int T, B, nbt;
  const float* nd;       // (T, B)
  float* c_all;          // ((T+1) B, H)
  float* gates;          // (T B, 4H)
  // forward
  const float* xp;       // (T B, 4H)
  __nv_bfloat16* hm;     // (T B, H)
  __nv_bfloat16* h_out;  // (T B, H)
  float* h_last;         // (B, H)
  // backward
  const float* dh_out;   // (T B, H) or null
  float* dh_rec;         // (B, H): in = gradient of the final hidden state, out = gradient of the initial one
  float* dc_rec;         // (B, H): same for the cell state
  __nv_bfloat16* dG;     // (T B, 4H)
  float* dbias;          // optional (4H): += sum over t, b of dG
  unsigned int* ready;   // ((T+1) * nbt * 16) arrival counters, zeroed before the launch
  long long* prof;       // optional (PVR_LSTM_PROF): [block][step < 64][16] clock64 stamps of the phases of a step
};

#define PROF(slot)                                                                                     \
  do {                                                                                                 \
    if (p.prof && s < 64) p.prof[((size_t)blockIdx.x * 64 + s) * 16 + (slot)] = clock64();             \
  } while (0)

// ---- small PTX helpers local to this kernel
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until the `n` (<= 32) arrival counters at `flags` have all reached `need`: one ld.acquire.gpu per lane and
// round trip. (Measured alternatives, both slower by ~1000 cycles per step: relaxed polls + one fence.acq_rel.gpu — the
// fence drains every store the SM has in flight — and several acquire polls in flight per lane.)
__device__ __forceinline__ bool wait_flags(const unsigned int* flags, int n, unsigned int need, int lane) {
  uint64_t t0 = 0;
  uint32_t polls = 0;
  for (;;) {
    unsigned int v = need;
    if (lane < n) v = ld_acquire_gpu(flags + lane);
    if (__all_sync(0xffffffffu, v >= need)) return true;
    if ((++polls & 63u) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) return false;
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++polls & 255u) == 0) {
      const uint64_t t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ull) {
        printf("pvr: lstm_persist exchange barrier timed out (block %d thread %d)\n", (int)blockIdx.x,
               (int)threadIdx.x);
        __trap();
      }
    }
  }
}
// Bulk copy of `bytes` from this CTA's shared memory into a peer CTA's (addresses in the shared::cluster window); the
// bytes are counted on the PEER's mbarrier (complete_tx), which is what its cell phase waits on.
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster_addr, uint32_t src_cta_addr, uint32_t bytes,
                                                  uint32_t mbar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster_addr), "r"(src_cta_addr), "r"(bytes), "r"(mbar_cluster_addr)
               : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }
__device__ __forceinline__ void ld8(const float* p, float* r) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
// (plain loads: data written earlier in this kernel by other CTAs must not go through the non-coherent path)
__device__ __forceinline__ void ld8_coherent(const float* p, float* r) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
  r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float* r) {
  reinterpret_cast<float4*>(p)[0] = make_float4(r[0], r[1], r[2], r[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(r[4], r[5], r[6], r[7]);
}
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float* r) {
  uint4 v;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(r[0], r[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r[2], r[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r[4], r[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(r[6], r[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
  *reinterpret_cast<uint4*>(p) = v;
}

// BWD = 0: forward recurrence; BWD = 1: backward (BPTT) recurrence.
template <int BWD>
__global__ void __launch_bounds__(NTHREADS, 1)
lstm_persist_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                    const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + SMEM_W;
  float* sRecv = reinterpret_cast<float*>(sA + SMEM_RING);
  float* sStage = reinterpret_cast<float*>(sA + SMEM_RING + SMEM_RECV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sRecv) + SMEM_RECV + SMEM_STAGE);
  uint64_t* full = bars;            // [NST]
  uint64_t* empty = bars + NST;     // [NST]
  uint64_t* w_full = bars + 2 * NST;
  uint64_t* acc_full = w_full + 1;
  uint64_t* acc_empty = w_full + 2;
  uint64_t* xchg = w_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = cluster_ctarank();
  const int cl = blockIdx.x >> 2;
  const int bt = cl / 16, J = cl % 16;
  const int T = p.T, B = p.B, nbt = p.nbt;
  // Clusters walk the chunk groups in rotated order, so that the 16 clusters of a batch tile do not all request the
  // same L2 lines at the same moment. The order is the same for the four CTAs of a cluster (forward: they share the
  // operand through multicast).
  const int rot = J & (NG - 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BWD ? 1 : 4);  // forward: freed by the MMAs of all four CTAs (multicast commit)
    }
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    mbar_init(xchg, 1);  // armed once per step with expect_tx of the three remote pieces
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_w);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();  // the peers' barriers exist before anything arrives on them remotely
  const uint32_t tmem = *tmem_slot;

  // step s = 0 .. T-1 handles time t = s (forward) or t = T-1-s (backward). The backward's first step has no
  // recurrent product unless a final-state gradient is given (it is added in the cell phase directly).
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(w_full, SMEM_W);
      for (int kc = 0; kc < CHUNKS; ++kc) {
        if (BWD) tma_load_2d(&tmap_w, w_full, sW + kc * CHUNK_BYTES, (int)q * 1024 + 64 * kc, 64 * J);
        else tma_load_2d(&tmap_w, w_full, sW + kc * CHUNK_BYTES, 64 * kc, (int)q * 1024 + 64 * J);
      }
    }
    __syncwarp();
    int it = 0;
    if (!BWD) {
      // Forward: the four CTAs of a cluster need the SAME operand rows (all 16 chunks of their batch tile). Each CTA
      // loads one group per step — the k-th group of the step is issued by CTA k — and multicasts it to all four, so
      // the L2 sees a quarter of the requests. A stage is reused when the MMAs of all four CTAs have released it
      // (count-4 `empty` barrier, multicast commit); every CTA arms its own `full` barrier.
      static_assert(NG == 4, "one group per CTA of the cluster");
      for (int s = 0; s < T; ++s) {
        const unsigned int* flags = p.ready + ((size_t)s * nbt + bt) * 16;
        const int a_row = s * B + 64 * bt;
        for (int k = 0; k < NG; ++k, ++it) {
          const int st = it % NST;
          const uint32_t ph = (uint32_t)(it / NST) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          if (elect_one()) mbar_expect_tx(&full[st], STAGE_BYTES);
          __syncwarp();
          if (k != (int)q) continue;
          const int g = (k + rot) & (NG - 1);
          if (s > 0) {  // the group's four chunks must have been written by their 16 producer CTAs
            uint32_t polls = 0;
            uint64_t t0 = 0;
            for (;;) {
              unsigned int v = 4;
              if (lane < CPI) v = ld_acquire_gpu(flags + CPI * g + lane);
              if (__all_sync(0xffffffffu, v >= 4u)) break;
              if ((++polls & 63u) == 0) {
                const uint64_t now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 4000000000ull) {
                  if (lane == 0)
                    printf("pvr: lstm_persist forward step %d group %d never became ready (block %d)\n", s, g,
                           (int)blockIdx.x);
                  __trap();
                }
              }
            }
          }
          if (lane == 0 && k == 0) PROF(0);
          fence_proxy_async_global();  // rows written with st.global by other SMs; TMA reads them next
          if (elect_one())
            tma_load_3d_multicast(&tmap_a, &full[st], sA + st * STAGE_BYTES, 0, a_row, CPI * g, (uint16_t)0xF);
          __syncwarp();
          if (lane == 0) PROF(2);
        }
      }
    }
    for (int s = 1; BWD && s < T; ++s) {
      const int t = T - 1 - s;
      const int src_t = t + 1;  // time index of the operand rows (dG_{t+1})
      const bool gated = true;
      const unsigned int* flags = p.ready + ((size_t)src_t * nbt + bt) * 16;
      const int a_row = src_t * B + 64 * bt;
      const int a_c0 = (int)q * 16;  // first K chunk of this CTA's slice (dG: gate q's 1024 columns)
      int next = 0;  // next chunk GROUP (CPI chunks = one TMA instruction = one ring stage), in this CTA's order
      uint32_t polls = 0;
      uint64_t t0 = 0;
      while (next < NG) {
        unsigned int mask = 0xffffu;
        if (gated) {
          unsigned int v = 4;
          if (lane < 16) v = ld_acquire_gpu(flags + lane);
          mask = __ballot_sync(0xffffffffu, v >= 4u) & 0xffffu;
        }
        constexpr unsigned int GM = (1u << CPI) - 1u;
        int upto = next;
        while (upto < NG && ((mask >> (CPI * ((upto + rot) & (NG - 1)))) & GM) == GM) ++upto;
        if (upto == next) {
          if ((++polls & 63u) == 0) {
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) {
              if (lane == 0)
                printf("pvr: lstm_persist step %d group %d never became ready (block %d, mask %x)\n", s, next,
                       (int)blockIdx.x, mask);
              __trap();
            }
          }
          continue;
        }
        if (lane == 0 && next == 0) PROF(0);               // first group seen ready
        if (lane == 0 && upto == NG) PROF(1);              // all chunks seen ready
        fence_proxy_async_global();  // rows written with st.global by other SMs; TMA reads them next
        for (; next < upto; ++next, ++it) {
          const int st = it % NST;
          const uint32_t ph = (uint32_t)(it / NST) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&full[st], STAGE_BYTES);
            tma_load_3d(&tmap_a, &full[st], sA + st * STAGE_BYTES, 0, a_row, a_c0 + CPI * ((next + rot) & (NG - 1)));
          }
          __syncwarp();
        }
      }
      if (lane == 0) PROF(2);                              // last TMA of the step issued
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    mbar_wait(w_full, 0);
    constexpr uint32_t idesc = umma_idesc_bf16(64, 64);
    int it = 0, gi = 0;
    for (int s = BWD ? 1 : 0; s < T; ++s, ++gi) {
      mbar_wait(acc_empty, ((uint32_t)gi & 1u) ^ 1u);  // the previous step's accumulator has been read
      tc_fence_after();
      for (int g = 0; g < NG; ++g, ++it) {
        const int st = it % NST;
        const uint32_t ph = (uint32_t)(it / NST) & 1u;
        mbar_wait(&full[st], ph);
        tc_fence_after();
        if (lane == 0 && g == 0) PROF(3);                  // first group landed
        if (lane == 0 && g == NG - 1) PROF(4);             // last group landed
        if (lane == 0 && g == NG / 4) PROF(12);
        if (lane == 0 && g == NG / 2) PROF(13);
        if (lane == 0 && g == 3 * NG / 4) PROF(14);
        if (elect_one()) {
          const int kc0 = CPI * ((g + rot) & (NG - 1));
#pragma unroll
          for (int c = 0; c < CPI; ++c) {
            const uint64_t ad = umma_desc_sw128(smem_u32(sA + st * STAGE_BYTES + c * CHUNK_BYTES));
            const uint64_t bd = umma_desc_sw128(smem_u32(sW + (kc0 + c) * CHUNK_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, ((g * CPI + c) | k) != 0 ? 1u : 0u);
          }
          if (BWD) umma_commit(&empty[st]);
          else umma_commit_multicast(&empty[st], (uint16_t)0xF);  // the stage is shared by the cluster's loads
          if (g == NG - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ exchange + cell (128 threads)
    const int e = threadIdx.x - 64;
    const int ew = warp & 3;            // TMEM lane quadrant this warp may read: rows 16 ew .. 16 ew + 15 in lanes 0-15
    const int erow = 16 * ew + lane;    // accumulator row held by this lane (lane < 16)
    const int crow = e & 63, half = e >> 6;
    const int b = 64 * bt + crow;
    const bool row_ok = b < B;
    const int j0 = 64 * J + 16 * (int)q + 8 * half;  // first of this thread's 8 hidden units
    const uint32_t recv_base = smem_u32(sRecv);
    const uint32_t xchg_addr = smem_u32(xchg);
    float state[8];  // forward: c_{t-1}; backward: dc flowing from step t+1
#pragma unroll
    for (int u = 0; u < 8; ++u) state[u] = 0.f;
    float dh_last[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) dh_last[u] = 0.f;
    if (row_ok) {
      if (!BWD) ld8_coherent(p.c_all + (size_t)b * HID + j0, state);
      else {
        ld8_coherent(p.dc_rec + (size_t)b * HID + j0, state);
        ld8_coherent(p.dh_rec + (size_t)b * HID + j0, dh_last);
      }
    }
    float bsum[BWD ? 32 : 1];  // backward: this thread's share of the bias gradient, summed over the time steps
#pragma unroll
    for (int i = 0; i < (BWD ? 32 : 1); ++i) bsum[i] = 0.f;
    uint32_t xphase = 0;
    int gi = 0;
    for (int s = 0; s < T; ++s) {
      const int t = BWD ? T - 1 - s : s;
      const bool has_gemm = BWD ? (s > 0) : true;
      const size_t row_t = (size_t)t * B + b;
      // ---- operands of the cell phase that do not depend on the recurrence: issued before the waits
      float op[32];
      float cp[8], cc[8], dho[8];
      float nd_t = 0.f, nd_n = 0.f;
      if (row_ok) {
        nd_t = __ldg(p.nd + row_t);
        if (t + 1 < T) nd_n = __ldg(p.nd + row_t + B);
        if (!BWD) {
#pragma unroll
          for (int g = 0; g < 4; ++g) ld8(p.xp + row_t * (4 * HID) + g * HID + j0, op + 8 * g);
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g) ld8(p.gates + row_t * (4 * HID) + g * HID + j0, op + 8 * g);
          ld8(p.c_all + row_t * HID + j0, cp);
          ld8(p.c_all + (row_t + B) * HID + j0, cc);
          if (p.dh_out) ld8(p.dh_out + row_t * HID + j0, dho);
          else {
#pragma unroll
            for (int u = 0; u < 8; ++u) dho[u] = 0.f;
          }
        }
      }
      float G[32];
      if (has_gemm) {
        // ---- accumulator -> 16-column pieces to the four CTAs of the cluster
        mbar_wait(acc_full, (uint32_t)gi & 1u);
        tc_fence_after();
        if (e == 0) PROF(5);                               // accumulator complete
        uint32_t v[64];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * ew) << 16), v);
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * ew) << 16) + 32, v + 32);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        if (lane < 16) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            // the piece this CTA keeps goes straight into its own receive buffer, the others are staged for the copies
            float* dstp = (r == (int)q ? sRecv + ((int)q * 64 + erow) * 16 : sStage + (r * 64 + erow) * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              reinterpret_cast<float4*>(dstp)[i] =
                  make_float4(__uint_as_float(v[16 * r + 4 * i]), __uint_as_float(v[16 * r + 4 * i + 1]),
                              __uint_as_float(v[16 * r + 4 * i + 2]), __uint_as_float(v[16 * r + 4 * i + 3]));
          }
          fence_proxy_async();  // generic-proxy writes to shared memory -> visible to the bulk-copy engine
        }
        named_bar_sync(2, 128);
        if (e == 0) {
          mbar_expect_tx(xchg, 3 * PIECE_BYTES);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r != (int)q)
              bulk_copy_to_peer(mapa(recv_base + (uint32_t)q * PIECE_BYTES, (uint32_t)r),
                                smem_u32(sStage) + (uint32_t)r * PIECE_BYTES, PIECE_BYTES, mapa(xchg_addr, (uint32_t)r));
          PROF(6);                                         // pieces sent
        }
        mbar_wait_cluster(xchg, xphase);
        if (e == 0) PROF(7);                               // pieces of all four CTAs received
        xphase ^= 1u;
        ++gi;
#pragma unroll
        for (int src = 0; src < 4; ++src) {
          const float4* rp = reinterpret_cast<const float4*>(sRecv + (src * 64 + crow) * 16 + 8 * half);
          const float4 a = rp[0], c4 = rp[1];
          G[8 * src + 0] = a.x; G[8 * src + 1] = a.y; G[8 * src + 2] = a.z; G[8 * src + 3] = a.w;
          G[8 * src + 4] = c4.x; G[8 * src + 5] = c4.y; G[8 * src + 6] = c4.z; G[8 * src + 7] = c4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) G[i] = 0.f;
      }
      // ---- cell. Order of the stores: (1) the operand of the next step (masked h_t / dG_t), (2) its arrival counter,
      // (3) whatever only the backward pass / the caller reads (gates, c_t, h_t) — off the critical path.
      const bool publish = BWD ? (t > 0) : (t + 1 < T);
      float o0[8], o1[8], o2[8], o3[8], hv[8];
      if (row_ok) {
        if (!BWD) {
          float hmv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float i = sigmoid_fast(G[u] + op[u]);
            const float f = sigmoid_fast(G[8 + u] + op[8 + u]);
            const float g = tanh_fast(G[16 + u] + op[16 + u]);
            const float o = sigmoid_fast(G[24 + u] + op[24 + u]);
            const float c = f * (nd_t * state[u]) + i * g;
            const float h = o * tanh_fast(c);
            state[u] = c;
            o0[u] = i; o1[u] = f; o2[u] = g; o3[u] = o;
            hv[u] = h;
            hmv[u] = h * nd_n;
          }
          if (t + 1 < T) st8_bf16(p.hm + (row_t + B) * HID + j0, hmv);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            // dh = gradient from above + what flows back from step t+1 through W_hh (masked by done[t+1]); at the
            // last step the gradient of the returned state instead (zero in BC training)
            const float rec = has_gemm ? nd_n * ((G[u] + G[8 + u]) + (G[16 + u] + G[24 + u])) : dh_last[u];
            const float dh = dho[u] + rec;
            const float i = op[u], f = op[8 + u], g = op[16 + u], o = op[24 + u];
            const float tc = tanh_fast(cc[u]);
            const float dc = state[u] + dh * o * (1.f - tc * tc);
            const float cpm = nd_t * cp[u];
            o0[u] = dc * g * i * (1.f - i);
            o1[u] = dc * cpm * f * (1.f - f);
            o2[u] = dc * i * (1.f - g * g);
            o3[u] = dh * tc * o * (1.f - o);
            state[u] = dc * f * nd_t;
            if (BWD) {
              bsum[u] += o0[u]; bsum[8 + u] += o1[u]; bsum[16 + u] += o2[u]; bsum[24 + u] += o3[u];
            }
          }
          __nv_bfloat16* dp = p.dG + row_t * (4 * HID) + j0;
          st8_bf16(dp, o0);
          st8_bf16(dp + HID, o1);
          st8_bf16(dp + 2 * HID, o2);
          st8_bf16(dp + 3 * HID, o3);
        }
      }
      if (e == 0) PROF(8);                                 // cell math done, operand stores issued
      if (publish) {
        fence_proxy_async_global();
        named_bar_sync(1, 128);
        if (e == 0) {
          PROF(9);
          const int dst_t = BWD ? t : t + 1;
          // release at gpu scope: cumulative over the stores of the 127 other threads ordered by the barrier
          unsigned int* flag = p.ready + ((size_t)dst_t * nbt + bt) * 16 + J;
          red_release_gpu_add(flag, 1u);
          PROF(11);
        }
        // the remaining stores wait for the release: a gpu-scope fence drains every store the SM has in flight, the
        // 28 KB below included if they were already issued
        named_bar_sync(3, 128);
      }
      if (!BWD && row_ok) {
        float* gp = p.gates + row_t * (4 * HID) + j0;
        st8(gp, o0);
        st8(gp + HID, o1);
        st8(gp + 2 * HID, o2);
        st8(gp + 3 * HID, o3);
        st8(p.c_all + (row_t + B) * HID + j0, state);
        st8_bf16(p.h_out + row_t * HID + j0, hv);
        if (t + 1 >= T) st8(p.h_last + (size_t)b * HID + j0, hv);
      }
    }
    if (BWD && p.dbias) {
      // bias gradient: the 32 rows a warp holds share `half`, so the rows are summed by shuffles first
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = bsum[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) atomicAdd(p.dbias + (i >> 3) * HID + j0 + (i & 7), v);
      }
    }
    if (BWD && row_ok) {
      // gradient w.r.t. the initial state: dc is complete; dh_0 needs one more product with W_hh, which training never
      // uses (main_bc_2.py:207 passes a fresh zero state every step) — the cell part is returned, dh_rec is zeroed.
      st8(p.dc_rec + (size_t)b * HID + j0, state);
      float z[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) z[u] = 0.f;
      st8(p.dh_rec + (size_t)b * HID + j0, z);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still write into its shared memory
  if (warp == 1) tmem_dealloc(tmem, 64);
}

// ---------------------------------------------------------------------------------------------------------------
// Forward recurrence, second form: the W_hh slice lives in TENSOR MEMORY.
//
// The kernel above is bound by what an SM can pull in per step (128 KB of h_{t-1} per CTA at ~32 B/clk, plus the
// exchange of accumulator pieces through DSMEM). Here the roles of the operands are swapped: a CTA owns 32 hidden
// units x 4 gates = 128 rows of W_hh as the A operand of an M = 128 MMA, held in TMEM as packed bf16 (128 lanes x 480
// columns = K 0..959; the last K chunk sits in shared memory — TMEM has 512 columns and the accumulator needs 32), and
// the B operand is a batch tile of 32 rows of h_{t-1} (64 KB per step: half the ingest, all four gates of a unit in
// the same CTA, so no exchange and no cluster). TMEM lane r = 4 u + g holds gate g of unit 32 c + u. 32 CTAs per batch
// tile x ceil(B / 32) tiles = 128 CTAs at B = 128. After the MMAs the (128 x 32) accumulator is transposed through
// shared memory so that a thread owns (batch row, 8 units, 4 gates) — the cell code and every saved tensor are the
// ones of the kernel above.
constexpr int F2_NB = 32;                            // batch rows per tile = MMA N
#ifndef PVR_F2_CPI
#define PVR_F2_CPI 8
#endif
constexpr int F2_CPI = PVR_F2_CPI;                   // K chunks per TMA instruction
constexpr int F2_NG = CHUNKS / F2_CPI;
constexpr int F2_CHUNK_BYTES = F2_NB * 128;          // 4096
constexpr int F2_SMEM_B = CHUNKS * F2_CHUNK_BYTES;   // 65536: one step's operand, every group its own slot
constexpr int F2_KT = 15;                            // K chunks of the W slice held in TMEM (columns 32 .. 511)
constexpr int F2_SMEM_WT = 128 * 128;                // the 16th chunk: (128 rows x 64) bf16, 128B-swizzled
constexpr int F2_XPITCH = 33;
constexpr int F2_SMEM_X = 128 * F2_XPITCH * 4;       // accumulator transpose
constexpr int F2_SMEM_TOTAL = F2_SMEM_B + F2_SMEM_WT + F2_SMEM_X + 256 + 1024;

__global__ void __launch_bounds__(NTHREADS, 1)
lstm_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __nv_bfloat16* __restrict__ w_hh,
                 const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sWt = smem + F2_SMEM_B;
  float* sX = reinterpret_cast<float*>(sWt + F2_SMEM_WT);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + F2_SMEM_X);
  uint64_t* full = bars;             // [F2_NG]
  uint64_t* empty = bars + F2_NG;    // [F2_NG]
  uint64_t* acc_full = bars + 2 * F2_NG;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bt = blockIdx.x >> 5, c = blockIdx.x & 31;
  const int T = p.T, B = p.B, nbt = p.nbt;
  const int rot = c & (F2_NG - 1);  // the CTAs of a tile walk the groups in rotated order (spreads the L2 requests)

  if (threadIdx.x == 0) {
    for (int g = 0; g < F2_NG; ++g) {
      mbar_init(&full[g], 1);
      mbar_init(&empty[g], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= 2) {
    // ---- W slice -> TMEM (packed bf16 pairs, the A-operand layout of tcgen05.mma with A in tensor memory) + the
    // last chunk -> shared memory. Row r of the slice is W_hh row (r & 3) * 1024 + 32 c + (r >> 2).
    const int ew = warp & 3;
    const int r = 32 * ew + lane;
    const __nv_bfloat16* wrow = w_hh + ((size_t)(r & 3) * HID + 32 * c + (r >> 2)) * HID;
    const uint32_t ta = tmem + ((uint32_t)(32 * ew) << 16) + 32;
#pragma unroll 6
    for (int ks = 0; ks < F2_KT * 4; ++ks) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(wrow + 16 * ks));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(wrow + 16 * ks) + 1);
      const uint32_t pk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      tmem_st_32x32b_x8(ta + 8 * ks, pk);
    }
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + 64 * F2_KT) + q8);
      *reinterpret_cast<uint4*>(sWt + r * 128 + ((q8 ^ (r & 7)) << 4)) = v;
    }
    tmem_wait_st();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int s = 0; s < T; ++s) {
      const unsigned int* flags = p.ready + ((size_t)s * nbt + bt) * 16;
      const int a_row = s * B + F2_NB * bt;
      for (int k = 0; k < F2_NG; ++k) {
        const int g = (k + rot) & (F2_NG - 1);
        mbar_wait(&empty[g], ((uint32_t)s & 1u) ^ 1u);
        // the group's chunks must have been written by their producers (2 CTAs per chunk of 64 units)
        if (s > 0 && !wait_flags(flags + F2_CPI * g, F2_CPI, 2u, lane)) {
          if (lane == 0) printf("pvr: lstm_fwd2 step %d group %d never became ready (block %d)\n", s, g, (int)blockIdx.x);
          __trap();
        }
        if (lane == 0 && k == 0) PROF(0);
        fence_proxy_async_global();  // rows written with st.global by other SMs; TMA reads them next
        if (elect_one()) {
          mbar_expect_tx(&full[g], F2_CPI * F2_CHUNK_BYTES);
          tma_load_3d(&tmap_a, &full[g], sB + g * F2_CPI * F2_CHUNK_BYTES, 0, a_row, F2_CPI * g);
        }
        __syncwarp();
        if (lane == 0 && k == F2_NG - 1) PROF(2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, F2_NB);
    const uint64_t wt_desc = umma_desc_sw128(smem_u32(sWt));
    for (int s = 0; s < T; ++s) {
      mbar_wait(acc_empty, ((uint32_t)s & 1u) ^ 1u);  // the previous step's accumulator has been read
      tc_fence_after();
      for (int k = 0; k < F2_NG; ++k) {
        const int g = (k + rot) & (F2_NG - 1);
        mbar_wait(&full[g], (uint32_t)s & 1u);
        tc_fence_after();
        if (lane == 0 && k == 0) PROF(3);
        if (lane == 0 && k == F2_NG - 1) PROF(4);
        if (elect_one()) {
#pragma unroll
          for (int cc = 0; cc < F2_CPI; ++cc) {
            const int kc = F2_CPI * g + cc;
            const uint64_t bd = umma_desc_sw128(smem_u32(sB + kc * F2_CHUNK_BYTES));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t acc = (k | cc | kk) != 0 ? 1u : 0u;
              if (kc < F2_KT) umma_bf16_ts(tmem, tmem + 32 + 8 * (4 * kc + kk), bd + 2 * kk, idesc, acc);
              else umma_bf16(tmem, wt_desc + 2 * kk, bd + 2 * kk, idesc, acc);
            }
          }
          umma_commit(&empty[g]);
          if (k == F2_NG - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ cell (128 threads)
    const int e = threadIdx.x - 64;
    const int ew = warp & 3;            // TMEM lane quadrant of this warp: rows 32 ew .. 32 ew + 31 of the slice
    const int crow = e >> 2, ug = e & 3;
    const int b = F2_NB * bt + crow;
    const bool row_ok = b < B;
    const int j0 = 32 * c + 8 * ug;     // first of this thread's 8 hidden units
    float state[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) state[u] = 0.f;
    if (row_ok) ld8_coherent(p.c_all + (size_t)b * HID + j0, state);
    for (int s = 0; s < T; ++s) {
      const int t = s;
      const size_t row_t = (size_t)t * B + b;
      float op[32];
      float nd_t = 0.f, nd_n = 0.f;
      if (row_ok) {
        nd_t = __ldg(p.nd + row_t);
        if (t + 1 < T) nd_n = __ldg(p.nd + row_t + B);
#pragma unroll
        for (int g = 0; g < 4; ++g) ld8(p.xp + row_t * (4 * HID) + g * HID + j0, op + 8 * g);
      }
      mbar_wait(acc_full, (uint32_t)s & 1u);
      tc_fence_after();
      if (e == 0) PROF(5);
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * ew) << 16), v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
        // row r = 32 ew + lane, column n (batch row of the tile) -> sX[r][(n + 8 ew) & 31]: conflict free both ways
        float* xr = sX + (32 * ew + lane) * F2_XPITCH;
#pragma unroll
        for (int n = 0; n < 32; ++n) xr[(n + 8 * ew) & 31] = __uint_as_float(v[n]);
      }
      named_bar_sync(2, 128);
      float G[32];
      {
        const float* xc = sX + (32 * ug) * F2_XPITCH + ((crow + 8 * ug) & 31);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int g = 0; g < 4; ++g) G[8 * g + u] = xc[(4 * u + g) * F2_XPITCH];
      }
      const bool publish = t + 1 < T;
      float o0[8], o1[8], o2[8], o3[8], hv[8];
      if (row_ok) {
        float hmv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float i = sigmoid_fast(G[u] + op[u]);
          const float f = sigmoid_fast(G[8 + u] + op[8 + u]);
          const float g = tanh_fast(G[16 + u] + op[16 + u]);
          const float o = sigmoid_fast(G[24 + u] + op[24 + u]);
          const float cv = f * (nd_t * state[u]) + i * g;
          const float h = o * tanh_fast(cv);
          state[u] = cv;
          o0[u] = i; o1[u] = f; o2[u] = g; o3[u] = o;
          hv[u] = h;
          hmv[u] = h * nd_n;
        }
        if (publish) st8_bf16(p.hm + (row_t + B) * HID + j0, hmv);
      }
      if (e == 0) PROF(8);
      if (publish) {
        fence_proxy_async_global();
        named_bar_sync(1, 128);
        if (e == 0) {
          PROF(9);
          red_release_gpu_add(p.ready + ((size_t)(t + 1) * nbt + bt) * 16 + (c >> 1), 1u);
          PROF(11);
        }
        named_bar_sync(3, 128);  // the stores below must not be in flight when the release drains the SM's stores
      }
      if (row_ok) {
        float* gp = p.gates + row_t * (4 * HID) + j0;
        st8(gp, o0);
        st8(gp + HID, o1);
        st8(gp + 2 * HID, o2);
        st8(gp + 3 * HID, o3);
        st8(p.c_all + (row_t + B) * HID + j0, state);
        st8_bf16(p.h_out + row_t * HID + j0, hv);
        if (t + 1 >= T) st8(p.h_last + (size_t)b * HID + j0, hv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// Backward recurrence, second form: dh_{t}[b, j] = sum_k dG_{t+1}[b, k] W_hh[k, j] with W_hh^T in tensor memory.
//
// CTA (J, q) of a cluster of 4 owns hidden units 128 J .. + 128 as the M = 128 rows of the A operand and gate q's 1024
// columns of dG as its K slice (960 in TMEM + 64 in shared memory, as in the forward kernel); the B operand is a batch
// tile of 32 rows of dG_{t+1} (64 KB per step instead of 128). The four partial (128 x 32) sums of a cluster are
// reduce-scattered through distributed shared memory: TMEM lane quadrant w (units 32 w .. + 32) goes to CTA w, which
// adds the four pieces and runs the cell update for those 32 units x 32 batch rows. 8 J x 4 q = 32 CTAs per batch tile.
constexpr int B2_PITCH = 33;
constexpr int B2_PIECE_BYTES = 32 * B2_PITCH * 4;    // (32 units x 32 batch rows) fp32, padded rows: 4224
constexpr int B2_SMEM_RECV = 4 * B2_PIECE_BYTES;     // [source CTA][unit][batch row]
constexpr int B2_SMEM_STAGE = 4 * B2_PIECE_BYTES;    // [destination CTA][unit][batch row]
constexpr int B2_SMEM_TOTAL = F2_SMEM_B + F2_SMEM_WT + B2_SMEM_RECV + B2_SMEM_STAGE + 256 + 1024;
static_assert(B2_PIECE_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

__global__ void __launch_bounds__(NTHREADS, 1)
lstm_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __nv_bfloat16* __restrict__ w_hh_t,
                 const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sWt = smem + F2_SMEM_B;
  float* sRecv = reinterpret_cast<float*>(sWt + F2_SMEM_WT);
  float* sStage = reinterpret_cast<float*>(sWt + F2_SMEM_WT + B2_SMEM_RECV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sWt + F2_SMEM_WT + B2_SMEM_RECV + B2_SMEM_STAGE);
  uint64_t* full = bars;             // [F2_NG]
  uint64_t* empty = bars + F2_NG;    // [F2_NG]
  uint64_t* acc_full = bars + 2 * F2_NG;
  uint64_t* acc_empty = acc_full + 1;
  uint64_t* xchg = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = cluster_ctarank();
  const int cl = blockIdx.x >> 2;
  const int bt = cl >> 3, J = cl & 7;
  const int T = p.T, B = p.B, nbt = p.nbt;
  const int rot = J & (F2_NG - 1);

  if (threadIdx.x == 0) {
    for (int g = 0; g < F2_NG; ++g) {
      mbar_init(&full[g], 1);
      mbar_init(&empty[g], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    mbar_init(xchg, 1);
    fence_barrier_init();
    prefetch_tmap(&tmap_a);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp >= 2) {
    // row r of the slice = row 128 J + r of W_hh^T (4096 columns), columns 1024 q .. + 1024
    const int ew = warp & 3;
    const int r = 32 * ew + lane;
    const __nv_bfloat16* wrow = w_hh_t + (size_t)(128 * J + r) * (4 * HID) + 1024 * (int)q;
    const uint32_t ta = tmem + ((uint32_t)(32 * ew) << 16) + 32;
#pragma unroll 6
    for (int ks = 0; ks < F2_KT * 4; ++ks) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(wrow + 16 * ks));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(wrow + 16 * ks) + 1);
      const uint32_t pk[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      tmem_st_32x32b_x8(ta + 8 * ks, pk);
    }
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(wrow + 64 * F2_KT) + q8);
      *reinterpret_cast<uint4*>(sWt + r * 128 + ((q8 ^ (r & 7)) << 4)) = v;
    }
    tmem_wait_st();
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();  // the peers' barriers exist before anything arrives on them remotely

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int s = 1; s < T; ++s) {
      const int src_t = T - s;  // time index of the operand rows (dG_{t+1}, t = T - 1 - s)
      const unsigned int* flags = p.ready + ((size_t)src_t * nbt + bt) * 16;
      const int a_row = src_t * B + F2_NB * bt;
      for (int k = 0; k < F2_NG; ++k) {
        const int g = (k + rot) & (F2_NG - 1);
        mbar_wait(&empty[g], ((uint32_t)s & 1u));  // step s = 1 is the first use (parity 1 passes at once)
        if (!wait_flags(flags + F2_CPI * g, F2_CPI, 2u, lane)) {
          if (lane == 0) printf("pvr: lstm_bwd2 step %d group %d never became ready (block %d)\n", s, g, (int)blockIdx.x);
          __trap();
        }
        if (lane == 0 && k == 0) PROF(0);
        fence_proxy_async_global();
        if (elect_one()) {
          mbar_expect_tx(&full[g], F2_CPI * F2_CHUNK_BYTES);
          tma_load_3d(&tmap_a, &full[g], sB + g * F2_CPI * F2_CHUNK_BYTES, 0, a_row, 16 * (int)q + F2_CPI * g);
        }
        __syncwarp();
        if (lane == 0 && k == F2_NG - 1) PROF(2);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(128, F2_NB);
    const uint64_t wt_desc = umma_desc_sw128(smem_u32(sWt));
    for (int s = 1; s < T; ++s) {
      mbar_wait(acc_empty, (uint32_t)s & 1u);  // the previous step's accumulator has been read
      tc_fence_after();
      for (int k = 0; k < F2_NG; ++k) {
        const int g = (k + rot) & (F2_NG - 1);
        mbar_wait(&full[g], ((uint32_t)s & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0 && k == 0) PROF(3);
        if (lane == 0 && k == F2_NG - 1) PROF(4);
        if (elect_one()) {
#pragma unroll
          for (int cc = 0; cc < F2_CPI; ++cc) {
            const int kc = F2_CPI * g + cc;
            const uint64_t bd = umma_desc_sw128(smem_u32(sB + kc * F2_CHUNK_BYTES));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint32_t acc = (k | cc | kk) != 0 ? 1u : 0u;
              if (kc < F2_KT) umma_bf16_ts(tmem, tmem + 32 + 8 * (4 * kc + kk), bd + 2 * kk, idesc, acc);
              else umma_bf16(tmem, wt_desc + 2 * kk, bd + 2 * kk, idesc, acc);
            }
          }
          umma_commit(&empty[g]);
          if (k == F2_NG - 1) umma_commit(acc_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ exchange + cell (128 threads)
    const int e = threadIdx.x - 64;
    const int ew = warp & 3;            // TMEM lane quadrant of this warp = the cluster rank its piece goes to
    const int crow = e >> 2, ug = e & 3;
    const int b = F2_NB * bt + crow;
    const bool row_ok = b < B;
    const int j0 = 128 * J + 32 * (int)q + 8 * ug;  // first of this thread's 8 hidden units
    const uint32_t recv_base = smem_u32(sRecv);
    const uint32_t xchg_addr = smem_u32(xchg);
    float state[8], dh_last[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) state[u] = dh_last[u] = 0.f;
    if (row_ok) {
      ld8_coherent(p.dc_rec + (size_t)b * HID + j0, state);
      ld8_coherent(p.dh_rec + (size_t)b * HID + j0, dh_last);
    }
    float bsum[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bsum[i] = 0.f;
    uint32_t xphase = 0;
    for (int s = 0; s < T; ++s) {
      const int t = T - 1 - s;
      const bool has_gemm = s > 0;
      const size_t row_t = (size_t)t * B + b;
      float op[32], cp[8], cc[8], dho[8];
      float nd_t = 0.f, nd_n = 0.f;
      if (row_ok) {
        nd_t = __ldg(p.nd + row_t);
        if (t + 1 < T) nd_n = __ldg(p.nd + row_t + B);
#pragma unroll
        for (int g = 0; g < 4; ++g) ld8(p.gates + row_t * (4 * HID) + g * HID + j0, op + 8 * g);
        ld8(p.c_all + row_t * HID + j0, cp);
        ld8(p.c_all + (row_t + B) * HID + j0, cc);
        if (p.dh_out) ld8(p.dh_out + row_t * HID + j0, dho);
        else {
#pragma unroll
          for (int u = 0; u < 8; ++u) dho[u] = 0.f;
        }
      }
      float G[32];
      if (has_gemm) {
        mbar_wait(acc_full, ((uint32_t)s & 1u) ^ 1u);
        tc_fence_after();
        if (e == 0) PROF(5);
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem + ((uint32_t)(32 * ew) << 16), v);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty);
          // this warp's (32 units x 32 batch rows) piece belongs to CTA ew: straight into the own receive buffer, or
          // staged for the copy
          float* dstp = (ew == (int)q ? sRecv + (int)q * (32 * B2_PITCH) : sStage + ew * (32 * B2_PITCH)) +
                        lane * B2_PITCH;
#pragma unroll
          for (int n = 0; n < 32; ++n) dstp[n] = __uint_as_float(v[n]);
        }
        fence_proxy_async();  // generic-proxy writes to shared memory -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0 && ew != (int)q)
          bulk_copy_to_peer(mapa(recv_base + q * B2_PIECE_BYTES, (uint32_t)ew),
                            smem_u32(sStage) + (uint32_t)ew * B2_PIECE_BYTES, B2_PIECE_BYTES,
                            mapa(xchg_addr, (uint32_t)ew));
        if (e == 0) {
          mbar_expect_tx(xchg, 3 * B2_PIECE_BYTES);
          PROF(6);
        }
        named_bar_sync(2, 128);  // the own piece is in sRecv (written by warp quadrant q)
        mbar_wait_cluster(xchg, xphase);
        if (e == 0) PROF(7);
        xphase ^= 1u;
#pragma unroll
        for (int src = 0; src < 4; ++src) {
          const float* rp = sRecv + src * (32 * B2_PITCH) + (8 * ug) * B2_PITCH + crow;
#pragma unroll
          for (int u = 0; u < 8; ++u) G[8 * src + u] = rp[u * B2_PITCH];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) G[i] = 0.f;
      }
      const bool publish = t > 0;
      if (row_ok) {
        float o0[8], o1[8], o2[8], o3[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float rec = has_gemm ? nd_n * ((G[u] + G[8 + u]) + (G[16 + u] + G[24 + u])) : dh_last[u];
          const float dh = dho[u] + rec;
          const float i = op[u], f = op[8 + u], g = op[16 + u], o = op[24 + u];
          const float tc = tanh_fast(cc[u]);
          const float dc = state[u] + dh * o * (1.f - tc * tc);
          const float cpm = nd_t * cp[u];
          o0[u] = dc * g * i * (1.f - i);
          o1[u] = dc * cpm * f * (1.f - f);
          o2[u] = dc * i * (1.f - g * g);
          o3[u] = dh * tc * o * (1.f - o);
          state[u] = dc * f * nd_t;
          bsum[u] += o0[u]; bsum[8 + u] += o1[u]; bsum[16 + u] += o2[u]; bsum[24 + u] += o3[u];
        }
        __nv_bfloat16* dp = p.dG + row_t * (4 * HID) + j0;
        st8_bf16(dp, o0);
        st8_bf16(dp + HID, o1);
        st8_bf16(dp + 2 * HID, o2);
        st8_bf16(dp + 3 * HID, o3);
      }
      if (e == 0) PROF(8);
      if (publish) {
        fence_proxy_async_global();
        named_bar_sync(1, 128);
        if (e == 0) {
          PROF(9);
          // 64-unit chunk 2 J + (q >> 1) of time t: written by this CTA and its neighbour of the cluster
          red_release_gpu_add(p.ready + ((size_t)t * nbt + bt) * 16 + 2 * J + ((int)q >> 1), 1u);
          PROF(11);
        }
      }
    }
    if (p.dbias) {
      // bias gradient: lanes with the same (lane & 3) hold the same 8 units for 8 different batch rows
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = bsum[i];
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < 4) atomicAdd(p.dbias + (i >> 3) * HID + j0 + (i & 7), v);
      }
    }
    if (row_ok) {
      st8(p.dc_rec + (size_t)b * HID + j0, state);
      float z[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) z[u] = 0.f;
      st8(p.dh_rec + (size_t)b * HID + j0, z);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still write into its shared memory
  if (warp == 1) tmem_dealloc(tmem, 512);
}

long long* g_prof = nullptr;      // set by pvr_lstm_persist_profile()
unsigned int* g_ready = nullptr;  // process-wide arrival counters, used when the caller passes none (stream ordered)
size_t g_ready_words = 0;
int g_max_clusters[2] = {-1, -1};

template <int BWD>
int max_active_clusters(int grid) {
  if (g_max_clusters[BWD] >= 0) return g_max_clusters[BWD];
  cudaFuncSetAttribute(lstm_persist_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_persist_kernel<BWD>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  g_max_clusters[BWD] = n;
  return n;
}

int g_fwd2_blocks = -1;
int fwd2_max_blocks() {
  if (g_fwd2_blocks >= 0) return g_fwd2_blocks;
  cudaFuncSetAttribute(lstm_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_TOTAL);
  int per_sm = 0, dev = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_fwd2_kernel, NTHREADS, F2_SMEM_TOTAL) !=
      cudaSuccess) {
    cudaGetLastError();
    per_sm = 0;
  }
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // one CTA per SM: each allocates all 512 TMEM columns
  g_fwd2_blocks = per_sm > 0 ? sms : 0;
  return g_fwd2_blocks;
}

int g_bwd2_clusters = -1;
void bwd2_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int grid, cudaStream_t st) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->gridDim = dim3(grid);
  cfg->blockDim = dim3(NTHREADS);
  cfg->dynamicSmemBytes = B2_SMEM_TOTAL;
  cfg->stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}
int bwd2_max_clusters() {
  if (g_bwd2_clusters >= 0) return g_bwd2_clusters;
  cudaFuncSetAttribute(lstm_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_SMEM_TOTAL);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  bwd2_config(&cfg, attr, 128, 0);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_bwd2_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  g_bwd2_clusters = n;
  return n;
}

bool shape_ok(int T, int B, int H) { return H == HID && T >= 1 && B >= 1 && B <= 128 && T <= 4096; }

// Arrival counters: grown on first use (never inside a stream capture), zeroed on the stream before every launch.
int ensure_ready(size_t words, cudaStream_t st, void* user, int64_t user_bytes, unsigned int** out) {
  if (user && user_bytes >= (int64_t)(words * sizeof(unsigned int))) {  // the caller's own counters
    *out = static_cast<unsigned int*>(user);
    cudaMemsetAsync(user, 0, words * sizeof(unsigned int), st);
    return PVR_OK;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cap);
  if (words > g_ready_words) {
    if (cap != cudaStreamCaptureStatusNone) {
      pvr_set_error("pvr_lstm_persist: first use inside a stream capture (run one eager step first)");
      return PVR_ERR_ARG;
    }
    if (g_ready) cudaFree(g_ready);
    g_ready_words = words < 65536 ? 65536 : words;
    if (cudaMalloc(&g_ready, g_ready_words * sizeof(unsigned int)) != cudaSuccess) {
      g_ready = nullptr;
      g_ready_words = 0;
      pvr_set_error("pvr_lstm_persist: cudaMalloc of the arrival counters failed");
      return PVR_ERR_CUDA;
    }
  }
  cudaMemsetAsync(g_ready, 0, words * sizeof(unsigned int), st);
  *out = g_ready;
  return PVR_OK;
}

template <int BWD>
int launch(const CUtensorMap& ta, const CUtensorMap& tw, const PersistParams& p, cudaStream_t st, void* user,
           int64_t user_bytes) {
  const int grid = p.nbt * 16 * 4;
  unsigned int* ready = nullptr;
  const int rc = ensure_ready((size_t)(p.T + 1) * p.nbt * 16, st, user, user_bytes, &ready);
  if (rc != PVR_OK) return rc;
  PersistParams q = p;
  q.ready = ready;
  q.prof = g_prof;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_persist_kernel<BWD>, ta, tw, q);
  if (e != cudaSuccess) {
    pvr_set_error("pvr_lstm_persist launch: %s", cudaGetErrorString(e));
    return PVR_ERR_CUDA;
  }
  return PVR_OK;
}

__global__ void __launch_bounds__(256) mask_state_bf16_kernel(const float* __restrict__ h0,
                                                               const float* __restrict__ nd0, int B, int H,
                                                               __nv_bfloat16* __restrict__ hm0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * H) hm0[idx] = __float2bfloat16_rn(h0[idx] * nd0[idx / H]);
}

}  // namespace

// 1 if the persistent kernels can run this shape on the current device (all CTAs co-resident), else 0.
int lstm_persist_supported(int T, int B, int H) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("PVR_LSTM_PERSIST");
    disabled = (e && e[0] == '0') ? 1 : 0;
  }
  if (disabled || !shape_ok(T, B, H)) return 0;
  const int clusters = ((B + 63) / 64) * 16;
  return max_active_clusters<0>(clusters * 4) >= clusters && max_active_clusters<1>(clusters * 4) >= clusters;
}

// PVR_LSTM_FWD2=0 keeps the forward pass on the cluster kernel (A/B measurements).
bool use_fwd2(int B) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("PVR_LSTM_FWD2");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  return !off && ((B + F2_NB - 1) / F2_NB) * 32 <= fwd2_max_blocks();
}

int lstm_persist_forward(const pvr_lstm_fwd* L, cudaStream_t st) {
  const int T = L->T, B = L->B;
  PersistParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.B = B; p.nbt = (B + 63) / 64;
  p.nd = L->nd; p.c_all = L->c_all; p.gates = L->gates; p.xp = L->xp;
  p.hm = static_cast<__nv_bfloat16*>(L->hm);
  p.h_out = static_cast<__nv_bfloat16*>(L->h_out);
  p.h_last = L->h_last;
  CUtensorMap ta, tw;
  const char* err = nullptr;
  if (use_fwd2(B)) {
    p.nbt = (B + F2_NB - 1) / F2_NB;
    if (!make_tmap_kchunks(&ta, L->hm, HID, (uint64_t)T * B, HID, F2_NB, F2_CPI, &err)) {
      pvr_set_error("pvr_lstm_persist_forward: tensor map: %s", err ? err : "?");
      return PVR_ERR_CUDA;
    }
    const int rc = ensure_ready((size_t)(T + 1) * p.nbt * 16, st, L->counters, L->counters_bytes, &p.ready);
    if (rc != PVR_OK) return rc;
    p.prof = g_prof;
    mask_state_bf16_kernel<<<(B * HID + 255) / 256, 256, 0, st>>>(L->h0, L->nd, B, HID, p.hm);
    lstm_fwd2_kernel<<<p.nbt * 32, NTHREADS, F2_SMEM_TOTAL, st>>>(ta, static_cast<const __nv_bfloat16*>(L->w_hh), p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      pvr_set_error("pvr_lstm_fwd2 launch: %s", cudaGetErrorString(e));
      return PVR_ERR_CUDA;
    }
    return PVR_OK;
  }
  if (!make_tmap_kchunks(&ta, L->hm, HID, (uint64_t)T * B, HID, 64, CPI, &err) ||
      !make_tmap_2d(&tw, L->w_hh, HID, 4 * HID, HID, 64, &err)) {
    pvr_set_error("pvr_lstm_persist_forward: tensor map: %s", err ? err : "?");
    return PVR_ERR_CUDA;
  }
  mask_state_bf16_kernel<<<(B * HID + 255) / 256, 256, 0, st>>>(L->h0, L->nd, B, HID, p.hm);
  return launch<0>(ta, tw, p, st, L->counters, L->counters_bytes);
}

// PVR_LSTM_BWD2=0 keeps the backward pass on the first cluster kernel (A/B measurements).
bool use_bwd2(int B) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("PVR_LSTM_BWD2");
    off = (e && e[0] == '0') ? 1 : 0;
  }
  return !off && ((B + F2_NB - 1) / F2_NB) * 8 <= bwd2_max_clusters();
}

int lstm_persist_backward(const pvr_lstm_bwd* L, cudaStream_t st) {
  const int T = L->T, B = L->B;
  PersistParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.B = B; p.nbt = (B + 63) / 64;
  p.nd = L->nd; p.c_all = const_cast<float*>(L->c_all); p.gates = const_cast<float*>(L->gates);
  p.dh_out = L->dh_out; p.dh_rec = L->dh_rec; p.dc_rec = L->dc_rec;
  p.dG = static_cast<__nv_bfloat16*>(L->dG);
  p.dbias = L->dbias;
  CUtensorMap ta, tw;
  const char* err = nullptr;
  if (use_bwd2(B)) {
    p.nbt = (B + F2_NB - 1) / F2_NB;
    if (!make_tmap_kchunks(&ta, L->dG, 4 * HID, (uint64_t)T * B, 4 * HID, F2_NB, F2_CPI, &err)) {
      pvr_set_error("pvr_lstm_persist_backward: tensor map: %s", err ? err : "?");
      return PVR_ERR_CUDA;
    }
    const int rc = ensure_ready((size_t)(T + 1) * p.nbt * 16, st, L->counters, L->counters_bytes, &p.ready);
    if (rc != PVR_OK) return rc;
    p.prof = g_prof;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    bwd2_config(&cfg, attr, p.nbt * 32, st);
    const cudaError_t e =
        cudaLaunchKernelEx(&cfg, lstm_bwd2_kernel, ta, static_cast<const __nv_bfloat16*>(L->w_hh_t), p);
    if (e != cudaSuccess) {
      pvr_set_error("pvr_lstm_bwd2 launch: %s", cudaGetErrorString(e));
      return PVR_ERR_CUDA;
    }
    return PVR_OK;
  }
  if (!make_tmap_kchunks(&ta, L->dG, 4 * HID, (uint64_t)T * B, 4 * HID, 64, CPI, &err) ||
      !make_tmap_2d(&tw, L->w_hh_t, 4 * HID, HID, 4 * HID, 64, &err)) {
    pvr_set_error("pvr_lstm_persist_backward: tensor map: %s", err ? err : "?");
    return PVR_ERR_CUDA;
  }
  return launch<1>(ta, tw, p, st, L->counters, L->counters_bytes);
}

}  // namespace pvr

extern "C" int pvr_lstm_persist_supported(int T, int B, int H) { return pvr::lstm_persist_supported(T, B, H); }

// Development aid: clock64 stamps of the phases of the first 64 steps of every CTA of the following launches are
// written to `buf` (device, 128 * 64 * 16 int64; NULL switches it off). Stamps of one CTA share one SM clock.
extern "C" int pvr_lstm_persist_profile(void* buf) {
  pvr::g_prof = static_cast<long long*>(buf);
  return PVR_OK;
}
