// Persistent CTA-pair tcgen05 GEMM (cta_group::2) for the FP16 hi/lo plane format -- the workhorse for every GEMM with N >= 384.
//
// Why (profiles/ncu_tc_r01.md, profiles/experiments_r01.md): the one-tile-per-CTA 128x256 kernel keeps the tensor pipe busy 55-59 % of the
// time.  Two causes: (1) its mainloop is bound by the L2 -> SM operand feed (96 KB of planes per k-block per SM; 1.6 GB through the
// crossbar for a 32768x1024x1024 GEMM, 6.7x the DRAM traffic); (2) per tile, CTA launch + barrier init + TMEM allocation + first-load
// latency + the whole store phase (~8 us) are serial with the mainloop.
//
// Here two CTAs on the SMs of one TPC compute a 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256):
//   * each CTA loads ITS 128 rows of the A planes and ITS 128 of the 256 B rows -- 64 KB per k-block per SM instead of 96 KB (the
//     tensor cores read the peer's B half through the pair's shared-memory path), so a 3-stage ring fits next to the staging buffers;
//   * one thread of the leader CTA issues the MMAs for both SMs; tcgen05.commit multicasts "stage free" / "partial ready" to both CTAs;
//   * the pair is PERSISTENT: it walks a static list of work items (split, 256-row tile, 256-column tile).  Barriers, TMEM and tensor
//     maps are set up once; the TMA producers run ahead into the next tile while the drain warps store the previous one.
//   * numerics: the tensor core accumulates with truncation, so nothing is ever accumulated in TMEM beyond ONE k-block (64 k): the 4
//     A_hi.B_hi MMAs and the 8 correction MMAs of a k-block go into a fresh 256-column TMEM buffer, the drain warps pull the partial out
//     with tcgen05.ld and add it to fp32 registers with round-to-nearest.  The two buffers (512 columns) let the MMA issuer run two
//     k-blocks ahead of the drain warps.  (gemm_tc256_kernel keeps the correction terms in a second tile across all of K instead; 12
//     truncating adds per partial instead of 4 cost < 1e-7 relative -- below the fp32 reference's own summation error.)
//   * store phase without a CTA-wide staging tile: each drain warp owns 32 rows x 128 columns in registers (thread = row), applies bias /
//     activation / mask bits / scale in that layout, and transposes 4 KB chunks (32 rows x 128 bytes, XOR-swizzled, conflict-free) through
//     a warp-private shared-memory buffer so that global stores are full 128-byte lines (4 rows x 128 B per instruction).  Column sums
//     (bias gradients) use a 124-shuffle butterfly.  Only the issuing warp's own progress gates the next tile's drain, so the next
//     tile's MMAs (which start as soon as the correction tile has been read out) overlap the stores.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace ase {

constexpr int TC2_DRAIN_WARPS = 16;                    // WG0..WG3: drain + store warps, TMEM lane quadrant (warp & 3) x 64-column quarter (warp >> 2)
constexpr int TC2_CW = 64;                             // accumulator columns per drain thread
constexpr int TC2_THREADS = 32 * (TC2_DRAIN_WARPS + 4);  // + WG4: warp 16 TMA producer, warp 17 MMA issuer (leader) / TMEM owner (2 idle warps)
// registers: the file is 4 x 16 K per SM and a sub-partition hosts 5 of the 20 warps, so the launch allocation is capped at 96 per thread.
// setmaxnreg moves registers between warpgroups WITHIN the CTA's launch allocation (640 x 96): WG4 releases (96 - 24) x 128 = 9216, which buys
// the 16 drain warps +16 each (112).  (Asking for 120 -- 3072 more than the pool holds -- parks the fourth warpgroup in setmaxnreg.inc
// forever: that was the deadlock of the first 16-warp build, profiles/experiments_r02.md.)
constexpr int TC2_STAGES = 3;
constexpr int TC2_PLANE_BYTES = 128 * 128;             // 128 operand rows x one 128-byte k-block row
constexpr int TC2_STAGE_BYTES = 4 * TC2_PLANE_BYTES;   // A_hi | A_lo | B_hi | B_lo of THIS CTA (its 128 rows of A, its 128 of the 256 B rows)
constexpr int TC2_STG_WARP_BYTES = 2048;               // per drain warp: 32 rows x 64 bytes
constexpr int TC2_SMEM_TOTAL = TC2_STAGES * TC2_STAGE_BYTES + TC2_DRAIN_WARPS * TC2_STG_WARP_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
static_assert(TC2_SMEM_TOTAL <= 232448, "exceeds the 227 KB a CTA can have");

__device__ __forceinline__ void sts128u(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128u(const float* p) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
// staging chunk of a drain warp: 32 rows x 64 bytes (4 pieces of 16 bytes); piece p of row r lives at physical piece p ^ ((r >> 1) & 3):
// a quarter warp writing one piece of 8 consecutive rows, or reading all pieces of 2 consecutive rows, touches all 32 banks exactly once
__device__ __forceinline__ float* stg_at(float* stg, int r, int p) { return stg + r * 16 + ((p ^ ((r >> 1) & 3)) << 2); }

// one stage of the column-sum butterfly: lanes whose `o` bit is clear keep columns [0, w), the others [w, 2w); partners exchange the rest
template <int o, int w>
__device__ __forceinline__ void colsum_stage(float (&a)[TC2_CW], int lane) {
  const bool up = (lane & o) != 0;
#pragma unroll
  for (int i = 0; i < w; ++i) {
    const float keep = up ? a[w + i] : a[i];
    const float send = up ? a[i] : a[w + i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
  }
}

// the two activity words of this thread's row for the warp's 64 columns (mask_mode 1 with bits): loaded BEFORE the tile's mainloop
__device__ __forceinline__ uint2 tc2_load_mask(const TcEpi& e, int64_t m, int nb) {
  uint2 w = make_uint2(0u, 0u);
  if (!(e.mask_mode == 1 && e.mask_bits) || nb >= e.N || m >= e.M) return w;
  const int nw = (min(TC2_CW, e.N - nb) + 31) >> 5;
  const uint32_t* mb = e.mask_bits + m * e.ldmb + (nb >> 5);
  if (nw == 2 && (e.ldmb & 1) == 0 && (reinterpret_cast<uintptr_t>(e.mask_bits) & 7) == 0) return __ldg(reinterpret_cast<const uint2*>(mb));
  w.x = __ldg(mb);
  if (nw > 1) w.y = __ldg(mb + 1);
  return w;
}
// bias of columns nb + 2 * lane, + 1 (zero past N): loaded BEFORE the tile's mainloop, broadcast through the staging buffer at the end
__device__ __forceinline__ float2 tc2_load_bias(const TcEpi& e, int nb, int lane, int z) {
  if (!e.bias || (e.accumulate && z != 0) || nb + 2 * lane >= e.N) return make_float2(0.0f, 0.0f);
  return __ldg(reinterpret_cast<const float2*>(e.bias + nb + 2 * lane));
}

// Store phase of one drain warp.  acc[c] = raw accumulator of C[mw0 + lane][nb + c], c < 64 (thread = row).  Everything that is arithmetic
// happens in this row layout, straight-line on registers (64 independent elements: full ILP, nothing from global memory in a dependency
// chain -- bias, mask words and scales were fetched before the mainloop): scale -> bias -> ReLU -> activity bits (one 8-byte store per row)
// -> mask bits -> max |C| -> FP16 hi/lo split.  Only finished BYTES go through the warp's 2 KB staging buffer (thread = row in, XOR-swizzled,
// conflict-free; 8 rows x 64 bytes per global store instruction out): fp32 C in 16-column chunks (or RED for split-K), half planes in
// 32-column chunks.  Column sums (bias gradients) use a 62-shuffle butterfly at the very end.  Same semantics and order as epilogue_rows /
// epilogue_fast in gemm_tc.cu.  Columns >= N hold exact zeros and are never stored.
// (History, profiles/experiments_r02.md: 8 drain warps x 128 columns ran the store phase at 35 % issue efficiency -- 2 warps per scheduler,
// spills at 128 accumulators + temporaries -- and it is serial with the next tile's drains (13.9 k of 39.3 k clocks per tile); 16 warps x 64
// columns halve the per-thread work and double the latency hiding (10.5 k clocks).  Earlier still: a version that loaded bias / masks here ran at 6 clocks per instruction on serialised L2
// latencies and instruction fetch; a rolled coalesced-layout version was bound by its dependency chains.)
__device__ __forceinline__ void tc2_store(const TcEpi& e, float (&acc)[TC2_CW], float* stg, int mw0, int nb, int lane, const float2 bias2, const uint2 maskw,
                                          float s1, float s2, float cscale) {
  const int ncols = min(TC2_CW, e.N - nb);           // valid columns of this warp (a multiple of 8: the host requires N % 8 == 0)
  if (ncols <= 0) return;
  const int64_t m = mw0 + lane;                      // this thread's row
  const int rr = lane >> 2, pp = lane & 3;           // coalesced pass: row 8 * it + rr, 16-byte piece pp
  // undo the operands' power-of-two plane scales (two exact multiplies; their product alone could underflow)
#pragma unroll
  for (int c = 0; c < TC2_CW; ++c) acc[c] = s2 * (s1 * acc[c]);
  if (e.bias) {                // uniform; bias2 is zero where it does not apply
    reinterpret_cast<float2*>(stg)[lane] = bias2;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < TC2_CW; c += 4) {
      const float4 b = lds128(stg + c);              // same address in every lane: a broadcast
      acc[c] += b.x; acc[c + 1] += b.y; acc[c + 2] += b.z; acc[c + 3] += b.w;
    }
    __syncwarp();
  }
  if (e.accumulate) {          // split-K / accumulating GEMMs (dW): fp32 RED into C, 8 rows x 64 bytes per instruction
#pragma unroll
    for (int q = 0; q < TC2_CW / 16; ++q) {
      if (q * 16 < ncols) {
#pragma unroll
        for (int p = 0; p < 4; ++p) sts128(stg_at(stg, lane, p), acc[q * 16 + 4 * p], acc[q * 16 + 4 * p + 1], acc[q * 16 + 4 * p + 2], acc[q * 16 + 4 * p + 3]);
        __syncwarp();
        const int col = q * 16 + pp * 4;
        float* cp = e.C + (int64_t)(mw0 + rr) * e.ldc + nb + col;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const float4 v = lds128(stg_at(stg, it * 8 + rr, pp));
          if (col < ncols) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          cp += 8 * e.ldc;
        }
        __syncwarp();
      }
    }
    return;
  }
  if (e.act == 1) {            // (tanh outputs are 64 wide in every network of the path: they never reach this kernel, see gemm_tc2_epilogue_ok)
#pragma unroll
    for (int c = 0; c < TC2_CW; ++c) acc[c] = fmaxf(acc[c], 0.0f);
  }
  const int nw = (ncols + 31) >> 5;                  // 32-column activity words this warp owns in its rows
  if (e.relu_bits) {
    uint32_t w[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      w[q] = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) w[q] |= (acc[q * 32 + i] > 0.0f) ? (1u << i) : 0u;
    }
    uint32_t* rb = e.relu_bits + m * e.ldrb + (nb >> 5);
    if (nw == 2 && (e.ldrb & 1) == 0 && (reinterpret_cast<uintptr_t>(e.relu_bits) & 7) == 0) *reinterpret_cast<uint2*>(rb) = make_uint2(w[0], w[1]);
    else {
      rb[0] = w[0];
      if (nw > 1) rb[1] = w[1];
    }
  }
  if (e.mask_mode == 1 && e.mask_bits) {
    const uint32_t w[2] = {maskw.x, maskw.y};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[q * 32 + i] = ((w[q] >> i) & 1u) ? acc[q * 32 + i] : 0.0f;
    }
  }
  if (e.c_amax || (e.Chi && e.flag)) {
    float amax = 0.0f;
#pragma unroll
    for (int c = 0; c < TC2_CW; ++c) amax = fmaxf(amax, fabsf(acc[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.0f) {
      if (e.c_amax) atomicMax(e.c_amax, __float_as_uint(amax));
      if (e.Chi && e.flag && !(amax * cscale <= 60000.0f)) atomicOr(e.flag, 1u);      // the predicted scale was too large: report, never saturate silently
      if (e.Chi && e.flag && cscale == 0.0f) atomicOr(e.flag, 2u);                    // the site only ever saw all-zero tensors, now there is data
    }
  }
  if (!e.skip_c) {             // fp32 C: 16-column chunks
#pragma unroll
    for (int q = 0; q < TC2_CW / 16; ++q) {
      if (q * 16 < ncols) {
#pragma unroll
        for (int p = 0; p < 4; ++p) sts128(stg_at(stg, lane, p), acc[q * 16 + 4 * p], acc[q * 16 + 4 * p + 1], acc[q * 16 + 4 * p + 2], acc[q * 16 + 4 * p + 3]);
        __syncwarp();
        const int col = q * 16 + pp * 4;
        float* cp = e.C + (int64_t)(mw0 + rr) * e.ldc + nb + col;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const float4 v = lds128(stg_at(stg, it * 8 + rr, pp));
          if (col < ncols) *reinterpret_cast<float4*>(cp) = v;
          cp += 8 * e.ldc;
        }
        __syncwarp();
      }
    }
  }
  if (e.Chi) {                 // half planes of C: 32-column chunks (64 bytes per row per plane); the lo words wait in registers for the hi chunk
#pragma unroll
    for (int ch = 0; ch < TC2_CW / 32; ++ch) {
      if (ch * 32 < ncols) {
        uint32_t lw[16];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          uint32_t hw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            split_f16x2(acc[ch * 32 + 8 * p + 2 * i] * cscale, acc[ch * 32 + 8 * p + 2 * i + 1] * cscale, hw[i], lw[4 * p + i]);
          sts128u(stg_at(stg, lane, p), hw[0], hw[1], hw[2], hw[3]);
        }
        const int col = ch * 32 + pp * 8;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          if (pl == 1) {
#pragma unroll
            for (int p = 0; p < 4; ++p) sts128u(stg_at(stg, lane, p), lw[4 * p], lw[4 * p + 1], lw[4 * p + 2], lw[4 * p + 3]);
          }
          __syncwarp();
          __half* dp = reinterpret_cast<__half*>(pl == 0 ? e.Chi : e.Clo) + (int64_t)(mw0 + rr) * e.ldp + nb + col;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const uint4 v = lds128u(stg_at(stg, it * 8 + rr, pp));
            if (col < ncols) *reinterpret_cast<uint4*>(dp) = v;
            dp += 8 * e.ldp;
          }
          __syncwarp();
        }
      }
    }
  }
  if (e.colsum) {              // bias gradient of the layer whose dZ this GEMM produces: sum over the warp's 32 rows, then one RED per column
    colsum_stage<16, 32>(acc, lane);
    colsum_stage<8, 16>(acc, lane);
    colsum_stage<4, 8>(acc, lane);
    colsum_stage<2, 4>(acc, lane);
    colsum_stage<1, 2>(acc, lane);
    const int cb = ((lane & 16) ? 32 : 0) + ((lane & 8) ? 16 : 0) + ((lane & 4) ? 8 : 0) + ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
#pragma unroll
    for (int j = 0; j < 2; ++j) if (cb + j < ncols) atomicAdd(e.colsum + nb + cb + j, acc[j]);
  }
}

// work item t -> (split z, 256-row tile, 256-column tile); consecutive items share A rows (L2 reuse across the concurrently running pairs)
struct Tc2Item { int m0, n0, kb_begin, nkb, z; };
__device__ __forceinline__ Tc2Item tc2_item(int t, int tiles_n, int tiles_m2, const TcEpi& e) {
  Tc2Item it;
  const int tn = t % tiles_n, tq = t / tiles_n;
  const int tm2 = tq % tiles_m2;
  it.z = tq / tiles_m2;
  it.m0 = tm2 * 256; it.n0 = tn * 256;
  it.kb_begin = it.z * e.kb_per_split;
  it.nkb = min(e.kb_per_split, e.kb_total - it.kb_begin);
  return it;
}

template <bool AMN, bool BMN>
__global__ void __launch_bounds__(TC2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const TcEpi e,
                const int tiles_n, const int tiles_m2, const int num_items) {
  using F = TcFmt<true>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + TC2_STAGES * TC2_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + TC2_DRAIN_WARPS * TC2_STG_WARP_BYTES);
  uint64_t* full = bars;                       // [3]  leader only: TMA bytes of BOTH CTAs -> MMA issuer
  uint64_t* empty = bars + TC2_STAGES;         // [3]  both CTAs: MMAs done with the stage (multicast commit) -> both producers
  uint64_t* buf_full = bars + 2 * TC2_STAGES;  // [2]  both CTAs: the k-block partial in TMEM buffer b is complete (multicast commit) -> drain warps
  uint64_t* buf_empty = buf_full + 2;          // [2]  leader only: the 32 drain warps of both CTAs pulled buffer b out of TMEM -> MMA issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(buf_full + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();                       // 0 = leader (issues the MMAs), 1 = peer
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int my_items = (num_items - pair + npairs - 1) / npairs;     // items pair, pair + npairs, ...

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAlo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBlo)) : "memory");
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&buf_full[b], 1); mbar_init(&buf_empty[b], 2 * TC2_DRAIN_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC2_DRAIN_WARPS + 1) {       // the same warp of BOTH CTAs allocates (and later frees) the pair's tensor memory: all 512 columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();          // the peer's barriers are initialised before any remote arrive / TMA completion reaches them
  pdl_sync();                  // nothing above touches global memory
  const uint32_t tmem_base = *tmem_slot;       // two 256-column accumulator buffers: k-block g lands in buffer g & 1

  if (warp >= TC2_DRAIN_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    if (warp == TC2_DRAIN_WARPS && lane == 0) {
      // ---------------- TMA producer (one per CTA): this CTA's 128 rows of A and its 128 of the tile's 256 B rows, every k-block of every
      // item; completion bytes go to the LEADER's full barrier
      const uint32_t full_leader = mapa_u32(smem_u32(&full[0]), 0);
      int g = 0;
      for (int j = 0; j < ((e.debug & 32) ? 0 : my_items); ++j) {
        const Tc2Item it = tc2_item(pair + j * npairs, tiles_n, tiles_m2, e);
        const bool pin = (e.debug & 16) != 0;       // ablation: every load hits the same (L2-resident) boxes
        const int m0 = (pin ? 0 : it.m0) + (int)rank * 128, nb0 = (pin ? 0 : it.n0) + (int)rank * 128;
        for (int kb = 0; kb < it.nkb; ++kb, ++g) {
          const int s = g % TC2_STAGES;
          const uint32_t ph = (uint32_t)(g / TC2_STAGES) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          if (rank == 0) mbar_expect_tx(&full[s], 2 * TC2_STAGE_BYTES);
          uint8_t* st = smem + s * TC2_STAGE_BYTES;
          const uint32_t bar = full_leader + 8u * (uint32_t)s;
          const int k0 = pin ? 0 : (it.kb_begin + kb) * F::BK;
          if (!AMN) {          // K-major planes [rows, K]: one box of 128 rows x one k-block
            tma_load_2d_pair(st, &tmAhi, bar, k0, m0);
            tma_load_2d_pair(st + TC2_PLANE_BYTES, &tmAlo, bar, k0, m0);
          } else {             // MN-major planes [K, rows]: boxes of 64 k-rows x 64 m
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              tma_load_2d_pair(st + b * F::MN_BOX_BYTES, &tmAhi, bar, m0 + b * F::MN_BOX, k0);
              tma_load_2d_pair(st + TC2_PLANE_BYTES + b * F::MN_BOX_BYTES, &tmAlo, bar, m0 + b * F::MN_BOX, k0);
            }
          }
          if (!BMN) {
            tma_load_2d_pair(st + 2 * TC2_PLANE_BYTES, &tmBhi, bar, k0, nb0);
            tma_load_2d_pair(st + 3 * TC2_PLANE_BYTES, &tmBlo, bar, k0, nb0);
          } else {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              tma_load_2d_pair(st + 2 * TC2_PLANE_BYTES + b * F::MN_BOX_BYTES, &tmBhi, bar, nb0 + b * F::MN_BOX, k0);
              tma_load_2d_pair(st + 3 * TC2_PLANE_BYTES + b * F::MN_BOX_BYTES, &tmBlo, bar, nb0 + b * F::MN_BOX, k0);
            }
          }
        }
      }
      // tail: every stage this producer filled has been consumed (and its multicast release has landed) before this CTA may exit
      for (int x = max(0, g - TC2_STAGES); x < g; ++x) mbar_wait(&empty[x % TC2_STAGES], (uint32_t)(x / TC2_STAGES) & 1u);
    } else if (warp == TC2_DRAIN_WARPS + 1 && lane == 0 && rank == 0) {
      // ---------------- MMA issuer (leader only): per k-block 4 main + 8 correction MMAs over both SMs into a FRESH TMEM buffer
      // (buffer g & 1); it may run two k-blocks ahead of the drain warps, which is what hides their store phase at tile boundaries
      const uint32_t idesc = tc_idesc2_f16(AMN, BMN, 256);
      int g = 0;
      for (int j = 0; j < my_items; ++j) {
        const Tc2Item it = tc2_item(pair + j * npairs, tiles_n, tiles_m2, e);
        for (int kb = 0; kb < it.nkb; ++kb, ++g) {
          const int s = g % TC2_STAGES;
          const uint32_t ph = (uint32_t)(g / TC2_STAGES) & 1u;
          const int b = g & 1;
          mbar_wait(&buf_empty[b], ((uint32_t)(g >> 1) & 1u) ^ 1u);     // both CTAs drained this buffer's previous partial (k-block g - 2)
          if (!(e.debug & 32)) mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * TC2_STAGE_BYTES);
          const uint32_t a_hi = sa, a_lo = sa + TC2_PLANE_BYTES, b_hi = sa + 2 * TC2_PLANE_BYTES, b_lo = sa + 3 * TC2_PLANE_BYTES;
          const uint32_t tmem_d = tmem_base + (uint32_t)(b * 256);
          // the correction terms go FIRST, into the fresh buffer: they are 2^-11 of the main term, so their truncating adds happen at a
          // 2^-11 finer scale; the 4 main MMAs on top then see exactly the 4 full-scale truncating adds per k-block of gemm_tc256_kernel
          const bool corr = !(e.debug & 4);
          if (corr)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc_mma2_f16(tmem_d, tc_desc<true, AMN>(a_lo, k), tc_desc<true, BMN>(b_hi, k), idesc, k > 0 ? 1u : 0u);
            tc_mma2_f16(tmem_d, tc_desc<true, AMN>(a_hi, k), tc_desc<true, BMN>(b_lo, k), idesc, 1u);
          }
          if (!(e.debug & 64))
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma2_f16(tmem_d, tc_desc<true, AMN>(a_hi, k), tc_desc<true, BMN>(b_hi, k), idesc, (corr || k > 0) ? 1u : 0u);
          tc_commit2_mc(&buf_full[b], 3);                        // the k-block's partial (main + correction terms) is complete in both CTAs
          tc_commit2_mc(&empty[s], 3);                           // ... and all 12 MMAs have read this stage
        }
      }
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ---------------- drain / store warps 0..15: TMEM lane quadrant lg, 64-column quarter cq of this CTA's 128 x 256 accumulator
    const int lg = warp & 3, cq = warp >> 2;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    float* stg = reinterpret_cast<float*>(stg_base + warp * TC2_STG_WARP_BYTES);
    // per-GEMM scalars (device memory, written by earlier kernels): fetched once, not in the store phase's dependency chains
    const float s1 = e.a_inv ? *e.a_inv : 1.0f;
    const float s2 = e.alpha * (e.b_inv ? *e.b_inv : 1.0f);
    const float cscale = (e.Chi && e.c_scale) ? *e.c_scale : 1.0f;
    const uint32_t buf_empty_leader = mapa_u32(smem_u32(&buf_empty[0]), 0);
    uint32_t g = 0;
    for (int j = 0; j < my_items; ++j) {
      const Tc2Item it = tc2_item(pair + j * npairs, tiles_n, tiles_m2, e);
      const int m0 = it.m0 + (int)rank * 128;
      const uint2 maskw = tc2_load_mask(e, (int64_t)m0 + lg * 32 + lane, it.n0 + cq * TC2_CW);     // in flight under the whole mainloop
      const float2 bias2 = tc2_load_bias(e, it.n0 + cq * TC2_CW, lane, it.z);
      float acc[TC2_CW];
#pragma unroll
      for (int i = 0; i < TC2_CW; ++i) acc[i] = 0.0f;
      for (int kb = 0; kb < it.nkb; ++kb, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait(&buf_full[b], (g >> 1) & 1u);
        tc_fence_after();
        if (!(e.debug & 2))
#pragma unroll
        for (int c = 0; c < TC2_CW; c += 32) {
          float v[32];
          tc_ld_32x32(tmem_base + lane_off + (uint32_t)(b * 256u + cq * TC2_CW + c), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[c + i] += v[i];      // round-to-nearest fp32 accumulation across k-blocks
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(buf_empty_leader + 8u * b);
      }
      if (m0 < e.M && !(e.debug & 1)) tc2_store(e, acc, stg, m0 + lg * 32, it.n0 + cq * TC2_CW, lane, bias2, maskw, s1, s2, cscale);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody exits (or frees tensor memory) while the peer may still signal this CTA or read its shared memory
  if (warp == TC2_DRAIN_WARPS + 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------ host side
bool gemm_tc2_epilogue_ok(const TcEpi& e) {
  if ((e.M & 127) || (e.N & 7)) return false;
  if ((e.ldc & 3) || (reinterpret_cast<uintptr_t>(e.C) & 15)) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  if (e.Chi && ((e.ldp & 7) || (reinterpret_cast<uintptr_t>(e.Chi) & 15) || (reinterpret_cast<uintptr_t>(e.Clo) & 15))) return false;
  if (e.mask_mode && !(e.mask_mode == 1 && e.mask_bits)) return false;       // fp32 masks (tanh', TF32-era callers) stay on the 128-row kernels
  if (e.act == 2) return false;                                              // tanh epilogues too (128 inlined tanhf would triple the code size)
  return true;
}

template <bool AMN, bool BMN>
static int tc2_slots() {
  static int slots = 0;
  if (slots) return slots;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int n = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * (sms / 2)); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = TC2_SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<AMN, BMN>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = sms / 2; }
  slots = n > sms / 2 ? sms / 2 : n;
  return slots;
}
int gemm_tc2_pair_slots() { return 74; }     // planning figure for the split-K heuristic (148 SMs); launches use the queried occupancy

template <bool AMN, bool BMN>
static int launch_tc2_t(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcEpi& e, int splits, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    ASE_CUDA_OK(cudaFuncSetAttribute(gemm_tc2_kernel<AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_TOTAL));
    attr_set = true;
  }
  const int tiles_n = ceil_div(e.N, 256), tiles_m2 = ceil_div(e.M, 256);
  const int items = tiles_n * tiles_m2 * splits;
  const int pairs = min(items, tc2_slots<AMN, BMN>());
  const bool prof = tc_prof_on();
  if (prof) tc_prof_mark(st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = TC2_SMEM_TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = tc_pdl();
  cfg.attrs = attr; cfg.numAttrs = 2;
  ASE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<AMN, BMN>, ah, al, bh, bl, e, tiles_n, tiles_m2, items));
  if (prof) { tc_prof_mark(st); tc_prof_add_flops(2.0 * (double)e.M * (double)e.N * (double)e.K); }
  ASE_LAUNCH_OK();
  return ASE_OK;
}

int launch_tc2(bool amn, bool bmn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcEpi& e,
               int splits, cudaStream_t st) {
  if (!amn && !bmn) return launch_tc2_t<false, false>(ah, al, bh, bl, e, splits, st);
  if (!amn && bmn) return launch_tc2_t<false, true>(ah, al, bh, bl, e, splits, st);
  if (amn && !bmn) return launch_tc2_t<true, false>(ah, al, bh, bl, e, splits, st);
  return launch_tc2_t<true, true>(ah, al, bh, bl, e, splits, st);
}

}  // namespace ase
