// tcgen05 GEMM for sm_100a with fp32-class accuracy: three tensor-core MMAs per product on hi/lo operand planes.
//
//   C[M,N] = epi(alpha * A . B^T),  A = A_hi + A_lo, B = B_hi + B_lo
//   A.B^T ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo       (dropped A_lo.B_lo term is 2^-22 relative)
//
// Plane formats (template parameter H): scaled FP16 halfs (kind::f16, backend 2, default) or fp32 words holding TF32 values
// (kind::tf32, backend 1).  Planes are kept in the tensors' natural row-major layout and read K-major or MN-major by TMA / UMMA
// descriptors -- nothing is transposed -- and are normally WRITTEN by the epilogue of the GEMM that produces the tensor.
//
// Kernels (all: warp 0 = TMA producer into a shared-memory ring with mbarrier completion, warp 1 = one thread issuing
// tcgen05.mma into TMEM, drain / epilogue warps that pull every k-block's A_hi.B_hi partial out of TMEM with tcgen05.ld and
// accumulate it in fp32 registers with round-to-nearest adds -- the tensor core's own accumulation truncates):
//   gemm_tc256_kernel  128 x 256 tile, 2-stage ring, 8 drain warps            (N >= 384; the workhorse)
//   gemm_tc_kernel     128 x 128 / 128 x 64 tile, double-buffered main tile   (narrow N)
// Store phase: registers -> shared staging tile -> coalesced pass with bias / ReLU / tanh / mask (fp32 or 1-bit), optional
// fp32 C, half planes (predicted power-of-two scale), ReLU activity bits, max |C|, fused column sums, split-K fp32 RED.
// Descriptor formats follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction descriptor" tables
// (cross-checked against cute/arch/mma_sm100_desc.hpp and cute/atom/mma_traits_sm100.hpp in the image's CUTLASS headers).
// DESIGN.md section 5 has the numerics and the measurements.
#include <cuda.h>
#include <cuda_fp16.h>
#include <vector>
#include <unordered_map>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace ase {


// Phase 2 of the epilogue, shared by the tile shapes: a warp writes rows [row0, row0+nrows) of the staged tile; a row is
// written as BNT/4 float4 by consecutive lanes (full 128-byte lines); bias / activation / mask operands are read with the
// same coalesced mapping; optional hi/lo planes of C, the ReLU activity bits, the running max |C| and the fused column sums
// (bias gradient; reduced over the tile's rows in shared memory first: s_colsum[BNT], zeroed by the caller) ride along.
// Every lane stays in both loops for the whole tile (columns past N are predicated off): the bit packing shuffles need all 32.
template <bool H>
__device__ __forceinline__ void epilogue_rows(const TcEpi& e, const float* cs, int cs_ld, float* s_colsum, int row0, int nrows, int BNT, int m0, int n0, int lane) {
  const bool vec_ok = ((e.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.C) & 15) == 0) &&
                      (!e.mask_mode || e.mask_bits || (((e.ldm & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.mask_src) & 15) == 0)));
  const bool add_bias = e.bias && (!e.accumulate || blockIdx.z == 0);
  const uintptr_t pal = H ? 7 : 15;         // 4 plane elements per lane: 8 bytes (halfs) / 16 bytes (fp32 words)
  const bool pvec = e.Chi && ((e.ldp & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.Chi) & pal) == 0) && ((reinterpret_cast<uintptr_t>(e.Clo) & pal) == 0);
  const float cscale = (H && e.Chi && e.c_scale) ? *e.c_scale : 1.0f;
  const bool use_bits = e.mask_mode == 1 && e.mask_bits != nullptr;
  const int nib_shift = 4 * (lane & 7);
  float amax = 0.0f;
#pragma unroll 1
  for (int cc = lane * 4; cc < ((BNT + 127) & ~127); cc += 128) {          // same trip count for every lane (BNT = 64: lanes 16.. idle along)
    const int n = n0 + cc;
    const int nvalid = (cc < BNT) ? max(0, min(4, e.N - n)) : 0;
    float bv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (add_bias) for (int j = 0; j < nvalid; ++j) bv[j] = e.bias[n + j];
    float cs4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
      const int row = row0 + r, m = m0 + row;
      if (m >= e.M) break;                 // warp-uniform
      const float4 t = (cc < BNT) ? lds128(cs + row * cs_ld + cc) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float x[4] = {t.x + bv[0], t.y + bv[1], t.z + bv[2], t.w + bv[3]};
      float* cp = e.C + (int64_t)m * e.ldc + n;
      if (e.accumulate) {
        if (vec_ok && nvalid == 4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(x[0]), "f"(x[1]), "f"(x[2]), "f"(x[3]) : "memory");
        else for (int j = 0; j < nvalid; ++j) atomicAdd(cp + j, x[j]);
        continue;
      }
      if (e.act == 1) { for (int j = 0; j < 4; ++j) x[j] = fmaxf(x[j], 0.0f); }
      else if (e.act == 2) { for (int j = 0; j < 4; ++j) x[j] = tanhf(x[j]); }
      if (e.relu_bits) {                   // 8 lanes x 4 columns = one 32-bit word of the row's activity mask
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) w |= (j < nvalid && x[j] > 0.0f) ? (1u << j) : 0u;
        w <<= nib_shift;
        w |= __shfl_xor_sync(0xffffffffu, w, 1); w |= __shfl_xor_sync(0xffffffffu, w, 2); w |= __shfl_xor_sync(0xffffffffu, w, 4);
        if ((lane & 7) == 0 && nvalid > 0) e.relu_bits[(int64_t)m * e.ldrb + (n >> 5)] = w;
      }
      if (use_bits) {
        const uint32_t w = (nvalid > 0 && !(e.debug & 16)) ? e.mask_bits[(int64_t)m * e.ldmb + (n >> 5)] >> nib_shift : 0xFu;
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ((w >> j) & 1u) ? x[j] : 0.0f;
      } else if (e.mask_mode && !(e.debug & 16) && nvalid > 0) {
        const float* mp = e.mask_src + (int64_t)m * e.ldm + n;
        float mv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (vec_ok && nvalid == 4) { const float4 q = *reinterpret_cast<const float4*>(mp); mv[0] = q.x; mv[1] = q.y; mv[2] = q.z; mv[3] = q.w; }
        else for (int j = 0; j < nvalid; ++j) mv[j] = mp[j];
        if (e.mask_mode == 1) { for (int j = 0; j < 4; ++j) x[j] = (mv[j] > 0.0f) ? x[j] : 0.0f; }
        else { for (int j = 0; j < 4; ++j) x[j] *= (1.0f - mv[j] * mv[j]); }
      }
      if (e.skip_c || (e.debug & 128)) {}
      else if (vec_ok && nvalid == 4) *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
      else for (int j = 0; j < nvalid; ++j) cp[j] = x[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) { cs4[j] += x[j]; if (j < nvalid) amax = fmaxf(amax, fabsf(x[j])); }
      if (e.Chi && !(e.debug & 32) && nvalid > 0) {       // the consumers of C read these planes directly through TMA: no separate split pass
        if (!H) {
          float h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) split_tf32(x[j], h[j], l[j]);
          float* hp = (float*)e.Chi + (int64_t)m * e.ldp + n; float* lp = (float*)e.Clo + (int64_t)m * e.ldp + n;
          if (pvec && nvalid == 4) { *reinterpret_cast<float4*>(hp) = make_float4(h[0], h[1], h[2], h[3]); *reinterpret_cast<float4*>(lp) = make_float4(l[0], l[1], l[2], l[3]); }
          else for (int j = 0; j < nvalid; ++j) { hp[j] = h[j]; lp[j] = l[j]; }
        } else {
          uint2 hv, lv;
          split_f16x2(x[0] * cscale, x[1] * cscale, hv.x, lv.x); split_f16x2(x[2] * cscale, x[3] * cscale, hv.y, lv.y);
          __half* hp = (__half*)e.Chi + (int64_t)m * e.ldp + n; __half* lp = (__half*)e.Clo + (int64_t)m * e.ldp + n;
          if (pvec && nvalid == 4) { *reinterpret_cast<uint2*>(hp) = hv; *reinterpret_cast<uint2*>(lp) = lv; }
          else {
            const uint32_t hw[2] = {hv.x, hv.y}, lw[2] = {lv.x, lv.y};
            for (int j = 0; j < nvalid; ++j) {
              hp[j] = __ushort_as_half((unsigned short)(hw[j >> 1] >> (16 * (j & 1))));
              lp[j] = __ushort_as_half((unsigned short)(lw[j >> 1] >> (16 * (j & 1))));
            }
          }
        }
      }
    }
    if (e.colsum && !e.accumulate && !(e.debug & 64)) for (int j = 0; j < nvalid; ++j) atomicAdd(s_colsum + cc + j, cs4[j]);   // shared-memory atomics
  }
  if ((e.c_amax || (H && e.Chi && e.flag)) && !e.accumulate) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.0f) {
      if (e.c_amax) atomicMax(e.c_amax, __float_as_uint(amax));
      if (H && e.Chi && e.flag && !(amax * cscale <= 60000.0f)) atomicOr(e.flag, 1u);     // the predicted scale was too large: report, never saturate silently
      if (H && e.Chi && e.flag && cscale == 0.0f) atomicOr(e.flag, 2u);                   // the site only ever saw all-zero tensors, now there is data
    }
  }
}

// Fast store phase of the 128x256 kernel for interior, aligned tiles (the generic epilogue_rows above handles everything else).
// The generic loop spends ~500 issue slots per (row, 4-column) item on predicates, 64-bit address arithmetic and scalar tails
// (ncu r01b: 43 % of the kernel's lifetime, issue-bound, not memory-bound).  Here a lane owns 8 consecutive columns of each of the
// warp's 16 rows: two LDS.128 from the staged tile, one 16-byte store per half plane, pointers advanced by the row strides, the
// rows' activity-mask words prefetched before the loop, no bounds checks.
__device__ __forceinline__ bool epilogue_fast_ok(const TcEpi& e, int m0, int n0, int bn, bool H) {
  if (e.accumulate || m0 + TC_BM > e.M || n0 + bn > e.N) return false;
  if ((e.ldc & 3) || (reinterpret_cast<uintptr_t>(e.C) & 15)) return false;
  if (e.bias && (reinterpret_cast<uintptr_t>(e.bias) & 15)) return false;
  if (e.Chi && (!H || (e.ldp & 7) || (reinterpret_cast<uintptr_t>(e.Chi) & 15) || (reinterpret_cast<uintptr_t>(e.Clo) & 15))) return false;
  if (e.mask_mode && !(e.mask_mode == 1 && e.mask_bits) && ((e.ldm & 3) || (reinterpret_cast<uintptr_t>(e.mask_src) & 15))) return false;
  return true;
}

// CPL = columns per lane (8: 256-wide tile, 4: 128-wide tile), NR = rows per warp (a multiple of 4)
template <bool H, int CPL, int NR>
__device__ __forceinline__ void epilogue_fast(const TcEpi& e, const float* cs, int cs_ld, float* s_colsum, int row0, int m0, int n0, int lane) {
  static_assert(CPL == 4 || CPL == 8, "columns per lane");
  constexpr int LPW = 32 / CPL;                        // lanes per 32-column activity word
  const int c8 = lane * CPL, n = n0 + c8;
  const int64_t m = m0 + row0;
  float bv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) bv[j] = 0.0f;
  if (e.bias) {
#pragma unroll
    for (int j = 0; j < CPL; j += 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n + j);
      bv[j] = b0.x; bv[j + 1] = b0.y; bv[j + 2] = b0.z; bv[j + 3] = b0.w;
    }
  }
  const float cscale = (H && e.Chi && e.c_scale) ? *e.c_scale : 1.0f;
  const bool use_bits = e.mask_mode == 1 && e.mask_bits != nullptr;
  const int sh = CPL * (lane & (LPW - 1));             // this lane's bits of the 32-column activity word
  // activity-mask words: 4 rows per group, the next group's words are in flight while the current one is processed
  const uint32_t* mb = use_bits ? e.mask_bits + m * e.ldmb + (n >> 5) : nullptr;
  uint32_t mw[4] = {0, 0, 0, 0}, mwn[4] = {0, 0, 0, 0};
  if (use_bits) {
#pragma unroll
    for (int r = 0; r < 4; ++r) mw[r] = __ldg(mb + (int64_t)r * e.ldmb);
  }
  float* cp = e.C + m * e.ldc + n;
  __half* hp = e.Chi ? (__half*)e.Chi + m * e.ldp + n : nullptr;
  __half* lp = e.Chi ? (__half*)e.Clo + m * e.ldp + n : nullptr;
  uint32_t* rb = e.relu_bits ? e.relu_bits + m * e.ldrb + (n >> 5) : nullptr;
  const float* mp = (e.mask_mode && !use_bits) ? e.mask_src + m * e.ldm + n : nullptr;
  const float* sp = cs + row0 * cs_ld + c8;
  float csum[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) csum[j] = 0.0f;
  float amax = 0.0f;
#pragma unroll 1
  for (int g = 0; g < NR / 4; ++g) {
    if (use_bits && g + 1 < NR / 4) {
#pragma unroll
      for (int r = 0; r < 4; ++r) mwn[r] = __ldg(mb + (int64_t)(4 * (g + 1) + r) * e.ldmb);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float x[CPL];
#pragma unroll
      for (int j = 0; j < CPL; j += 4) {
        const float4 t = lds128(sp + j);
        x[j] = t.x + bv[j]; x[j + 1] = t.y + bv[j + 1]; x[j + 2] = t.z + bv[j + 2]; x[j + 3] = t.w + bv[j + 3];
      }
      if (e.act == 1) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) x[j] = fmaxf(x[j], 0.0f);
      } else if (e.act == 2) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) x[j] = tanhf(x[j]);
      }
      if (rb) {              // a lane's bits are one byte (CPL = 8) / one nibble (CPL = 4) of the row's little-endian activity words
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) w |= (x[j] > 0.0f) ? (1u << j) : 0u;
        if (CPL == 8) reinterpret_cast<uint8_t*>(rb)[lane & 3] = (uint8_t)w;
        else {
          w |= __shfl_xor_sync(0xffffffffu, w << 4, 1) & 0xF0u;          // even lanes pick up the odd neighbour's nibble
          if (!(lane & 1)) reinterpret_cast<uint8_t*>(rb)[(lane & 7) >> 1] = (uint8_t)w;
        }
        rb += e.ldrb;
      }
      if (use_bits) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) x[j] = ((mw[r] >> (sh + j)) & 1u) ? x[j] : 0.0f;
      } else if (mp) {
        float mv[CPL];
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
          const float4 q = *reinterpret_cast<const float4*>(mp + j);
          mv[j] = q.x; mv[j + 1] = q.y; mv[j + 2] = q.z; mv[j + 3] = q.w;
        }
        if (e.mask_mode == 1) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) x[j] = (mv[j] > 0.0f) ? x[j] : 0.0f;
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) x[j] *= (1.0f - mv[j] * mv[j]);
        }
        mp += e.ldm;
      }
      if (!e.skip_c) {
#pragma unroll
        for (int j = 0; j < CPL; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      }
#pragma unroll
      for (int j = 0; j < CPL; ++j) { csum[j] += x[j]; amax = fmaxf(amax, fabsf(x[j])); }
      if (H && hp) {
        uint32_t hw[CPL / 2], lw[CPL / 2];
#pragma unroll
        for (int j = 0; j < CPL; j += 2) split_f16x2(x[j] * cscale, x[j + 1] * cscale, hw[j >> 1], lw[j >> 1]);
        if (CPL == 8) {
          *reinterpret_cast<uint4*>(hp) = make_uint4(hw[0], hw[1], hw[CPL / 2 - 2], hw[CPL / 2 - 1]);
          *reinterpret_cast<uint4*>(lp) = make_uint4(lw[0], lw[1], lw[CPL / 2 - 2], lw[CPL / 2 - 1]);
        } else {
          *reinterpret_cast<uint2*>(hp) = make_uint2(hw[0], hw[1]);
          *reinterpret_cast<uint2*>(lp) = make_uint2(lw[0], lw[1]);
        }
        hp += e.ldp; lp += e.ldp;
      }
      cp += e.ldc; sp += cs_ld;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) mw[r] = mwn[r];
  }
  if (e.colsum) {
#pragma unroll
    for (int j = 0; j < CPL; ++j) atomicAdd(s_colsum + c8 + j, csum[j]);
  }
  if (e.c_amax || (H && e.Chi && e.flag)) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.0f) {
      if (e.c_amax) atomicMax(e.c_amax, __float_as_uint(amax));
      if (H && e.Chi && e.flag && !(amax * cscale <= 60000.0f)) atomicOr(e.flag, 1u);
      if (H && e.Chi && e.flag && cscale == 0.0f) atomicOr(e.flag, 2u);
    }
  }
}

template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BM * 128;                // 16 KB per plane: one 128-byte k-block row per operand row
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};


// fp32-exact accumulation ("accumulate outside the tensor core", Ootomo & Yokota 2022, mapped to TMEM):
// the tensor core adds into its accumulator with truncation (measured: one-sided drift of ~0.1 ulp per MMA), which
// over K/8 x 3 sequential MMAs costs ~1e-5 absolute on O(1) sums -- enough to flip ReLU masks against the fp32
// reference.  So the main term A_hi.B_hi is accumulated in TMEM only WITHIN one 32-wide k-block (4 MMAs, fresh
// accumulator), double buffered; the epilogue warps drain each k-block's partial tile with tcgen05.ld and add it
// into fp32 registers with round-to-nearest FADDs while the tensor core works on the next k-block.  The two
// correction terms (2^-11 smaller) accumulate across all of K in a third TMEM tile: their truncation is negligible.
template <int BN, int STAGES, bool AMN, bool BMN, bool H>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const TcEpi e) {
  using SM = TcSmem<BN, STAGES>;
  using F = TcFmt<H>;
  constexpr uint32_t TMEM_COLS = (3 * BN <= 256) ? 256u : 512u;      // main[0], main[1], corr
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * SM::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* main_full = bars + 2 * STAGES;         // [2]
  uint64_t* main_empty = bars + 2 * STAGES + 2;    // [2]
  uint64_t* corr_full = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int kb_begin = blockIdx.z * e.kb_per_split;
  const int nkb = min(e.kb_per_split, e.kb_total - kb_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&main_full[b], 1); mbar_init(&main_empty[b], 4); }   // 4 epilogue warps arrive
    mbar_init(corr_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_corr = tmem_base + 2 * BN;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], SM::STAGE_BYTES);
        uint8_t* st = smem + s * SM::STAGE_BYTES;
        const int k0 = (kb_begin + kb) * F::BK;
        if (!AMN) {          // K-major planes [rows, K]: one box of 128 rows x one k-block
          tma_load_2d(st, &tmAhi, &full[s], k0, m0);
          tma_load_2d(st + SM::A_BYTES, &tmAlo, &full[s], k0, m0);
        } else {             // MN-major planes [K, rows]: boxes of BK k-rows x 128 bytes of m
#pragma unroll
          for (int b = 0; b < TC_BM / F::MN_BOX; ++b) {
            tma_load_2d(st + b * F::MN_BOX_BYTES, &tmAhi, &full[s], m0 + b * F::MN_BOX, k0);
            tma_load_2d(st + SM::A_BYTES + b * F::MN_BOX_BYTES, &tmAlo, &full[s], m0 + b * F::MN_BOX, k0);
          }
        }
        if (!BMN) {
          tma_load_2d(st + 2 * SM::A_BYTES, &tmBhi, &full[s], k0, n0);
          tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES, &tmBlo, &full[s], k0, n0);
        } else {
#pragma unroll
          for (int b = 0; b < BN / F::MN_BOX; ++b) {
            tma_load_2d(st + 2 * SM::A_BYTES + b * F::MN_BOX_BYTES, &tmBhi, &full[s], n0 + b * F::MN_BOX, k0);
            tma_load_2d(st + 2 * SM::A_BYTES + SM::B_BYTES + b * F::MN_BOX_BYTES, &tmBlo, &full[s], n0 + b * F::MN_BOX, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc_idesc<H>(AMN, BMN, BN);
      // per MMA slice (32 bytes of K): K-major operands advance 32 bytes along the swizzled row, MN-major ones F::MN_KSTEP
      auto adesc = [](uint32_t base, int k) { return tc_desc<H, AMN>(base, k); };
      auto bdesc = [](uint32_t base, int k) { return tc_desc<H, BMN>(base, k); };
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        const int b = kb & 1;
        const uint32_t bph = (kb >> 1) & 1;
        mbar_wait(&full[s], ph);
        mbar_wait(&main_empty[b], bph ^ 1);       // the drain of the k-block that used this TMEM tile two steps ago
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint32_t a_hi = sa, a_lo = sa + SM::A_BYTES, b_hi = sa + 2 * SM::A_BYTES, b_lo = b_hi + SM::B_BYTES;
        const uint32_t tmem_main = tmem_base + (uint32_t)b * BN;
#pragma unroll
        for (int k = 0; k < 4; ++k)     // 4 MMAs per k-block: 8 tf32 / 16 f16 = 32 bytes along the swizzled row each
          tc_mma<H>(tmem_main, adesc(a_hi, k), bdesc(b_hi, k), idesc, k > 0 ? 1u : 0u);
        tc_commit(&main_full[b]);              // this k-block's A_hi.B_hi partial tile is complete
        if (!(e.debug & 4))
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma<H>(tmem_corr, adesc(a_lo, k), bdesc(b_hi, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_mma<H>(tmem_corr, adesc(a_hi, k), bdesc(b_lo, k), idesc, 1u);
        }
        tc_commit(&empty[s]);     // all 12 MMAs have read this smem stage
      }
      tc_commit(corr_full);
    }
    __syncwarp();
  } else {
    // epilogue / drain warps 2..5 own TMEM lane groups (warp % 4); one output row per thread
    const int lg = warp & 3;
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
    for (int kb = 0; kb < nkb; ++kb) {
      const int b = kb & 1;
      const uint32_t bph = (kb >> 1) & 1;
      mbar_wait(&main_full[b], bph);
      tc_fence_after();
      if (!(e.debug & 2))
#pragma unroll
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc_ld_32x32(tmem_base + lane_off + (uint32_t)(b * BN + c), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c + j] += v[j];      // round-to-nearest fp32 accumulation across k-blocks
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&main_empty[b]);
    }
    mbar_wait(corr_full, 0);
    tc_fence_after();
    // Phase 1: (main + correction) -> shared staging tile [128][BN+4].  The mainloop stages are free now: every MMA
    // has completed (corr_full) and every TMA load was consumed.  One accumulator row per thread.
    float* cs = reinterpret_cast<float*>(smem);
    constexpr int CS_LD = BN + 4;
    // FP16 planes: undo the operands' power-of-two scales (two exact multiplies; their product alone could underflow)
    const float s1 = (H && e.a_inv) ? *e.a_inv : 1.0f;
    const float s2 = e.alpha * ((H && e.b_inv) ? *e.b_inv : 1.0f);
    {
      float* crow_s = cs + (lg * 32 + lane) * CS_LD;
#pragma unroll
      for (int c = 0; c < BN; c += 32) {
        float v[32];
        tc_ld_32x32(tmem_corr + lane_off + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(crow_s + c + j, s2 * (s1 * (acc[c + j] + v[j])), s2 * (s1 * (acc[c + j + 1] + v[j + 1])),
                 s2 * (s1 * (acc[c + j + 2] + v[j + 2])), s2 * (s1 * (acc[c + j + 3] + v[j + 3])));
      }
    }
    float* s_colsum = cs + TC_BM * CS_LD;               // [BN] per-tile column sums, behind the staging tile (pipeline smem is idle)
    if (e.colsum) for (int c = (lg * 32 + lane); c < BN; c += 128) s_colsum[c] = 0.0f;
    asm volatile("bar.sync 1, 128;" ::: "memory");      // the 4 epilogue warps only
    // Phase 2: coalesced epilogue -- warp w owns rows [32w, 32w+32)
    if (!(e.debug & 1)) {
      if (BN == 128 && epilogue_fast_ok(e, m0, n0, BN, H) && !(e.debug & 256)) epilogue_fast<H, 4, 32>(e, cs, CS_LD, s_colsum, lg * 32, m0, n0, lane);
      else epilogue_rows<H>(e, cs, CS_LD, s_colsum, lg * 32, 32, BN, m0, n0, lane);
    }
    if (e.colsum && !e.accumulate) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = (lg * 32 + lane); c < BN; c += 128) if (n0 + c < e.N) atomicAdd(e.colsum + n0 + c, s_colsum[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}


// ---------------------------------------------------------------------------------------------------------
// 128 x 256 tile variant.  A 128x128 tile needs 64 KB of operand planes per 768 MMA clocks = 85 B/clk, above the
// ~64 B/clk a B200 SM can ingest (profiles/experiments_r01.md), and its MMAs read shared memory at the full
// 128 B/clk.  Doubling N amortises the A planes: 96 KB per 1536 clocks = 62 B/clk ingest, 96 B/clk smem reads.
// TMEM: main (256 cols, single buffered) + corr (256 cols) = all 512 columns.  8 drain/epilogue warps (two per TMEM lane
// quadrant, 128 columns each) keep the per-thread accumulator at 128 registers; setmaxnreg moves registers from the
// producer / MMA warpgroup to them.  The drain of k-block kb runs under the 8 correction MMAs of kb (1024 clocks).
// ---------------------------------------------------------------------------------------------------------
constexpr int TC256_THREADS = 384;       // WG0: warp 0 TMA, warp 1 MMA (+2 idle); WG1, WG2: drain / epilogue
constexpr int TC256_BN = 256;
constexpr int TC256_STAGES = 2;

// One k-block of operand planes into the 128x256 kernel's ring (2 stages x 96 KB; producer thread only); it is called from
// two places (before and after the CTA-wide setup barrier).
template <bool AMN, bool BMN, bool H>
__device__ __forceinline__ void tc256_issue_kb(const CUtensorMap* tmAhi, const CUtensorMap* tmAlo, const CUtensorMap* tmBhi, const CUtensorMap* tmBlo,
                                               uint8_t* smem, uint64_t* full, uint64_t* empty, int kb, int kb_begin, int m0, int n0) {
  constexpr int BN = 256, STAGES = 2;
  using F = TcFmt<H>;
  constexpr int KW = F::BK;                                   // k elements per ring slot
  constexpr int ROWB = 128;                                   // bytes per K-major operand row in a slot
  constexpr int A_BYTES = TC_BM * ROWB, B_BYTES = BN * ROWB, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  constexpr int MN_BOX_BYTES = F::MN_BOX_BYTES;
  const int s = kb % STAGES;
  const uint32_t ph = (kb / STAGES) & 1;
  mbar_wait(&empty[s], ph ^ 1);
  mbar_expect_tx(&full[s], STAGE_BYTES);
  uint8_t* st = smem + s * STAGE_BYTES;
  const int k0 = (kb_begin + kb) * KW;
  if (!AMN) {
    tma_load_2d(st, tmAhi, &full[s], k0, m0);
    tma_load_2d(st + A_BYTES, tmAlo, &full[s], k0, m0);
  } else {
#pragma unroll
    for (int b = 0; b < TC_BM / F::MN_BOX; ++b) {
      tma_load_2d(st + b * MN_BOX_BYTES, tmAhi, &full[s], m0 + b * F::MN_BOX, k0);
      tma_load_2d(st + A_BYTES + b * MN_BOX_BYTES, tmAlo, &full[s], m0 + b * F::MN_BOX, k0);
    }
  }
  if (!BMN) {        // the B maps have 128-row boxes: two per plane
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tma_load_2d(st + 2 * A_BYTES + h * (128 * ROWB), tmBhi, &full[s], k0, n0 + h * 128);
      tma_load_2d(st + 2 * A_BYTES + B_BYTES + h * (128 * ROWB), tmBlo, &full[s], k0, n0 + h * 128);
    }
  } else {
#pragma unroll
    for (int b = 0; b < BN / F::MN_BOX; ++b) {
      tma_load_2d(st + 2 * A_BYTES + b * MN_BOX_BYTES, tmBhi, &full[s], n0 + b * F::MN_BOX, k0);
      tma_load_2d(st + 2 * A_BYTES + B_BYTES + b * MN_BOX_BYTES, tmBlo, &full[s], n0 + b * F::MN_BOX, k0);
    }
  }
}

template <bool AMN, bool BMN, bool H>
__global__ void __launch_bounds__(TC256_THREADS, 1)
gemm_tc256_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
                  const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const TcEpi e) {
  constexpr int BN = TC256_BN, STAGES = TC256_STAGES;
  using SM = TcSmem<BN, TC256_STAGES>;
  using F = TcFmt<H>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC256_STAGES * SM::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* main_full = bars + 2 * STAGES;
  uint64_t* main_empty = bars + 2 * STAGES + 1;
  uint64_t* corr_full = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;
  const int kb_begin = blockIdx.z * e.kb_per_split;
  const int nkb = min(e.kb_per_split, e.kb_total - kb_begin);

  auto issue_kb = [&](int kb) {
    tc256_issue_kb<AMN, BMN, H>(&tmAhi, &tmAlo, &tmBhi, &tmBlo, smem, full, empty, kb, kb_begin, m0, n0);
  };
  const int nslots = nkb;                     // ring slots this CTA streams
  const int kb_early = min(STAGES, nslots);   // slots whose loads are issued before the CTA-wide setup barrier
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAlo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBhi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBlo)) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(main_full, 1); mbar_init(main_empty, 8); mbar_init(corr_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first stages' loads fly while TMEM is allocated and the CTA synchronises (they only need the producer's own barriers)
    pdl_sync();
    for (int kb = 0; kb < kb_early; ++kb) issue_kb(kb);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_corr = tmem_base + BN;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      for (int kb = kb_early; kb < nslots; ++kb) issue_kb(kb);
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc = tc_idesc<H>(AMN, BMN, BN);
      auto adesc = [](uint32_t base, int k) { return tc_desc<H, AMN>(base, k); };
      auto bdesc = [](uint32_t base, int k) { return tc_desc<H, BMN>(base, k); };
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full[s], ph);
        mbar_wait(main_empty, (uint32_t)((kb & 1) ^ 1));       // drain of k-block kb-1 (runs under its correction MMAs)
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * SM::STAGE_BYTES);
        const uint32_t a_hi = sa, a_lo = sa + SM::A_BYTES, b_hi = sa + 2 * SM::A_BYTES, b_lo = b_hi + SM::B_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma<H>(tmem_base, adesc(a_hi, k), bdesc(b_hi, k), idesc, k > 0 ? 1u : 0u);
        tc_commit(main_full);
        if (!(e.debug & 4))
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          tc_mma<H>(tmem_corr, adesc(a_lo, k), bdesc(b_hi, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          tc_mma<H>(tmem_corr, adesc(a_hi, k), bdesc(b_lo, k), idesc, 1u);
        }
        tc_commit(&empty[s]);
      }
      tc_commit(corr_full);
    }
    __syncwarp();
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int lg = warp & 3;                 // TMEM lane quadrant
    const int half = (warp - 4) >> 2;        // which 128 columns of the 256
    const uint32_t lane_off = (uint32_t)(lg * 32) << 16;
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.0f;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(main_full, (uint32_t)(kb & 1));
      tc_fence_after();
      if (!(e.debug & 2))
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        float v[32];
        tc_ld_32x32(tmem_base + lane_off + (uint32_t)(half * 128 + c), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[c + j] += v[j];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(main_empty);
    }
    mbar_wait(corr_full, 0);
    tc_fence_after();
    float* cs = reinterpret_cast<float*>(smem);
    constexpr int CS_LD = BN + 4;
    const float s1 = (H && e.a_inv) ? *e.a_inv : 1.0f;
    const float s2 = e.alpha * ((H && e.b_inv) ? *e.b_inv : 1.0f);
    {
      float* crow_s = cs + (lg * 32 + lane) * CS_LD + half * 128;
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        float v[32];
        tc_ld_32x32(tmem_corr + lane_off + (uint32_t)(half * 128 + c), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(crow_s + c + j, s2 * (s1 * (acc[c + j] + v[j])), s2 * (s1 * (acc[c + j + 1] + v[j + 1])),
                 s2 * (s1 * (acc[c + j + 2] + v[j + 2])), s2 * (s1 * (acc[c + j + 3] + v[j + 3])));
      }
    }
    float* s_colsum = cs + TC_BM * CS_LD;               // [BN] per-tile column sums, behind the staging tile
    const int et = threadIdx.x - 128;                   // 0..255 within the epilogue warps
    if (e.colsum) s_colsum[et] = 0.0f;
    asm volatile("bar.sync 1, 256;" ::: "memory");      // the 8 epilogue warps only
    if (!(e.debug & 1)) {
      if (epilogue_fast_ok(e, m0, n0, BN, H) && !(e.debug & 256)) epilogue_fast<H, 8, 16>(e, cs, CS_LD, s_colsum, (warp - 4) * 16, m0, n0, lane);
      else epilogue_rows<H>(e, cs, CS_LD, s_colsum, (warp - 4) * 16, 16, BN, m0, n0, lane);
    }
    if (e.colsum && !e.accumulate) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (n0 + et < e.N) atomicAdd(e.colsum + n0 + et, s_colsum[et]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------ operand prep

// split src[rows, cols] (ld) into zero-padded hi/lo planes [rows_p, cols_p]
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols, int rows_p, int cols_p, float* __restrict__ hi,
               float* __restrict__ lo) {
  const int64_t total = (int64_t)rows_p * cols_p;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols_p), c = (int)(i - (int64_t)r * cols_p);
    float h = 0.0f, l = 0.0f;
    if (r < rows && c < cols) split_tf32(src[(int64_t)r * ld + c], h, l);
    hi[i] = h; lo[i] = l;
  }
}

// FP16 format: max |src| over the [rows, cols] view -> atomicMax on the uint bits (non-negative floats order like uints)
__global__ void __launch_bounds__(256)
tc_amax_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols, unsigned* __restrict__ amax) {
  const int64_t total = (int64_t)rows * cols;
  float m = 0.0f;
  if (cols == ld) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(src[i]));
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
      m = fmaxf(m, fabsf(src[(int64_t)r * ld + c]));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sm[w]);
    if (m > 0.0f) atomicMax(amax, __float_as_uint(m));
  }
}

// FP16 format: split src[rows, cols] (ld) into zero-padded half planes [rows_p, cols_p]; 4 consecutive columns per thread
// (cols_p is a multiple of 8).  Two modes, both without a host sync:
//   amax_in != null : EXACT  -- the scale is derived from the tensor's max (already in device memory) and published with its
//                              inverse in scale_io[0..1] for the epilogues of the GEMMs that consume these planes;
//   amax_in == null : PREDICTED -- scale_io holds the scale derived from the previous call's max at this site; this pass tracks
//                              the current max into amax_out for the next call and raises flag bit 0 if a value does not fit.
__global__ void __launch_bounds__(256)
tc_prep_h_kernel(const float* __restrict__ src, int64_t ld, int rows, int cols, int rows_p, int cols_p, __half* __restrict__ hi,
                 __half* __restrict__ lo, const unsigned* __restrict__ amax_in, float* __restrict__ scale_io, float* __restrict__ scale_copy,
                 unsigned* __restrict__ amax_out, unsigned* __restrict__ flag, int top) {
  float s;
  if (amax_in) {
    s = scale_from_amax(__uint_as_float(*amax_in), top);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scale_io[0] = s; scale_io[1] = 1.0f / s; }
  } else s = scale_io[0];
  if (scale_copy && blockIdx.x == 0 && threadIdx.x == 0) { scale_copy[0] = s; scale_copy[1] = (s != 0.0f) ? 1.0f / s : 0.0f; }
  const int c4n = cols_p >> 2;
  const int64_t total = (int64_t)rows_p * c4n;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4n), c = (int)(i - (int64_t)r * c4n) * 4;
    float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (r < rows) {
      const float* sp = src + (int64_t)r * ld + c;
      if (vec && c + 4 <= cols) { const float4 t = *reinterpret_cast<const float4*>(sp); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w; }
      else for (int j = 0; j < 4; ++j) if (c + j < cols) x[j] = sp[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) m = fmaxf(m, fabsf(x[j]));
    uint2 hv, lv;
    split_f16x2(x[0] * s, x[1] * s, hv.x, lv.x); split_f16x2(x[2] * s, x[3] * s, hv.y, lv.y);
    *reinterpret_cast<uint2*>(hi + (int64_t)r * cols_p + c) = hv;
    *reinterpret_cast<uint2*>(lo + (int64_t)r * cols_p + c) = lv;
  }
  if (!amax_in) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) {
      if (amax_out) atomicMax(amax_out, __float_as_uint(m));
      if (flag && !(m * s <= 60000.0f)) atomicOr(flag, 1u);
      if (flag && s == 0.0f) atomicOr(flag, 2u);       // the site only ever saw an all-zero tensor, now there is data
    }
  }
}

// FP16 format: all weight tensors of the learner in ONE launch (they are re-split after every optimizer step; one launch
// per tensor cost ~7 us each for a few MB of work).  blockIdx.y = tensor, blockIdx.x = slice of it.
__global__ void __launch_bounds__(256)
tc_amax_batch_kernel(TcPrepBatch b, unsigned* __restrict__ amax) {
  const TcPrepItem it = b.item[blockIdx.y];
  const int64_t total = (int64_t)it.rows * it.cols;
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(it.src[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(amax + it.site, __float_as_uint(m));
}
__global__ void __launch_bounds__(256)
tc_prep_h_batch_kernel(TcPrepBatch b, unsigned* __restrict__ amax, float* __restrict__ scale, float* __restrict__ bscale, int exact,
                       unsigned* __restrict__ flag) {
  const TcPrepItem it = b.item[blockIdx.y];
  float s;
  if (exact) {
    s = scale_from_amax(__uint_as_float(amax[it.site]), TOP_SITE);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scale[2 * it.site] = s; scale[2 * it.site + 1] = 1.0f / s; }
  } else s = scale[2 * it.site];
  if (blockIdx.x == 0 && threadIdx.x == 0) { bscale[2 * it.buf] = s; bscale[2 * it.buf + 1] = (s != 0.0f) ? 1.0f / s : 0.0f; }
  __half* hi = (__half*)it.hi; __half* lo = (__half*)it.lo;
  const int c4n = it.ldp >> 2;
  const int64_t total = (int64_t)it.rows * c4n;
  const bool vec = (it.cols & 3) == 0;         // weight rows are contiguous (ld = cols) and the arena offsets 128-byte aligned
  float m = 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4n), c = (int)(i - (int64_t)r * c4n) * 4;
    float x[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const float* sp = it.src + (int64_t)r * it.cols + c;
    if (vec && c + 4 <= it.cols) { const float4 t = *reinterpret_cast<const float4*>(sp); x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w; }
    else for (int j = 0; j < 4; ++j) if (c + j < it.cols) x[j] = sp[j];
#pragma unroll
    for (int j = 0; j < 4; ++j) m = fmaxf(m, fabsf(x[j]));
    uint2 hv, lv;
    split_f16x2(x[0] * s, x[1] * s, hv.x, lv.x); split_f16x2(x[2] * s, x[3] * s, hv.y, lv.y);
    *reinterpret_cast<uint2*>(hi + (int64_t)r * it.ldp + c) = hv;
    *reinterpret_cast<uint2*>(lo + (int64_t)r * it.ldp + c) = lv;
  }
  if (!exact) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) {
      atomicMax(amax + it.site, __float_as_uint(m));
      if (!(m * s <= 60000.0f)) atomicOr(flag, 1u);
      if (s == 0.0f) atomicOr(flag, 2u);
    }
  }
}

// FP16 format, start of every top-level call: fold the maxima tracked during the previous call into the sites' scales (the
// prediction for this call), check that the previous prediction did not lose precision, and clear the maxima.
__global__ void __launch_bounds__(256)
tc_site_update_kernel(unsigned* __restrict__ amax, float* __restrict__ scale, int n, float* __restrict__ static_scale, float static_value,
                      unsigned* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { static_scale[0] = static_value; static_scale[1] = 1.0f / static_value; }
  if (i >= n) return;
  const float a = __uint_as_float(amax[i]);
  const float s0 = scale[2 * i];
  if (a > 0.0f) {
    if (s0 != 0.0f && a * s0 < 0.015625f) atomicOr(flag, 2u);        // the tensor shrank by > 2^12 between two calls: the split lost bits
    const float s = scale_from_amax(a, TOP_SITE);
    scale[2 * i] = s; scale[2 * i + 1] = 1.0f / s;
    amax[i] = 0u;
  }                                          // a site that only ever saw all-zero tensors keeps scale 0 (zero planes, zero inverse)
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// cuTensorMapEncodeTiled costs microseconds and the learner re-issues the same ~230 maps every minibatch: memoise.
struct MapKey {
  const void* base; int rows, cols; int64_t ld; int box_rows; int mn;     // mn: bit 0 = MN-major, bit 1 = half planes
  bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && mn == o.mn; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.base;
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (size_t)k.ld * 0x165667B19E3779F9ull + (size_t)k.box_rows * 31 + (size_t)k.mn;
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() { static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c; return c; }

// 2D map over [rows, cols] elements (cols contiguous, row stride ld elements); the box is one 128-byte row chunk
// (32 fp32 words / 64 halfs) x box_rows
static int encode_cached(CUtensorMap* tm, const void* base, int rows, int cols, int64_t ld, int box_rows, bool mn_major, bool half = false) {
  MapKey k{base, rows, cols, ld, box_rows, (mn_major ? 1 : 0) | (half ? 2 : 0)};
  auto& c = map_cache();
  auto it = c.find(k);
  if (it != c.end()) { memcpy(tm, &it->second, sizeof(CUtensorMap)); return ASE_OK; }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return ASE_ERR_UNSUPPORTED; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * (half ? 2 : 4)};
  cuuint32_t box[2] = {half ? 64u : 32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  // MN-major: fp32 words need the 32-byte-atom flavour of the 128B swizzle, halfs the plain one (see the smem descriptors)
  CUresult r = enc(tm, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, (mn_major && !half) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%d x %d, ld %lld, box %d) failed with CUresult %d", rows, cols, (long long)ld, box_rows, (int)r); return ASE_ERR_CUDA; }
  if (c.size() > 8192) c.clear();
  c.emplace(k, *tm);
  return ASE_OK;
}

// 2D tensor map over a zero-padded plane [rows_p, cols_p] (cols contiguous); box = [box_rows x 32 cols], 128B swizzle
static int make_map(CUtensorMap* tm, const void* base, int rows_p, int cols_p, int box_rows, bool mn_major = false, bool half = false) {
  return encode_cached(tm, base, rows_p, cols_p, cols_p, box_rows, mn_major, half);
}

static inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// planes are padded to whole tiles in both dimensions (zero filled by the prep kernel): no reliance on TMA OOB fill
// (the FP16 format needs half of it: [*, pad(K, 64)] halfs <= [*, pad(K, 32)] words; the first TC_WS_HEAD bytes hold the
// amax / scale slots of a registry-less call)
constexpr int64_t TC_WS_HEAD = 1024;
int64_t gemm_tc_workspace_bytes(int M, int N, int K) {
  const int64_t Mp = pad_to(M, 128), Np = pad_to(N, 128), Kp = pad_to(K, TC_BK);
  return TC_WS_HEAD + 2 * align_up(Mp * Kp * 4, 1024) + 2 * align_up(Np * Kp * 4, 1024);
}

// every shape runs on the tensor cores (small heads are padded up to one tile)
bool gemm_tc_supported(const AseGemmParams& p) { return p.M >= 1 && p.N >= 1 && p.K >= 1; }

// a missing / small / misaligned workspace is an ERROR, never a silent fallback
int gemm_tc_check_workspace(const AseGemmParams& p) {
  if (!p.workspace || p.workspace_bytes < gemm_tc_workspace_bytes(p.M, p.N, p.K)) {
    set_error("tcgen05 GEMM %dx%dx%d: workspace %lld bytes < required %lld", p.M, p.N, p.K, (long long)p.workspace_bytes,
              (long long)gemm_tc_workspace_bytes(p.M, p.N, p.K));
    return ASE_ERR_WORKSPACE;
  }
  if (reinterpret_cast<uintptr_t>(p.workspace) & 1023) { set_error("tcgen05 GEMM: workspace must be 1024-byte aligned"); return ASE_ERR_WORKSPACE; }
  return ASE_OK;
}

// operand with `rows` = its M/N extent and reduction length K.  trans == 0: stored [rows, K] (K-major planes
// [rows_p, Kp]); trans == 1: stored [K, rows] (MN-major planes [Kp, rows_p]) -- no transposition anywhere.
static int prep_operand(const float* src, int64_t ld, int trans, int rows, int K, int rows_p, int Kp, float* hi, float* lo, cudaStream_t st) {
  const int pr = trans ? Kp : rows_p, pc = trans ? rows_p : Kp;
  const int sr = trans ? K : rows, sc = trans ? rows : K;
  const int64_t total = (int64_t)pr * pc;
  tc_prep_kernel<<<(int)imin64((total + 255) / 256, 148 * 16), 256, 0, st>>>(src, ld, sr, sc, pr, pc, hi, lo);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

static int launch_amax(const float* src, int64_t ld, int rows, int cols, unsigned* amax, cudaStream_t st) {
  const int64_t total = (int64_t)rows * cols;
  tc_amax_kernel<<<(int)imin64((total + 1023) / 1024, 148 * 8), 256, 0, st>>>(src, ld, rows, cols, amax);
  ASE_LAUNCH_OK();
  return ASE_OK;
}
static int launch_prep_h(const float* src, int64_t ld, int rows, int cols, int rows_p, int cols_p, void* hi, void* lo, const unsigned* amax_in,
                         float* scale_io, float* scale_copy, unsigned* amax_out, unsigned* flag, cudaStream_t st, int top = TOP_SITE) {
  const int64_t total = (int64_t)rows_p * (cols_p / 4);
  tc_prep_h_kernel<<<(int)imin64((total + 255) / 256, 148 * 16), 256, 0, st>>>(src, ld, rows, cols, rows_p, cols_p, (__half*)hi, (__half*)lo, amax_in,
                                                                              scale_io, scale_copy, amax_out, flag, top);
  ASE_LAUNCH_OK();
  return ASE_OK;
}
// FP16 format: materialise the planes of a tensor nobody split yet.  predicted: the site's scale from the previous call is used
// (one pass); else a max pass runs first (two passes, exact scale).
static int materialize_h(const float* src, int64_t ld, int sr, int sc, int pr, int pc, void* hi, void* lo, unsigned* amax, float* scale,
                         float* scale_copy, bool predicted, unsigned* flag, cudaStream_t st, int top = TOP_SITE) {
  if (predicted) return launch_prep_h(src, ld, sr, sc, pr, pc, hi, lo, nullptr, scale, scale_copy, amax, flag, st);
  int rc;
  if ((rc = launch_amax(src, ld, sr, sc, amax, st))) return rc;
  return launch_prep_h(src, ld, sr, sc, pr, pc, hi, lo, amax, scale, scale_copy, nullptr, nullptr, st, top);
}

// Optional per-launch timing of the main kernel (bench.py's live roofline measurement): CUDA events recorded on
// the launching stream around every gemm_tc_kernel launch; read back (with a sync) by ase_gemm_tc_profile_read.
struct TcProfile {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t used = 0;
  double flops = 0.0;
};
static TcProfile g_prof;

static void prof_mark(cudaStream_t st) {
  if (g_prof.used == g_prof.ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) { g_prof.on = false; return; }
    g_prof.ev.push_back(e);
  }
  cudaEventRecord(g_prof.ev[g_prof.used++], st);
}

bool tc_prof_on() { return g_prof.on; }
void tc_prof_mark(cudaStream_t st) { prof_mark(st); }
void tc_prof_add_flops(double f) { g_prof.flops += f; }

int tc_pdl() {   // env ASE_TC_PDL=0 launches the GEMMs fully stream-serialised
  static int v = -1;
  if (v < 0) { const char* d = getenv("ASE_TC_PDL"); v = d ? (atoi(d) != 0) : 1; }
  return v;
}

template <int BN, int STAGES, bool AMN, bool BMN, bool H>
static int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcEpi& e,
                     int splits, cudaStream_t st) {
  using SM = TcSmem<BN, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ASE_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, AMN, BMN, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  dim3 grid(ceil_div(e.N, BN), ceil_div(e.M, TC_BM), splits);
  const bool prof = g_prof.on;
  if (prof) prof_mark(st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl();
  cfg.attrs = attr; cfg.numAttrs = 1;
  ASE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, AMN, BMN, H>, ah, al, bh, bl, e));
  if (prof) { prof_mark(st); g_prof.flops += 2.0 * (double)e.M * (double)e.N * (double)e.K; }
  ASE_LAUNCH_OK();
  return ASE_OK;
}

template <int BN, int STAGES, bool H>
static int launch_tc_major(bool amn, bool bmn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                           const TcEpi& e, int splits, cudaStream_t st) {
  if (!amn && !bmn) return launch_tc<BN, STAGES, false, false, H>(ah, al, bh, bl, e, splits, st);
  if (!amn && bmn) return launch_tc<BN, STAGES, false, true, H>(ah, al, bh, bl, e, splits, st);
  if (amn && !bmn) return launch_tc<BN, STAGES, true, false, H>(ah, al, bh, bl, e, splits, st);
  return launch_tc<BN, STAGES, true, true, H>(ah, al, bh, bl, e, splits, st);
}

template <bool AMN, bool BMN, bool H>
static int launch_tc256(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcEpi& e,
                        int splits, cudaStream_t st) {
  using SM = TcSmem<TC256_BN, TC256_STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    ASE_CUDA_OK(cudaFuncSetAttribute(gemm_tc256_kernel<AMN, BMN, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  dim3 grid(ceil_div(e.N, TC256_BN), ceil_div(e.M, TC_BM), splits);
  const bool prof = g_prof.on;
  if (prof) prof_mark(st);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC256_THREADS); cfg.dynamicSmemBytes = SM::TOTAL; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = tc_pdl();
  cfg.attrs = attr; cfg.numAttrs = 1;
  ASE_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tc256_kernel<AMN, BMN, H>, ah, al, bh, bl, e));
  if (prof) { prof_mark(st); g_prof.flops += 2.0 * (double)e.M * (double)e.N * (double)e.K; }
  ASE_LAUNCH_OK();
  return ASE_OK;
}
template <bool H>
static int launch_tc256_major(bool amn, bool bmn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                              const TcEpi& e, int splits, cudaStream_t st) {
  if (!amn && !bmn) return launch_tc256<false, false, H>(ah, al, bh, bl, e, splits, st);
  if (!amn && bmn) return launch_tc256<false, true, H>(ah, al, bh, bl, e, splits, st);
  if (amn && !bmn) return launch_tc256<true, false, H>(ah, al, bh, bl, e, splits, st);
  return launch_tc256<true, true, H>(ah, al, bh, bl, e, splits, st);
}
int gemm_tc_tile_n(int N);
static int tc_tile256();
static int tc_pair() {      // env ASE_TC_PAIR=0 keeps the one-tile-per-CTA kernels (read per call: the tests compare both paths)
  const char* d = getenv("ASE_TC_PAIR");
  return d ? atoi(d) : 1;
}
bool gemm_tc_pair_candidate(int backend, int M, int N) { return backend == 2 && tc_pair() && tc_tile256() && N >= 384 && (M % 128 == 0) && (N % 8 == 0); }
static int tc_tile256() {   // env ASE_TC_TILE256=0 keeps every GEMM on the 128x128 kernel
  static int v = -1;
  if (v < 0) { const char* d = getenv("ASE_TC_TILE256"); v = d ? atoi(d) : 1; }
  return v;
}

int gemm_tc_tile_n(int N) { return (tc_tile256() && N >= 384) ? 256 : (N > 64 ? 128 : 64); }

// ---------------------------------------------------------------------------------------------------------
// Operand-plane registry: fp32 buffers whose TF32 hi/lo planes are kept next to them so that a tensor is split at
// most once (by the epilogue of the GEMM that produced it, or by one prep pass on first use) no matter how many
// GEMMs consume it, in either major-ness.  Host-side bookkeeping in stream-issue order (single stream).
// ---------------------------------------------------------------------------------------------------------
PlaneBuf* PlaneRegistry::find(const float* p) {
  for (int i = 0; i < n; ++i)
    if (p >= b[i].base && p < b[i].base + b[i].capacity) return &b[i];
  return nullptr;
}
void PlaneRegistry::add(const float* base, int64_t capacity, float* hi, float* lo, int64_t plane_capacity) {
  if (n >= MAX) return;
  PlaneBuf& x = b[n++];
  x.base = base; x.capacity = capacity; x.hi = hi; x.lo = lo; x.plane_capacity = plane_capacity;
  x.ld = 0; x.rows = x.cols = 0; x.ldp = 0; x.valid = false; x.scale_ptr = nullptr; x.amax_site = -1; x.is_static = false; x.fp32_stale = false;
}
PlaneBuf* PlaneRegistry::declare(const float* base, int64_t ld, int rows, int cols) {
  PlaneBuf* x = find(base);
  if (!x || x->base != base) return nullptr;
  x->amax_site = -1; x->is_static = false; x->fp32_stale = false;
  const int64_t ldp = f16 ? (cols + 7) / 8 * 8 : (cols + 3) / 4 * 4;
  if ((int64_t)rows * ldp > (f16 ? 2 : 1) * x->plane_capacity) { x->valid = false; return nullptr; }
  x->ld = ld; x->rows = rows; x->cols = cols; x->ldp = ldp; x->valid = true;
  if (f16) { x->is_static = true; x->scale_ptr = static_scale; }
  return x;
}
void* PlaneRegistry::plane(const PlaneBuf* x, bool lo, int64_t r0, int64_t c0) const {
  float* p = lo ? x->lo : x->hi;
  const int64_t off = r0 * x->ldp + c0;
  return f16 ? (void*)((__half*)p + off) : (void*)(p + off);
}
void PlaneRegistry::invalidate(const float* p) { if (PlaneBuf* x = find(p)) { x->valid = false; x->amax_site = -1; x->is_static = false; x->fp32_stale = false; } }
void PlaneRegistry::invalidate_range(const float* lo_, const float* hi_) {
  for (int i = 0; i < n; ++i) if (b[i].base >= lo_ && b[i].base < hi_) { b[i].valid = false; b[i].amax_site = -1; b[i].is_static = false; b[i].fp32_stale = false; }
}
int PlaneRegistry::begin_call(cudaStream_t st, int base) {
  call_base = base; gemm_index = 0;
  if (!f16) return ASE_OK;
  if (reset_pending) {      // new parameters: every scale is re-derived exactly on this call, nothing is compared with the old ones
    ASE_CUDA_OK(cudaMemsetAsync(amax, 0, (size_t)SITES * 12, st));      // amax[SITES] + scale[SITES][2] are contiguous
    reset_pending = false;
  }
  for (int i = 0; i < SITES; ++i) { if (touched[i]) { known[i] = true; touched[i] = false; } }
  for (int i = 0; i < n; ++i) {
    b[i].amax_site = -1;                                 // maxima tracked by GEMM epilogues are only meaningful within one call
    // planes written by a GEMM epilogue refer to their site's scale slot, which is re-predicted right now: drop them
    // (planes split by a prep pass -- the weights -- carry their own copy of the scale and stay valid)
    if (b[i].valid && !b[i].is_static && b[i].scale_ptr >= scale && b[i].scale_ptr < scale + 2 * SITES) b[i].valid = false;
    if (b[i].is_static) { b[i].valid = false; b[i].is_static = false; }
  }
  tc_site_update_kernel<<<ceil_div(SITES, 256), 256, 0, st>>>(amax, scale, SITES, static_scale, STATIC_SCALE, flag);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

// (Re)split every listed weight tensor [rows, cols] (contiguous) into its registered planes: one launch (two on the first call
// after the parameters were announced: exact maxima first).  Sites are fixed per tensor (WEIGHT_SITE0 + i).
int PlaneRegistry::prep_weights(const float* const* src, const int* rows, const int* cols, int count, cudaStream_t st) {
  if (!f16 || count <= 0) return ASE_OK;
  TcPrepBatch batch; int nb = 0; bool all_known = true;
  for (int i = 0; i < count && nb < TcPrepBatch::MAX; ++i) {
    PlaneBuf* x = find(src[i]);
    if (!x || x->base != src[i]) continue;
    const int64_t ldp = (cols[i] + 7) / 8 * 8;
    if ((int64_t)rows[i] * ldp > 2 * x->plane_capacity) continue;
    const int site = WEIGHT_SITE0 + i;
    if (site >= SITES) break;
    TcPrepItem& it = batch.item[nb++];
    it.src = src[i]; it.hi = x->hi; it.lo = x->lo; it.rows = rows[i]; it.cols = cols[i]; it.ldp = (int)ldp; it.site = site; it.buf = (int)(x - b);
    x->ld = cols[i]; x->rows = rows[i]; x->cols = cols[i]; x->ldp = ldp; x->valid = true; x->is_static = false; x->amax_site = -1; x->fp32_stale = false;
    x->scale_ptr = bscale + 2 * it.buf;
    all_known = all_known && known[site];
    touched[site] = true;
  }
  if (nb == 0) return ASE_OK;
  dim3 grid(48, nb);      // 28 MB of weights per optimizer step: 16 slices per tensor ran at 1 TB/s (ncu r02: 56 us), 48 x 16 tensors fill the 148 SMs 5 deep
  if (!all_known) {
    tc_amax_batch_kernel<<<grid, 256, 0, st>>>(batch, amax);
    ASE_LAUNCH_OK();
  }
  tc_prep_h_batch_kernel<<<grid, 256, 0, st>>>(batch, amax, scale, bscale, all_known ? 0 : 1, flag);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

struct OpView { const void* hi; const void* lo; int64_t ldp; bool ok; const float* scale; };

// View [nat_rows, nat_cols] (ld) at `ptr` as planes.  Geometry of a registered buffer is whatever its last full
// writer declared; a first read of a buffer without valid planes splits the WHOLE declared buffer once.
// site: this reader's scale site (FP16 format; used when the buffer's max is not already tracked at its producer's site).
static int resolve_operand(PlaneRegistry* reg, const float* ptr, int64_t ld, int nat_rows, int nat_cols, OpView* v, int site, cudaStream_t st) {
  v->ok = false; v->scale = nullptr;
  if (!reg) return ASE_OK;
  PlaneBuf* x = reg->find(ptr);
  if (!x) return ASE_OK;
  const bool H = reg->f16;
  const int pal = H ? 8 : 4;                 // plane leading dimensions / column offsets: 16-byte granules
  if (!x->valid) {
    if (x->fp32_stale) { set_error("tcgen05 GEMM: a planes-only tensor lost its planes before it was consumed"); return ASE_ERR_INVALID; }
    // adopt the reader's geometry if it starts at the buffer base (inputs written by non-GEMM kernels, weights)
    if (ptr != x->base) return ASE_OK;
    if (!H) {
      const int64_t ldp = pad_to(nat_cols, 4);
      if ((int64_t)nat_rows * ldp > x->plane_capacity || (int64_t)(nat_rows - 1) * ld + nat_cols > x->capacity) return ASE_OK;
      x->ld = ld; x->rows = nat_rows; x->cols = nat_cols; x->ldp = ldp;
      const int64_t total = (int64_t)nat_rows * ldp;
      tc_prep_kernel<<<(int)imin64((total + 255) / 256, 148 * 16), 256, 0, st>>>(ptr, ld, nat_rows, nat_cols, nat_rows, (int)ldp, x->hi, x->lo);
      ASE_LAUNCH_OK();
    } else {
      int rc;
      float* bs = reg->bscale + 2 * (x - reg->b);      // the buffer's own copy of the scale: survives the next begin_call
      if (x->amax_site >= 0) {
        // written earlier in this call by a GEMM whose scale was not known yet: it declared the geometry and tracked max |C|
        const int ts = x->amax_site;
        x->ldp = pad_to(x->cols, 8);
        if ((rc = launch_prep_h(ptr, x->ld, x->rows, x->cols, x->rows, (int)x->ldp, x->hi, x->lo, reg->amax + ts, reg->scale + 2 * ts, bs, nullptr, nullptr, st))) return rc;
      } else {
        // written by a non-GEMM kernel (or accumulated into): split the reader's view with this reader's site
        const int64_t ldp = pad_to(nat_cols, 8);
        if (site < 0 || (int64_t)nat_rows * ldp > 2 * x->plane_capacity || (int64_t)(nat_rows - 1) * ld + nat_cols > x->capacity) return ASE_OK;
        x->ld = ld; x->rows = nat_rows; x->cols = nat_cols; x->ldp = ldp;
        if ((rc = materialize_h(ptr, ld, nat_rows, nat_cols, nat_rows, (int)ldp, x->hi, x->lo, reg->amax + site, reg->scale + 2 * site, bs, reg->known[site],
                                reg->flag, st))) return rc;
        reg->touched[site] = true;
      }
      x->scale_ptr = bs; x->is_static = false;
    }
    x->valid = true;
  }
  if (ld != x->ld) return ASE_OK;
  const int64_t off = ptr - x->base;
  const int64_t r0 = off / x->ld, c0 = off - r0 * x->ld;
  if ((c0 & (pal - 1)) || r0 + nat_rows > x->rows || c0 + nat_cols > x->cols) return ASE_OK;
  v->hi = reg->plane(x, false, r0, c0); v->lo = reg->plane(x, true, r0, c0);
  if (H) v->scale = x->scale_ptr;
  v->ldp = x->ldp; v->ok = true;
  return ASE_OK;
}

// tensor map over a (sub-)view of a plane with TRUE extents: TMA zero-fills everything outside [rows, cols]
static int make_view_map(CUtensorMap* tm, const void* base, int rows, int cols, int64_t ldp, int box_rows, bool mn_major, bool half) {
  return encode_cached(tm, base, rows, cols, ldp, box_rows, mn_major, half);
}

int gemm_tc(const AseGemmParams& p, cudaStream_t st, PlaneRegistry* reg) {
  const bool H = p.backend == 2;             // scaled FP16 hi/lo planes instead of TF32 ones
  if (reg && reg->f16 != H) { set_error("tcgen05 GEMM: plane registry format does not match backend %d", p.backend); return ASE_ERR_INVALID; }
  const int BK = H ? 64 : 32;
  const int BN = (p.N > 64) ? 128 : 64;
  const bool use256 = tc_tile256() && p.N >= 384;               // 128x256 tiles (B maps keep 128-row boxes: two per stage)
  // persistent CTA-pair kernel (gemm_tc2.cu): the default for every wide GEMM whose rows come in whole 128-row tiles; its tensor maps
  // are those of the 128x256 kernel (128-row K-major boxes, 64x64 MN-major boxes), so the final choice can wait for the epilogue block
  const bool pair_shape = H && tc_pair() && use256 && (p.M % 128 == 0) && (p.N % 8 == 0);
  const int a_box = TC_BM;
  const int Mp = pad_to(p.M, 128), Np = pad_to(p.N, 128), Kp = pad_to(p.K, BK);
  int rc;
  // ---- operands: cached planes when the buffer is registered, else a split pass into the shared workspace
  OpView va, vb;
  const int a_rows = p.a_trans ? p.K : p.M, a_cols = p.a_trans ? p.M : p.K;
  const int b_rows = p.b_trans ? p.K : p.N, b_cols = p.b_trans ? p.N : p.K;
  // FP16 format: scale sites of this GEMM (A, B, C); without a registry the workspace head holds two transient slots
  const int site_a = (H && reg) ? reg->site(0) : -1, site_b = (H && reg) ? reg->site(1) : -1, site_c = (H && reg) ? reg->site(2) : -1;
  if (H && reg) { if (site_c < 0) { set_error("tcgen05 FP16 GEMM: more GEMMs in one call than scale sites"); return ASE_ERR_WORKSPACE; } reg->gemm_index++; }
  if ((rc = resolve_operand(reg, p.A, p.lda, a_rows, a_cols, &va, site_a, st))) return rc;
  if ((rc = resolve_operand(reg, p.B, p.ldb, b_rows, b_cols, &vb, site_b, st))) return rc;
  if (reg) {      // a planes-only tensor (fp32 store elided) can only be consumed through its planes / activity bits: anything else is a bug, not a fallback
    auto stale = [&](const float* ptr) { PlaneBuf* x = ptr ? reg->find(ptr) : nullptr; return x && x->fp32_stale; };
    if ((!va.ok && stale(p.A)) || (!vb.ok && stale(p.B)) || (p.mask_src && p.mask_mode && !(p.mask_mode == 1 && p.mask_bits) && stale(p.mask_src))) {
      set_error("tcgen05 GEMM %dx%dx%d: an operand's fp32 copy was elided (c_planes_only) but it is not consumed through its planes", p.M, p.N, p.K);
      return ASE_ERR_INVALID;
    }
  }
  CUtensorMap ah, al, bh, bl;
  char* ws = (char*)p.workspace;
  if (!va.ok || !vb.ok) { if ((rc = gemm_tc_check_workspace(p))) return rc; }
  unsigned* t_amax[2] = {nullptr, nullptr}; float* t_scale[2] = {nullptr, nullptr}; bool t_pred[2] = {false, false};
  if (H && (!va.ok || !vb.ok)) {
    if (reg) {
      const int sites[2] = {site_a, site_b};
      for (int i = 0; i < 2; ++i) {
        if (i == 0 ? va.ok : vb.ok) continue;
        t_amax[i] = reg->amax + sites[i]; t_scale[i] = reg->scale + 2 * sites[i]; t_pred[i] = reg->known[sites[i]]; reg->touched[sites[i]] = true;
      }
    } else {
      ASE_CUDA_OK(cudaMemsetAsync(ws, 0, 2 * sizeof(unsigned), st));
      t_amax[0] = (unsigned*)ws; t_amax[1] = (unsigned*)ws + 1; t_scale[0] = (float*)(ws + 16); t_scale[1] = (float*)(ws + 32);
    }
  }
  unsigned* oflag = (H && reg) ? reg->flag : nullptr;
  char* wsp = ws + TC_WS_HEAD;
  if (va.ok) {
    if ((rc = make_view_map(&ah, va.hi, a_rows, a_cols, va.ldp, p.a_trans ? BK : a_box, p.a_trans != 0, H)) ||
        (rc = make_view_map(&al, va.lo, a_rows, a_cols, va.ldp, p.a_trans ? BK : a_box, p.a_trans != 0, H))) return rc;
  } else {
    float* Ahi = (float*)wsp; float* Alo = (float*)(wsp + align_up((int64_t)Mp * Kp * 4, 1024));
    if (!H) { if ((rc = prep_operand(p.A, p.lda, p.a_trans, p.M, p.K, Mp, Kp, Ahi, Alo, st))) return rc; }
    else {
      if ((rc = materialize_h(p.A, p.lda, a_rows, a_cols, p.a_trans ? Kp : Mp, p.a_trans ? Mp : Kp, Ahi, Alo, t_amax[0], t_scale[0], nullptr, t_pred[0], oflag, st, reg ? TOP_SITE : TOP_EXACT))) return rc;
      va.scale = t_scale[0];
    }
    if (!p.a_trans) { if ((rc = make_map(&ah, Ahi, Mp, Kp, a_box, false, H)) || (rc = make_map(&al, Alo, Mp, Kp, a_box, false, H))) return rc; }
    else            { if ((rc = make_map(&ah, Ahi, Kp, Mp, BK, true, H)) || (rc = make_map(&al, Alo, Kp, Mp, BK, true, H))) return rc; }
  }
  if (vb.ok) {
    if ((rc = make_view_map(&bh, vb.hi, b_rows, b_cols, vb.ldp, p.b_trans ? BK : BN, p.b_trans != 0, H)) ||
        (rc = make_view_map(&bl, vb.lo, b_rows, b_cols, vb.ldp, p.b_trans ? BK : BN, p.b_trans != 0, H))) return rc;
  } else {
    char* wb = wsp + 2 * align_up((int64_t)Mp * Kp * 4, 1024);
    float* Bhi = (float*)wb; float* Blo = (float*)(wb + align_up((int64_t)Np * Kp * 4, 1024));
    if (!H) { if ((rc = prep_operand(p.B, p.ldb, p.b_trans, p.N, p.K, Np, Kp, Bhi, Blo, st))) return rc; }
    else {
      if ((rc = materialize_h(p.B, p.ldb, b_rows, b_cols, p.b_trans ? Kp : Np, p.b_trans ? Np : Kp, Bhi, Blo, t_amax[1], t_scale[1], nullptr, t_pred[1], oflag, st, reg ? TOP_SITE : TOP_EXACT))) return rc;
      vb.scale = t_scale[1];
    }
    if (!p.b_trans) { if ((rc = make_map(&bh, Bhi, Np, Kp, BN, false, H)) || (rc = make_map(&bl, Blo, Np, Kp, BN, false, H))) return rc; }
    else            { if ((rc = make_map(&bh, Bhi, Kp, Np, BK, true, H)) || (rc = make_map(&bl, Blo, Kp, Np, BK, true, H))) return rc; }
  }
  TcEpi e;
  e.C = p.C; e.ldc = p.ldc; e.M = p.M; e.N = p.N; e.K = p.K; e.alpha = p.alpha; e.bias = p.bias; e.act = p.act;
  e.mask_src = p.mask_src; e.ldm = p.ldm; e.mask_mode = p.mask_src ? p.mask_mode : 0; e.accumulate = p.accumulate;
  e.Chi = e.Clo = nullptr; e.ldp = 0; e.colsum = p.colsum_out;
  e.a_inv = (H && va.scale) ? va.scale + 1 : nullptr; e.b_inv = (H && vb.scale) ? vb.scale + 1 : nullptr;
  e.c_scale = nullptr; e.c_amax = nullptr; e.flag = nullptr;
  e.relu_bits = p.relu_bits_out; e.ldrb = p.ldrb; e.mask_bits = (e.mask_mode == 1) ? p.mask_bits : nullptr; e.ldmb = p.ldmb; e.skip_c = 0;
  // ---- output planes: a full write at the base of a registered buffer (re)declares its geometry; a partial write
  // keeps planes in sync only if they are currently valid with the same leading dimension; accumulation invalidates
  if (reg && !H) {
    if (PlaneBuf* x = reg->find(p.C)) {
      if (p.accumulate) x->valid = false;
      else if (p.C == x->base && (int64_t)p.M * pad_to(p.N, 4) <= x->plane_capacity && !(x->valid && x->ld == p.ldc && (x->rows > p.M || x->cols > p.N))) {
        x->ld = p.ldc; x->rows = p.M; x->cols = p.N; x->ldp = pad_to(p.N, 4); x->valid = true;
        e.Chi = x->hi; e.Clo = x->lo; e.ldp = x->ldp;
      } else if (x->valid && x->ld == p.ldc) {
        const int64_t off = p.C - x->base, r0 = off / x->ld, c0 = off - r0 * x->ld;
        if (r0 + p.M <= x->rows && c0 + p.N <= x->cols) { e.Chi = x->hi + r0 * x->ldp + c0; e.Clo = x->lo + r0 * x->ldp + c0; e.ldp = x->ldp; }
        else x->valid = false;
      } else x->valid = false;
    }
  } else if (reg && H) {
    // FP16 format.  A full write at the buffer base: the epilogue tracks max |C| at this GEMM's C site; if the site's scale is
    // already known (predicted from the previous call) it also writes the planes, else the first consumer splits with the exact
    // scale.  A tanh-bounded partial write into statically scaled planes (the style columns behind the normalised
    // observations) keeps them in sync.  Anything else leaves the buffer without planes.
    if (PlaneBuf* x = reg->find(p.C)) {
      if (x->fp32_stale && p.accumulate) { set_error("tcgen05 GEMM: accumulating into a planes-only tensor"); return ASE_ERR_INVALID; }
      x->fp32_stale = false;
      const bool full = !p.accumulate && p.C == x->base && (int64_t)p.M * pad_to(p.N, 8) <= 2 * x->plane_capacity && (int64_t)(p.M - 1) * p.ldc + p.N <= x->capacity &&
                        !(x->valid && x->is_static && x->ld == p.ldc && (x->rows > p.M || x->cols > p.N));
      if (full) {
        x->ld = p.ldc; x->rows = p.M; x->cols = p.N; x->ldp = pad_to(p.N, 8); x->is_static = false;
        reg->touched[site_c] = true; e.c_amax = reg->amax + site_c;
        if (reg->known[site_c]) {
          x->valid = true; x->amax_site = -1; x->scale_ptr = reg->scale + 2 * site_c;
          e.Chi = x->hi; e.Clo = x->lo; e.ldp = x->ldp; e.c_scale = x->scale_ptr; e.flag = reg->flag;
          if (p.c_planes_only) { e.skip_c = 1; x->fp32_stale = true; }
        } else { x->valid = false; x->amax_site = site_c; }
      } else if (!p.accumulate && x->valid && x->is_static && x->ld == p.ldc && p.act == 2 && !p.mask_src) {
        const int64_t off = p.C - x->base, r0 = off / x->ld, c0 = off - r0 * x->ld;
        if (r0 + p.M <= x->rows && c0 + p.N <= x->cols) { e.Chi = reg->plane(x, false, r0, c0); e.Clo = reg->plane(x, true, r0, c0); e.ldp = x->ldp; e.c_scale = x->scale_ptr; e.flag = reg->flag; }
        else { x->valid = false; x->amax_site = -1; x->is_static = false; }
      } else { x->valid = false; x->amax_site = -1; x->is_static = false; }
    }
  }
  e.kb_total = Kp / BK;
  { static int dbg = -1; if (dbg < 0) { const char* d = getenv("ASE_TC_DEBUG"); dbg = d ? atoi(d) : 0; } e.debug = dbg; }
  int splits = (p.accumulate && p.split_k > 1) ? p.split_k : 1;
  splits = min(splits, e.kb_total);
  e.kb_per_split = ceil_div(e.kb_total, splits);
  splits = ceil_div(e.kb_total, e.kb_per_split);
  const bool amn = p.a_trans != 0, bmn = p.b_trans != 0;
  if (H) {
    if (pair_shape && gemm_tc2_epilogue_ok(e)) return launch_tc2(amn, bmn, ah, al, bh, bl, e, splits, st);
    if (use256) return launch_tc256_major<true>(amn, bmn, ah, al, bh, bl, e, splits, st);
    if (BN == 128) return launch_tc_major<128, 3, true>(amn, bmn, ah, al, bh, bl, e, splits, st);
    return launch_tc_major<64, 2, true>(amn, bmn, ah, al, bh, bl, e, splits, st);      // 2 stages = 97 KB: two CTAs per SM hide each other's latencies
  }
  if (use256) return launch_tc256_major<false>(amn, bmn, ah, al, bh, bl, e, splits, st);
  if (BN == 128) return launch_tc_major<128, 3, false>(amn, bmn, ah, al, bh, bl, e, splits, st);
  return launch_tc_major<64, 4, false>(amn, bmn, ah, al, bh, bl, e, splits, st);
}

}  // namespace ase

extern "C" int ase_gemm_tc_profile(int enable) {
  ase::g_prof.on = enable != 0;
  ase::g_prof.used = 0;
  ase::g_prof.flops = 0.0;
  return ASE_OK;
}

extern "C" int ase_gemm_tc_profile_read(double* total_ms, int64_t* launches, double* flops) {
  using namespace ase;
  double tot = 0.0;
  const size_t n = g_prof.used / 2;
  for (size_t i = 0; i < n; ++i) {
    float ms = 0.0f;
    ASE_CUDA_OK(cudaEventSynchronize(g_prof.ev[2 * i + 1]));
    ASE_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int64_t)n;
  if (flops) *flops = g_prof.flops;
  return ASE_OK;
}
