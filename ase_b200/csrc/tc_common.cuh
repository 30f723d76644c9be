// Shared device-side pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one-tile-per-CTA kernels; gemm_tc2.cu: persistent CTA-pair kernel):
// PTX wrappers (mbarrier, TMA, tcgen05), shared-memory / instruction descriptors, the hi/lo split helpers and the epilogue parameter block.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "kernels.h"

namespace ase {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                 // fp32 elements per k-block = one 128-byte swizzle row
constexpr int TC_THREADS = 192;

// Operand-plane formats.  H = false: TF32 hi/lo planes stored as fp32 words (kind::tf32, K = 8 per MMA).
// H = true: SCALED FP16 hi/lo planes (kind::f16, K = 16 per MMA, twice the MMA rate and half the plane bytes): each tensor is
// multiplied by a per-tensor power of two that puts its max |x| in [2^8, 2^9) before the split, so hi + lo carries 22
// significant bits for every element within 2^-22 of the tensor max (absolute floor 2^-25 / scale); the epilogue multiplies
// by the two inverse scales (exact).  In both formats a k-block is ONE 128-byte swizzle row per operand row, so the shared
// memory tiles, the TMA transaction bytes and the 4-MMAs-per-k-block structure are identical.
template <bool H> struct TcFmt {
  static constexpr int BK = H ? 64 : 32;              // elements per k-block (128 bytes)
  static constexpr int MN_BOX = H ? 64 : 32;          // MN elements per MN-major TMA box row (128 bytes)
  static constexpr int MN_BOX_BYTES = H ? 8192 : 4096;   // BK k-rows x 128 bytes
  static constexpr int MN_KSTEP = H ? 2048 : 1024;    // bytes between the MN-major k-slices of consecutive MMAs (16 / 8 k-rows)
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// Programmatic dependent launch: the GEMMs are launched with programmatic stream serialization, so a CTA of the NEXT kernel may
// be scheduled (on an SM the previous kernel no longer needs) and run its prologue -- barrier init, TMEM allocation, descriptor
// fetch -- while the previous kernel's last wave is still computing.  launch_dependents lets our own successor do the same;
// wait blocks until the predecessor grid has completed and its writes are visible: nothing before it touches global memory.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
template <bool H>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (H) tc_mma_f16(tmem_d, adesc, bdesc, idesc, accum); else tc_mma_tf32(tmem_d, adesc, bdesc, idesc, accum);
}
// 32 lanes x 32 consecutive fp32 columns: thread `lane` receives row (lane_base + lane), columns [col, col+32)
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte swizzle: 8-row x 128 B atoms, 1024 B between 8-row groups.
//   bits [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4 (=64)
//   bits [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// MN-major (operand stored with the M/N index contiguous, reduction index strided).  For 32-bit operands the only
// MN-major layout the tensor core accepts is "128B swizzle with a 32-byte atom" (layout type 1, Swizzle<2,5,2>:
// 32-byte chunks XORed with the row index mod 4; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -- see
// cutlass/gemm/collective/builders/sm100_common.inl "for mn-major tf32 operands, SW128_32B is the only available smem
// layout".  The tile is (BM or BN)/32 boxes of [32 k-rows x 32 mn] = 4 KB each, box b at +4096*b; a K=8 MMA slice is
// 8 rows = two 4-row swizzle groups.  LBO = byte distance between 32-wide MN blocks (4096), SBO = distance between
// 4-row k groups (512).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

// MN-major 16-bit operands: the canonical SWIZZLE_128B layout (layout type 2), ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units
// (cute/atom/mma_traits_sm100.hpp): a TMA box is [64 k-rows x 64 mn] = 8 KB with the 16-byte chunks of row r XORed with r mod 8;
// LBO = distance between 64-wide MN blocks (one box, 8192), SBO = distance between 8-row k groups (1024); a K=16 MMA slice is
// two k groups.
__device__ __forceinline__ uint64_t make_smem_desc_mn_h(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
template <bool H, bool MN>
__device__ __forceinline__ uint64_t tc_desc(uint32_t base, int k) {
  if (!MN) return make_smem_desc(base + k * 32);                       // K-major: 32 bytes along the swizzled row per MMA
  return H ? make_smem_desc_mn_h(base + k * TcFmt<true>::MN_KSTEP) : make_smem_desc_mn(base + k * TcFmt<false>::MN_KSTEP);
}
// instruction descriptor: D = f32 (bits 4-5 = 1); A/B format at bits 7-9 / 10-12 (kind::tf32: 2 = TF32; kind::f16: 0 = F16);
// bit 15 / 16 = A / B MN-major; N >> 3 at 17, M >> 4 at 24
template <bool H>
__device__ __forceinline__ uint32_t tc_idesc(bool amn, bool bmn, int bn) {
  const uint32_t fmt = H ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((amn ? 1u : 0u) << 15) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(bn >> 3) << 17) |
         ((uint32_t)(TC_BM >> 4) << 24);
}

// scale that puts amax into [2^(top-1), 2^top); 1 for an all-zero / non-finite tensor.  Absolute split error 2^-25 / scale
// = 2^-(24+top) of the tensor max.  TOP_SITE leaves 2^7 of headroom below the FP16 maximum for scales PREDICTED from the
// previous call's max; TOP_EXACT is for a scale derived from the very tensor being split.
constexpr int TOP_SITE = 9, TOP_EXACT = 13;
__device__ __forceinline__ float scale_from_amax(float amax, int top) {
  if (!(amax > 0.0f) || !(amax < 3.0e38f)) return 1.0f;
  int e; frexpf(amax, &e);                      // amax = m * 2^e, m in [0.5, 1)
  return ldexpf(1.0f, max(-100, min(100, top - e)));
}
// two elements at once: cvt.rn.f16x2.f32 (F2FP, full rate) instead of two scalar F2F conversions (quarter-rate pipe; ncu r01:
// the store phase stalled on MIO at the F2Fs); the residuals are exact in fp32.  Returns the packed hi / lo half2 words.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// explicit shared-space accesses for the staging tile (a generic pointer makes the compiler emit LD.E / ST.E with their longer latency)
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ void sts128(float* p, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void split_f16(float xs, __half& hi, __half& lo) {
  hi = __float2half_rn(xs);
  lo = __float2half_rn(xs - __half2float(hi));   // the residual is exact in fp32
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float r = x - hi;                       // exact
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

struct TcEpi {
  float* C; int64_t ldc;
  int M, N, K;
  int kb_total, kb_per_split;
  float alpha;
  const float* bias;
  int act;
  const float* mask_src; int64_t ldm; int mask_mode;
  int accumulate;
  float* colsum;                           // optional: colsum[n] += sum_m C[m,n] (bias gradient of the layer whose dZ this GEMM produces)
  void* Chi; void* Clo; int64_t ldp;       // optional hi/lo planes of C (operand cache for the consumers of C): fp32 words (TF32) or halfs
  const float* a_inv; const float* b_inv;  // FP16 planes: device pointers to the operands' inverse scales (null = 1)
  const float* c_scale;                    // FP16 planes of C: device pointer to the scale they are written with
  unsigned* c_amax;                        // optional: atomicMax of |C| (as uint bits) -- the scale source for the consumers of C
  unsigned* flag;                          // FP16 planes of C written with a PREDICTED scale: sticky overflow flag (bit 0)
  uint32_t* relu_bits; int64_t ldrb;       // optional: bit (n % 32) of word n / 32 of row m := C[m,n] > 0
  const uint32_t* mask_bits; int64_t ldmb; // optional: replaces mask_src for mask_mode 1 (same bit layout)
  int skip_c;                              // the planes are the only consumers of C: no fp32 store
  int debug;   // experiments only (env ASE_TC_DEBUG): 1 skip the whole store phase, 2 skip TMEM drain loads, 4 skip correction MMAs,
               // 16 skip mask loads, 32 skip plane stores, 64 skip column-sum atomics, 128 skip the fp32 C store
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------ CTA-pair (cta_group::2) wrappers
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa_u32)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair into ITS OWN shared memory; the transaction bytes are credited to the mbarrier at
// cluster address `bar_cluster` (the leader CTA's barrier: the leader issues the MMAs that read both CTAs' tiles)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
// tcgen05.mma over a CTA pair: M = 256 (128 rows per CTA), A / D split by rows, B split by N across the two CTAs' shared memory;
// issued by ONE thread of the leader CTA (cluster rank 0)
__device__ __forceinline__ void tc_mma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// completion of all prior tcgen05 ops of this thread -> one arrive on the barrier at this shared-memory offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit2_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// instruction descriptor of the pair MMA: M = 256, N = bn
__device__ __forceinline__ uint32_t tc_idesc2_f16(bool amn, bool bmn, int bn) {
  return (1u << 4) | ((amn ? 1u : 0u) << 15) | ((bmn ? 1u : 0u) << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

}  // namespace ase

// ------------------------------------------------------------------------------------------ host side (shared between the .cu files)
namespace ase {
// per-launch CUDA-event timing of the GEMM kernels (ase_gemm_tc_profile); defined in gemm_tc.cu
bool tc_prof_on();
void tc_prof_mark(cudaStream_t st);
void tc_prof_add_flops(double f);
int tc_pdl();
// persistent CTA-pair kernel (gemm_tc2.cu): FP16 planes, N >= 384, M % 128 == 0
bool gemm_tc2_epilogue_ok(const TcEpi& e);
int launch_tc2(bool amn, bool bmn, const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, const TcEpi& e,
               int splits, cudaStream_t st);
int gemm_tc2_pair_slots();          // CTA pairs that can be co-resident (74 on a full B200)

}  // namespace ase
