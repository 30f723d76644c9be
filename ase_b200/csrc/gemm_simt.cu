// fp32 SIMT GEMM with fused epilogue -- the exact-fp32 backend (backend 0) of ase_gemm and the fallback for
// shapes the tcgen05 backend does not take (tiny N heads, unaligned leading dimensions).
//   C[M,N] = epi(alpha * op(A).op(B));  128x128x8 block tile, 256 threads, 8x8 register tile (2x2 quads of 4x4 so
//   that shared-memory reads are conflict-free float4), register-prefetch double buffering, optional split-K
//   with fp32 RED accumulation.
#include "common.cuh"
#include "kernels.h"

namespace ase {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

struct SimtArgs {
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  float* C; int64_t ldc;
  int M, N, K;
  float alpha;
  const float* bias;
  int act;
  const float* mask_src; int64_t ldm; int mask_mode;
  int accumulate;
  float* colsum;
  int k_per_split;
  int vecA, vecB;    // 16-byte vector loads allowed (alignment of base + leading dimension)
};

// Loads one 128x8 operand tile slice owned by this thread into r[4].
//  TRANS=false: operand stored [rows, K] (k contiguous): thread -> row t/2, k-offset (t%2)*4
//  TRANS=true : operand stored [K, rows] (row index contiguous): thread -> k t/32, row-offset (t%32)*4
template <bool TRANS>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t ld, int rows, int K, int row0, int k0, int kend,
                                          int vec, float (&r)[4]) {
  const int t = threadIdx.x;
  if (!TRANS) {
    const int row = row0 + (t >> 1), k = k0 + (t & 1) * 4;
    r[0] = r[1] = r[2] = r[3] = 0.0f;
    if (row < rows) {
      const float* p = P + (int64_t)row * ld + k;
      if (vec && k + 3 < kend) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (k + i < kend) r[i] = p[i];
      }
    }
  } else {
    const int k = k0 + (t >> 5), row = row0 + (t & 31) * 4;
    r[0] = r[1] = r[2] = r[3] = 0.0f;
    if (k < kend) {
      const float* p = P + (int64_t)k * ld + row;
      if (vec && row + 3 < rows) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (row + i < rows) r[i] = p[i];
      }
    }
  }
}

template <bool TRANS>
__device__ __forceinline__ void store_tile(float (*S)[SG_BM + 4], const float (&r)[4]) {
  const int t = threadIdx.x;
  if (!TRANS) {
    const int row = t >> 1, k = (t & 1) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) S[k + i][row] = r[i];
  } else {
    const int k = t >> 5, row = (t & 31) * 4;
    *reinterpret_cast<float4*>(&S[k][row]) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(SG_THREADS)
gemm_simt_kernel(SimtArgs g) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int kbeg = blockIdx.z * g.k_per_split, kend = min(g.K, kbeg + g.k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[4], rb[4];
  const int ntiles = (kend - kbeg + SG_BK - 1) / SG_BK;
  if (ntiles > 0) {
    load_tile<AT>(g.A, g.lda, g.M, g.K, m0, kbeg, kend, g.vecA, ra);
    load_tile<BT>(g.B, g.ldb, g.N, g.K, n0, kbeg, kend, g.vecB, rb);
    store_tile<AT>(As[0], ra);
    store_tile<BT>(Bs[0], rb);
  }
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < ntiles) {
      load_tile<AT>(g.A, g.lda, g.M, g.K, m0, kbeg + (t + 1) * SG_BK, kend, g.vecA, ra);
      load_tile<BT>(g.B, g.ldb, g.N, g.K, n0, kbeg + (t + 1) * SG_BK, kend, g.vecB, rb);
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < ntiles) {
      store_tile<AT>(As[cur ^ 1], ra);
      store_tile<BT>(Bs[cur ^ 1], rb);
    }
    __syncthreads();
  }

  // epilogue
  const bool first_split = (blockIdx.z == 0);
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // this thread's 8-row column sums (fused bias gradient)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.accumulate) {
        // split-K partial sums: bias is added by split 0 only; act/mask are not allowed with accumulate (host checks)
        if (g.bias && first_split) v += g.bias[n];
        atomicAdd(&g.C[(int64_t)m * g.ldc + n], v);
      } else {
        if (g.bias) v += g.bias[n];
        if (g.act == 1) v = fmaxf(v, 0.0f);
        else if (g.act == 2) v = tanhf(v);
        if (g.mask_mode == 1) v = (g.mask_src[(int64_t)m * g.ldm + n] > 0.0f) ? v : 0.0f;
        else if (g.mask_mode == 2) { const float s = g.mask_src[(int64_t)m * g.ldm + n]; v *= (1.0f - s * s); }
        g.C[(int64_t)m * g.ldc + n] = v;
        csum[j] += v;
      }
    }
  }
  // column sums: registers (8 rows) -> shared memory tree over the 16 row groups -> ONE atomic per column per 128-row tile.  (An atomic
  // per element chains 16384 fp32 adds into one word at B = 16384: 1e-4 of max|db| on the critic's bias gradients, found by the fp64
  // three-way test.)
  if (g.colsum && !g.accumulate) {
    float* red = &As[0][0][0];                 // 16 x 128 floats; every thread passed the mainloop's last barrier
#pragma unroll
    for (int j = 0; j < 8; ++j) red[ty * SG_BN + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4))] = csum[j];
    __syncthreads();
    if (threadIdx.x < SG_BN) {
      float s = 0.0f;
#pragma unroll
      for (int y = 0; y < 16; y += 2) s += red[y * SG_BN + threadIdx.x] + red[(y + 1) * SG_BN + threadIdx.x];
      const int n = n0 + threadIdx.x;
      if (n < g.N) atomicAdd(&g.colsum[n], s);
    }
  }
}

int gemm_simt(const AseGemmParams& p, cudaStream_t st) {
  SimtArgs g;
  g.A = p.A; g.lda = p.lda; g.B = p.B; g.ldb = p.ldb; g.C = p.C; g.ldc = p.ldc;
  g.M = p.M; g.N = p.N; g.K = p.K; g.alpha = p.alpha; g.bias = p.bias; g.act = p.act;
  g.mask_src = p.mask_src; g.ldm = p.ldm; g.mask_mode = p.mask_src ? p.mask_mode : 0; g.accumulate = p.accumulate; g.colsum = p.colsum_out;
  int splits = p.split_k > 1 ? p.split_k : 1;
  if (!p.accumulate) splits = 1;
  int kps = (p.K + splits - 1) / splits;
  kps = (kps + SG_BK - 1) / SG_BK * SG_BK;
  if (kps < SG_BK) kps = SG_BK;
  splits = (p.K + kps - 1) / kps;
  if (splits < 1) splits = 1;
  g.k_per_split = kps;
  g.vecA = ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (p.lda % 4) == 0) ? 1 : 0;
  g.vecB = ((reinterpret_cast<uintptr_t>(p.B) & 15) == 0 && (p.ldb % 4) == 0) ? 1 : 0;
  dim3 grid(ceil_div(p.N, SG_BN), ceil_div(p.M, SG_BM), splits);
  if (!p.a_trans && !p.b_trans) gemm_simt_kernel<false, false><<<grid, SG_THREADS, 0, st>>>(g);
  else if (!p.a_trans && p.b_trans) gemm_simt_kernel<false, true><<<grid, SG_THREADS, 0, st>>>(g);
  else if (p.a_trans && !p.b_trans) gemm_simt_kernel<true, false><<<grid, SG_THREADS, 0, st>>>(g);
  else gemm_simt_kernel<true, true><<<grid, SG_THREADS, 0, st>>>(g);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

}  // namespace ase
