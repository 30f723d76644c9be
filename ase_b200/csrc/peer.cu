// Gradient allreduce + Adam as ONE kernel over NVLink peer memory (one process per GPU, up to 8 GPUs of one NVSwitch node).
//
// What it replaces: ncclAllReduce of the 28 MB gradient arena followed by adam_kernel (learning/amp_agent.py:348-363: Horovod averages the
// gradients inside optimizer.step).  Every rank keeps its gradient arena in a buffer whose CUDA IPC handle the other ranks have opened, so a
// kernel can load and store any peer's arena directly over NVLink / NVSwitch:
//
//   1. barrier "ready":  rank r writes its call counter into slot r of every peer's flag array (system-scope release) and waits until all
//                        slots of its own array have reached the counter (acquire)  -- every rank's backward pass is complete
//   2. reduce-scatter + all-gather, in place: rank r owns the slice [lo_r, hi_r) of the arena; for each element of its slice it loads the
//                        value from every rank (fixed order 0..N-1: the sum is computed once, by one rank, so all ranks end up with
//                        bit-identical gradients) and stores the sum back into every rank's arena.  No other rank touches that element.
//                        Traffic per rank: (N-1)/N of the arena in, (N-1)/N out -- what a ring allreduce moves, without its 2(N-1) steps.
//   3. barrier "done":   the last block of a rank to finish step 2 signals all peers; every block waits for all peers' signals
//   4. Adam over the FULL local arena (optimizer state stays replicated: checkpoints and the single-GPU path are unchanged), same
//                        arithmetic, in the same order, as adam_kernel; the rank's own slice is updated inside step 2 from the sums in
//                        registers, under the latency of its peer stores.
//
// All blocks spin on flags, so the grid must be resident at once: one block per SM, launched after every earlier kernel of the stream
// has finished.  A spin that exceeds ~15 s of clock64() gives up and raises the error word (bit 2 of the learner's status flag when there is
// one): a lost peer produces an error at the next status check, not a hung GPU.
// Peer loads bypass L1 (ld.global.cg): the same addresses are read again in step 4 and by the next call.
#include <new>
#include <string.h>
#include "common.cuh"
#include "kernels.h"

namespace ase {

constexpr int PEER_MAX = 8;
constexpr int PEER_FLAG_BYTES = 4096;      // [0, 64): ready[8] u64, [64, 128): done[8] u64, 128: block counter u32, 132: error u32
constexpr int PEER_THREADS = 512;

struct PeerTab {
  float* g[PEER_MAX];                 // every rank's gradient arena (g[rank] is local)
  unsigned long long* flags[PEER_MAX];
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// wait until slots [base, base + world) of this rank's flag array have reached `epoch`; false on timeout
__device__ __forceinline__ bool peer_wait(const unsigned long long* mine, int base, int world, unsigned long long epoch, long long timeout) {
  bool ok = true;
  if ((int)threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + base + threadIdx.x) < epoch) {
      if (clock64() - t0 > timeout) { ok = false; break; }
      __nanosleep(64);
    }
  }
  return __syncthreads_and((int)ok) != 0;
}

struct AdamK { float grad_scale, b1, b2, omb1, omb2, step_size, bc2_sqrt, eps; };
// torch.optim.Adam single-tensor arithmetic on 4 consecutive elements, in the order of adam_kernel (loss_kernels.cu)
__device__ __forceinline__ void adam_update4(float4* __restrict__ p4, float4* __restrict__ m4, float4* __restrict__ v4, int64_t i, const float4 g, const AdamK& k) {
  float4 mm = m4[i], vv = v4[i], pp = p4[i];
  const float gi[4] = {g.x * k.grad_scale, g.y * k.grad_scale, g.z * k.grad_scale, g.w * k.grad_scale};
  float* mp = &mm.x; float* vp = &vv.x; float* ppp = &pp.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float mi = k.b1 * mp[j] + k.omb1 * gi[j];
    const float vi = k.b2 * vp[j] + k.omb2 * gi[j] * gi[j];
    mp[j] = mi; vp[j] = vi;
    const float denom = sqrtf(vi) / k.bc2_sqrt + k.eps;
    ppp[j] = ppp[j] - k.step_size * (mi / denom);
  }
  m4[i] = mm; v4[i] = vv; p4[i] = pp;
}

__global__ void __launch_bounds__(PEER_THREADS, 1)
peer_allreduce_adam_kernel(PeerTab t, unsigned long long epoch, int64_t n4 /* arena / 4 */, int64_t lo4, int64_t hi4,
                           float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                           const AdamK k, long long timeout, unsigned* status) {
  float4* p4 = reinterpret_cast<float4*>(p); float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
  unsigned long long* mine = t.flags[t.rank];
  unsigned* counter = reinterpret_cast<unsigned*>(mine + 16);
  unsigned* error = counter + 1;
  long long* dbg = reinterpret_cast<long long*>(mine + 32);      // [256, 320): phase timestamps of block 0 (clock64), for tools/peer_adam_check.py
  const bool tick = blockIdx.x == 0 && threadIdx.x == 0;
  if (tick) dbg[0] = clock64();
  // ---- 1. ready
  if (blockIdx.x == 0 && (int)threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + t.rank, epoch);
  }
  bool ok = peer_wait(mine, 0, t.world, epoch, timeout);
  if (tick) dbg[1] = clock64();
  // ---- 2. reduce my slice from every rank, write the sum to every rank (in place)
  if (ok) {
    // 4 independent elements per thread and trip: the loop is bound by NVLink round trips (~3 us), not by issue
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi4; i0 += 4 * stride) {
      float4 s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * stride;
        s[u] = (i < hi4) ? __ldcg(reinterpret_cast<const float4*>(t.g[0]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int q = 1; q < t.world; ++q) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = i0 + u * stride;
          x[u] = (i < hi4) ? __ldcg(reinterpret_cast<const float4*>(t.g[q]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s[u].x += x[u].x; s[u].y += x[u].y; s[u].z += x[u].z; s[u].w += x[u].w; }
      }
      for (int q = 0; q < t.world; ++q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = i0 + u * stride;
          if (i < hi4) __stcg(reinterpret_cast<float4*>(t.g[q]) + i, s[u]);
        }
      }
      // the optimizer step of this rank's own slice, straight from the sums in registers: it runs while the peer stores above are in flight
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < hi4) adam_update4(p4, m4, v4, i, s[u], k);
      }
    }
  }
  // ---- 3. done: the rank's last block tells everybody
  if (tick) dbg[2] = clock64();
  __threadfence_system();
  __syncthreads();
  if (tick) dbg[3] = clock64();
  __shared__ unsigned last;
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    last = (prev == gridDim.x - 1) ? 1u : 0u;
    if (last) *counter = 0u;                       // every block of this launch has passed: ready for the next call
  }
  __syncthreads();
  if (last && (int)threadIdx.x < t.world) {
    __threadfence_system();
    st_release_sys(t.flags[threadIdx.x] + 8 + t.rank, epoch);
  }
  ok = peer_wait(mine, 8, t.world, epoch, timeout) && ok;
  if (tick) dbg[4] = clock64();
  if (!ok) {
    if (threadIdx.x == 0) { atomicOr(error, 1u); if (status) atomicOr(status, 4u); }
    return;                                          // no update from gradients that are not known to be complete
  }
  // ---- 4. Adam over the rest of the local arena: the slices the peers reduced and wrote here
  const float4* g4 = reinterpret_cast<const float4*>(t.g[t.rank]);
  const int64_t own = hi4 - lo4;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n4 - own; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = j < lo4 ? j : j + own;
    adam_update4(p4, m4, v4, i, __ldcg(g4 + i), k);
  }
  if (tick) dbg[5] = clock64();
}

}  // namespace ase

using namespace ase;

struct AsePeer {
  int world, rank;
  void* base[PEER_MAX];          // mapped allocations (base[rank] = the local cudaMalloc)
  int64_t arena;                 // floats, a multiple of 4
  unsigned long long epoch;
  int sms;
  long long timeout_clk;         // ~15 s of SM clocks (queried once: cudaDevAttrClockRate is a slow attribute)
};

extern "C" int64_t ase_peer_buffer_bytes(int64_t arena_floats) { return PEER_FLAG_BYTES + ((arena_floats + 3) / 4 * 4) * 4; }

// cudaMalloc (NOT the caller's caching allocator: IPC handles need a whole allocation), zeroed; returns the local pointer and its IPC handle
extern "C" int ase_peer_alloc(int64_t arena_floats, void** ptr, uint8_t* handle64) {
  ASE_CHECK_ARG(ptr && handle64 && arena_floats > 0, "ase_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void* p = nullptr;
  const int64_t bytes = ase_peer_buffer_bytes(arena_floats);
  ASE_CUDA_OK(cudaMalloc(&p, bytes));
  ASE_CUDA_OK(cudaMemset(p, 0, bytes));
  ASE_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return ASE_ERR_UNSUPPORTED; }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return ASE_OK;
}

// handles: world x 64 bytes in rank order (the host gathers them over whatever channel it has)
extern "C" int ase_peer_open(const uint8_t* handles, int world, int rank, void* local, int64_t arena_floats, AsePeer** out) {
  ASE_CHECK_ARG(handles && local && out && world >= 2 && world <= PEER_MAX && rank >= 0 && rank < world, "ase_peer_open: bad argument (2 <= world <= 8)");
  AsePeer* pr = new (std::nothrow) AsePeer;
  ASE_CHECK_ARG(pr != nullptr, "ase_peer_open: out of host memory");
  memset(pr, 0, sizeof(*pr));
  pr->world = world; pr->rank = rank; pr->arena = (arena_floats + 3) / 4 * 4; pr->epoch = 0;
  int dev = 0; cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&pr->sms, cudaDevAttrMultiProcessorCount, dev);
  if (pr->sms <= 0) pr->sms = 148;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  pr->timeout_clk = 15LL * 1000LL * (long long)(clk_khz > 0 ? clk_khz : 1900000);
  for (int q = 0; q < world; ++q) {
    if (q == rank) { pr->base[q] = local; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, handles + 64 * q, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&pr->base[q], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", q, cudaGetErrorString(e)); cudaGetLastError();
      for (int k = 0; k < q; ++k) if (k != rank && pr->base[k]) cudaIpcCloseMemHandle(pr->base[k]);
      delete pr;
      return ASE_ERR_UNSUPPORTED;
    }
  }
  *out = pr;
  return ASE_OK;
}

extern "C" void ase_peer_close(AsePeer* pr, int free_local) {
  if (!pr) return;
  for (int q = 0; q < pr->world; ++q) if (q != pr->rank && pr->base[q]) cudaIpcCloseMemHandle(pr->base[q]);
  if (free_local && pr->base[pr->rank]) cudaFree(pr->base[pr->rank]);
  delete pr;
}

extern "C" int ase_peer_debug(AsePeer* pr, long long* out8) {
  ASE_CHECK_ARG(pr && out8, "ase_peer_debug: null argument");
  ASE_CUDA_OK(cudaMemcpy(out8, reinterpret_cast<uint8_t*>(pr->base[pr->rank]) + 256, 64, cudaMemcpyDeviceToHost));
  return ASE_OK;
}

extern "C" float* ase_peer_grads(void* local) { return reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(local) + PEER_FLAG_BYTES); }

extern "C" int ase_peer_status(AsePeer* pr, int* err, void* stream) {
  ASE_CHECK_ARG(pr && err, "ase_peer_status: null argument");
  unsigned e = 0;
  ASE_CUDA_OK(cudaMemcpyAsync(&e, reinterpret_cast<uint8_t*>(pr->base[pr->rank]) + 132, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  ASE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  *err = (int)e;
  return ASE_OK;
}

namespace ase {
int launch_peer_adam(AsePeer* pr, float* p, float* m, float* v, float grad_scale, float b1, float b2, float lr, float eps, int64_t step,
                     unsigned* status, cudaStream_t st) {
  PeerTab t; memset(&t, 0, sizeof(t));
  t.world = pr->world; t.rank = pr->rank;
  for (int q = 0; q < pr->world; ++q) {
    t.flags[q] = reinterpret_cast<unsigned long long*>(pr->base[q]);
    t.g[q] = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pr->base[q]) + PEER_FLAG_BYTES);
  }
  const int64_t n4 = pr->arena / 4;
  const int64_t per = (n4 + pr->world - 1) / pr->world;
  const int64_t lo4 = per * pr->rank < n4 ? per * pr->rank : n4;
  const int64_t hi4 = lo4 + per < n4 ? lo4 + per : n4;
  AdamConsts c = adam_consts(b1, b2, lr, eps, step);
  pr->epoch += 1;
  const long long timeout = pr->timeout_clk;
  AdamK k{grad_scale, c.b1, c.b2, c.omb1, c.omb2, c.step_size, c.bc2_sqrt, c.eps};
  peer_allreduce_adam_kernel<<<pr->sms, PEER_THREADS, 0, st>>>(t, pr->epoch, n4, lo4, hi4, p, m, v, k, timeout, status);
  ASE_LAUNCH_OK();
  return ASE_OK;
}
}  // namespace ase
