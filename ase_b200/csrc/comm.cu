// Gradient averaging behind the C ABI (SURVEY.md section 8b `grad_allreduce`, section 2b): one NCCL sum-allreduce of the flat gradient arena per
// minibatch (learning/amp_agent.py:348-363: Horovod averages inside optimizer.step; the 1/world factor is folded into the Adam kernel).
// libnccl.so.2 is resolved at run time with dlopen (the library PyTorch already loaded in the process), so libase_b200.so has no link-time
// dependency on it and single-GPU users never touch it.  The communicator is the library's own (ncclCommInitRankConfig from a unique id the
// host exchanges over whatever channel it has -- torch.distributed's store in the Python mirror).
#include <dlfcn.h>
#include <string.h>
#include <new>
#include <nccl.h>
#include "common.cuh"

namespace ase {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(const char* path) {
  if (g_nccl.handle) return ASE_OK;
  const char* cands[3] = {path, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; ++i) if (cands[i] && cands[i][0]) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("ase_comm: cannot dlopen libnccl.so.2 (%s)", dlerror()); return ASE_ERR_UNSUPPORTED; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRankConfig = (decltype(g_nccl.CommInitRankConfig))dlsym(h, "ncclCommInitRankConfig");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
    set_error("ase_comm: libnccl.so.2 lacks a required symbol");
    return ASE_ERR_UNSUPPORTED;
  }
  g_nccl.handle = h;
  return ASE_OK;
}
#define ASE_NCCL_OK(expr)                                                                                           \
  do {                                                                                                              \
    ncclResult_t _r = (expr);                                                                                       \
    if (_r != ncclSuccess) {                                                                                        \
      ase::set_error("%s failed: %s", #expr, ase::g_nccl.GetErrorString ? ase::g_nccl.GetErrorString(_r) : "?");    \
      return ASE_ERR_CUDA;                                                                                          \
    }                                                                                                               \
  } while (0)
}  // namespace ase

struct AseComm { ncclComm_t comm; int rank, world; };

using namespace ase;

extern "C" int ase_comm_load(const char* libnccl_path) { return nccl_load(libnccl_path); }

extern "C" int ase_comm_unique_id(uint8_t* out128) {
  ASE_CHECK_ARG(out128 != nullptr, "ase_comm_unique_id: null pointer");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  ASE_NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return ASE_OK;
}

extern "C" int ase_comm_create(const uint8_t* id128, int rank, int world, AseComm** out) {
  ASE_CHECK_ARG(id128 && out && world >= 1 && rank >= 0 && rank < world, "ase_comm_create: bad argument");
  int rc = nccl_load(nullptr);
  if (rc) return rc;
  AseComm* c = new (std::nothrow) AseComm;
  ASE_CHECK_ARG(c != nullptr, "ase_comm_create: out of host memory");
  c->rank = rank; c->world = world; c->comm = nullptr;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); delete c; return ASE_ERR_CUDA; }
  *out = c;
  return ASE_OK;
}

extern "C" void ase_comm_destroy(AseComm* c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

// in-place fp32 sum over the ranks of the communicator, enqueued on `stream` (the learner's stream: it sits between the last dW GEMM and
// adam_kernel; see DESIGN.md section 7 for why it is NOT overlapped with the persistent GEMM kernels)
extern "C" int ase_grad_allreduce(AseComm* c, float* buf, int64_t count, void* stream) {
  ASE_CHECK_ARG(c && c->comm && buf && count >= 0, "ase_grad_allreduce: bad argument");
  if (count == 0 || c->world == 1) return ASE_OK;
  ASE_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, c->comm, (cudaStream_t)stream));
  return ASE_OK;
}
// fp64 variant: RunningMeanStd statistics are averaged once per epoch (hvd.sync_stats)
extern "C" int ase_comm_allreduce_f64(AseComm* c, double* buf, int64_t count, void* stream) {
  ASE_CHECK_ARG(c && c->comm && buf && count >= 0, "ase_comm_allreduce_f64: bad argument");
  if (count == 0 || c->world == 1) return ASE_OK;
  ASE_NCCL_OK(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ncclSum, c->comm, (cudaStream_t)stream));
  return ASE_OK;
}
