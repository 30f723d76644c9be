// Internal (non-ABI) declarations shared by the .cu files of libase_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ase_b200.h"

namespace ase {

// ------------------------------------------------------------------ RunningMeanStd
struct RmsBatchList { const float* x[3]; int64_t ld[3]; int rows; };
struct RmsDst { float* y[3]; int64_t ld[3]; float* hi[3]; float* lo[3]; int64_t ldp[3]; };   // optional TF32 planes per destination
int64_t rms_scratch_bytes(int cols, int rows, int nbatch);
int rms_update_batches(const RmsBatchList& bl, int nbatch, int cols, double* mean, double* var, double* count, float eps,
                       int update, void* scratch, float** meanf_out, float** stdf_out, cudaStream_t st);
int rms_normalize(const float* x, int64_t ldx, int rows, int cols, const float* meanf, const float* stdf, int unnorm,
                  const RmsDst& dst, cudaStream_t st);
int rms_apply(const float* x, int64_t ldx, int rows, int cols, const double* mean, const double* var, float eps, int unnorm,
              float* y, int64_t ldy, cudaStream_t st);
int copy_cols(const float* x, int64_t ldx, int rows, int cols, float* y, int64_t ldy, cudaStream_t st, float* hi = nullptr, float* lo = nullptr,
              int64_t ldp = 0);

// ------------------------------------------------------------------ GEMM backends
int gemm_simt(const AseGemmParams& p, cudaStream_t st);
// TF32 hi/lo operand planes kept next to registered fp32 buffers (gemm_tc.cu)
struct PlaneBuf {
  const float* base; int64_t capacity;          // fp32 buffer (floats)
  float* hi; float* lo; int64_t plane_capacity; // planes (floats each; the FP16 format stores halfs in the same space)
  int64_t ld; int rows, cols; int64_t ldp;      // geometry declared by the last full writer / first reader
  bool valid;
  int amax_slot;                                // FP16 format: transient slot holding max |x| of the current contents, -1 = unknown
};
struct PlaneRegistry {
  static constexpr int MAX = 160;
  static constexpr int SLOTS = 1024;
  PlaneBuf b[MAX]; int n = 0;
  // FP16 format (backend 2): per-tensor power-of-two scales live in device memory (no host sync anywhere)
  bool f16 = false;
  unsigned* amax = nullptr;     // [SLOTS] transient max |x| slots (uint bits), zeroed by begin_call
  float* tscale = nullptr;      // [SLOTS][2] scale / inverse of unregistered operands split into the shared workspace
  float* bscale = nullptr;      // [MAX][2] scale / inverse of each registered buffer's current planes
  int n_slots = 0, next_slot = 0;
  static int64_t device_bytes() { return (int64_t)SLOTS * 4 + (int64_t)SLOTS * 8 + (int64_t)MAX * 8; }
  void attach_device(void* mem) { amax = (unsigned*)mem; tscale = (float*)((char*)mem + SLOTS * 4); bscale = tscale + 2 * SLOTS; n_slots = SLOTS; }
  int begin_call(cudaStream_t st);
  int new_slot();
  PlaneBuf* find(const float* p);
  // a non-GEMM kernel is about to write the whole buffer [rows, cols] (ld) INCLUDING its planes: returns the entry (valid) or null
  PlaneBuf* declare(const float* base, int64_t ld, int rows, int cols);
  void add(const float* base, int64_t capacity, float* hi, float* lo, int64_t plane_capacity);
  void invalidate(const float* p);
  void invalidate_range(const float* lo_, const float* hi_);
};
int gemm_tc(const AseGemmParams& p, cudaStream_t st, PlaneRegistry* reg = nullptr);   // tcgen05 3xTF32 (backend 1) / 3xFP16-scaled (backend 2) (gemm_tc.cu)
bool gemm_tc_supported(const AseGemmParams& p);
int64_t gemm_tc_workspace_bytes(int M, int N, int K);
int gemm_tc_tile_n(int N);      // N extent of the output tile the tcgen05 backend will use for this N
int gemm_dispatch(const AseGemmParams& p, cudaStream_t st, PlaneRegistry* reg = nullptr);    // picks the backend named in p.backend (falls back to SIMT for shapes tc rejects)

// ------------------------------------------------------------------ loss-side accumulators (doubles)
enum {
  ACC_MSUM = 0, ACC_ALOSS, ACC_CLOSS, ACC_BLOSS, ACC_CLIPPED, ACC_KL, ACC_DIV,
  ACC_BCE_AGENT, ACC_BCE_DEMO, ACC_ACC_AGENT, ACC_ACC_DEMO, ACC_LOGIT_AGENT, ACC_LOGIT_DEMO,
  ACC_GP, ACC_WLOGIT2, ACC_WDISC2, ACC_ENC, ACC_COUNT = 24
};

struct PpoHeadArgs {
  const float* mu; int64_t ld_mu;      // [B or 2B, A]; rows B.. hold the diversity pass
  const float* values;                 // [B]
  const float* actions; const float* old_logp; const float* adv; const float* old_mu; const float* old_sigma;
  const float* returns; const float* mask; const float* logstd;
  const float* z; const float* z2; int Z;
  int B, A;
  int has_div;
  int mu_tanh;                         // mu = tanh(raw): gradients are taken to the pre-activation
  float e_clip, critic_coef, bounds_coef, div_bonus, div_tar;
  float* dmu;                          // same layout as mu
  float* dv;                           // [B]
  double* acc;
};

struct FinalizeArgs {
  const double* acc; float* out; const float* logstd;
  int kind, B, Ba, A;
  float critic_coef, entropy_coef, bounds_coef, disc_coef, logit_reg, gp_coef, weight_decay, enc_coef, div_bonus;
};

int launch_mask_sum(const float* mask, int rows, double* acc, cudaStream_t st);
int launch_ppo_head(const PpoHeadArgs& a, cudaStream_t st);
int launch_disc_head(const float* logit, int Ba, float disc_coef, float* dlogit, double* acc, float* out_agent, float* out_demo, cudaStream_t st);
int launch_enc_head(const float* e, int rows, int Z, const float* z, float enc_coef, float* de, float* enc_pred, double* acc, cudaStream_t st);
int launch_gp_u_last(const float* h, int64_t ldh, int rows, int cols, const float* w, float* u, cudaStream_t st);
int launch_gp_scale(float* g, int64_t total, float scale, double* acc, cudaStream_t st);
int launch_colsum(const float* dz, int64_t ld, int rows, int cols, float* db, cudaStream_t st);
int launch_weight_reg(const float* w, float* g, int64_t n, float coef, double* acc, int idx, int idx2, cudaStream_t st);
int launch_finalize(const FinalizeArgs& f, cudaStream_t st);
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float grad_scale, float b1, float b2, float lr, float eps,
                int64_t step, cudaStream_t st);

}  // namespace ase
