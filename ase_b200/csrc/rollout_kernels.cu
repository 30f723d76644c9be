// Rollout-side scans / reductions (HBM-bound): GAE, AMP/ASE rewards, advantage normalisation.
#include "common.cuh"
#include "kernels.h"

namespace ase {

// learning/common_agent.py:437-449.  One thread per env, reverse scan over the horizon; [H,N] row-major so
// a warp reads 32 consecutive envs per step (coalesced).  Algorithmic bytes: (3*4 + 1 + 2*4) * H * N.
__global__ void __launch_bounds__(128)
gae_kernel(const uint8_t* __restrict__ dones, const float* __restrict__ values, const float* __restrict__ rewards,
           const float* __restrict__ next_values, int H, int N, float gamma, float tau,
           float* __restrict__ advs, float* __restrict__ returns) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  float last = 0.0f;
  for (int t = H - 1; t >= 0; --t) {
    const int64_t i = (int64_t)t * N + e;
    const float nd = 1.0f - (float)dones[i];
    const float v = values[i];
    const float delta = rewards[i] + gamma * next_values[i] - v;
    last = delta + gamma * tau * nd * last;
    advs[i] = last;
    if (returns) returns[i] = last + v;
  }
}

// amp_agent.py:570-577, ase_agent.py:404-411,469-472,484-490.  One warp per row (latent dot product).
__global__ void __launch_bounds__(256)
amp_rewards_kernel(const float* __restrict__ logits, const float* __restrict__ enc_pred, const float* __restrict__ z,
                   int zdim, int rows, float disc_scale, float enc_scale, const float* __restrict__ task_r,
                   float task_w, float disc_w, float enc_w, float* __restrict__ disc_r, float* __restrict__ enc_r,
                   float* __restrict__ combined) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float er = 0.0f;
  if (enc_pred) {
    float d = 0.0f;
    for (int j = lane; j < zdim; j += 32) d += enc_pred[(int64_t)row * zdim + j] * z[(int64_t)row * zdim + j];
    d = warp_sum(d);
    er = fmaxf(d, 0.0f) * enc_scale;      // clamp_min(-err, 0), err = -sum(enc*z)
  }
  if (lane == 0) {
    const float l = logits[row];
    const float prob = 1.0f / (1.0f + expf(-l));
    const float dr = -logf(fmaxf(1.0f - prob, 0.0001f)) * disc_scale;
    if (disc_r) disc_r[row] = dr;
    if (enc_r && enc_pred) enc_r[row] = er;
    if (combined) combined[row] = task_w * (task_r ? task_r[row] : 0.0f) + disc_w * dr + (enc_pred ? enc_w * er : 0.0f);
  }
}

// Gaussian sampling head (eval mode) + eps-greedy override; one warp per row.
__global__ void __launch_bounds__(256)
policy_sample_kernel(const float* __restrict__ mu, const float* __restrict__ logstd, const float* __restrict__ noise,
                     const float* __restrict__ rand_mask, int rows, int A, float* __restrict__ actions,
                     float* __restrict__ neglogp, float* __restrict__ sigma_out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bool det = rand_mask && rand_mask[row] == 0.0f;
  float s = 0.0f, sumlog = 0.0f;
  for (int j = lane; j < A; j += 32) {
    const float ls = logstd[j], sg = expf(ls), m = mu[(int64_t)row * A + j];
    const float a = m + sg * noise[(int64_t)row * A + j];
    const float t = (a - m) / sg;
    s += t * t; sumlog += ls;
    actions[(int64_t)row * A + j] = det ? m : a;
    if (sigma_out) sigma_out[(int64_t)row * A + j] = sg;
  }
  s = warp_sum(s); sumlog = warp_sum(sumlog);
  if (lane == 0 && neglogp) neglogp[row] = 0.5f * s + (float)(0.5 * 1.8378770664093453 * (double)A) + sumlog;
}

// stats[0..2] = sum(m), sum(v*m), sum((v*m)^2) ; unmasked: m = 1
__global__ void __launch_bounds__(256)
adv_stats_kernel(const float* __restrict__ ret, const float* __restrict__ val, const float* __restrict__ mask, int rows,
                 double* __restrict__ stats) {
  __shared__ double sm[32 * 3];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) {
    const float a = ret[i] - val[i];
    const float m = mask ? mask[i] : 1.0f;
    const double vm = (double)(a * m);
    acc[0] += (double)m; acc[1] += vm; acc[2] += vm * vm;
  }
  block_sum<3>(acc, sm);
  if (threadIdx.x == 0) { atomicAdd(&stats[0], acc[0]); atomicAdd(&stats[1], acc[1]); atomicAdd(&stats[2], acc[2]); }
}

__global__ void __launch_bounds__(256)
adv_apply_kernel(const float* __restrict__ ret, const float* __restrict__ val, int rows, const double* __restrict__ stats,
                 float* __restrict__ advs) {
  // torch_ext.normalization_with_masks: var = (E[(vm)^2] - E[vm]^2) * n/(n-1)   (for mask == 1 this is the
  // unbiased variance that advantages.std() uses in common_agent.py:543)
  const double n = stats[0];
  const double mean = stats[1] / n;
  const double min_sqr = stats[2] / n - mean * mean;
  const double var = min_sqr * n / (n - 1.0);
  const float meanf = (float)mean, stdf = (float)sqrt(var > 0.0 ? var : 0.0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x)
    advs[i] = ((ret[i] - val[i]) - meanf) / (stdf + 1e-8f);
}

}  // namespace ase

using namespace ase;

extern "C" int ase_gae(const uint8_t* dones, const float* values, const float* rewards, const float* next_values,
                       int horizon, int num_envs, float gamma, float tau, float* advs, float* returns, void* stream) {
  ASE_CHECK_ARG(dones && values && rewards && next_values && advs, "ase_gae: null pointer");
  if (horizon <= 0 || num_envs <= 0) return ASE_OK;
  gae_kernel<<<ceil_div(num_envs, 128), 128, 0, (cudaStream_t)stream>>>(dones, values, rewards, next_values, horizon, num_envs,
                                                                         gamma, tau, advs, returns);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

extern "C" int ase_amp_rewards(const float* disc_logits, const float* enc_pred, const float* latents, int latent_dim, int rows,
                               float disc_scale, float enc_scale, const float* task_rewards, float task_w, float disc_w,
                               float enc_w, float* disc_r, float* enc_r, float* combined, void* stream) {
  ASE_CHECK_ARG(disc_logits, "ase_amp_rewards: null logits");
  ASE_CHECK_ARG((enc_pred == nullptr) == (latents == nullptr), "ase_amp_rewards: enc_pred and latents go together");
  if (rows <= 0) return ASE_OK;
  amp_rewards_kernel<<<ceil_div((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      disc_logits, enc_pred, latents, latent_dim, rows, disc_scale, enc_scale, task_rewards, task_w, disc_w, enc_w, disc_r, enc_r,
      combined);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

extern "C" int ase_adv_normalize(const float* returns, const float* values, const float* mask, int rows, float* advs,
                                 void* scratch, void* stream) {
  ASE_CHECK_ARG(returns && values && advs && scratch, "ase_adv_normalize: null pointer");
  if (rows <= 0) return ASE_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ASE_CUDA_OK(cudaMemsetAsync(scratch, 0, 3 * sizeof(double), st));
  const int blocks = min(ceil_div(rows, 256), 148 * 4);
  adv_stats_kernel<<<blocks, 256, 0, st>>>(returns, values, mask, rows, (double*)scratch);
  ASE_LAUNCH_OK();
  adv_apply_kernel<<<blocks, 256, 0, st>>>(returns, values, rows, (const double*)scratch, advs);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

extern "C" int ase_policy_sample(const float* mu, const float* logstd, const float* noise, const float* rand_mask, int rows,
                                 int act_dim, float* actions, float* neglogp, float* sigma_out, void* stream) {
  ASE_CHECK_ARG(mu && logstd && noise && actions, "ase_policy_sample: null pointer");
  if (rows <= 0) return ASE_OK;
  policy_sample_kernel<<<ceil_div((int64_t)rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(mu, logstd, noise, rand_mask, rows, act_dim,
                                                                                           actions, neglogp, sigma_out);
  ASE_LAUNCH_OK();
  return ASE_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Minibatch gather (learning/amp_datasets.py:14-27 AMPDataset._get_item + the demo / replay row fetches of
// amp_agent.py:194-202): dst_i[r, :] = src_i[idx_i[r], :] for up to ASE_GATHER_MAX tensors in ONE launch
// (the reference issues one advanced-indexing kernel per tensor, 13 per minibatch).  blockIdx.y = tensor.
// ------------------------------------------------------------------------------------------------------------
namespace ase {
__global__ void __launch_bounds__(256)
gather_rows_kernel(AseGatherBatch b) {
  const AseGatherItem it = b.item[blockIdx.y];
  const bool vec = ((it.cols & 3) == 0) && ((it.src_ld & 3) == 0) && ((it.dst_ld & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(it.src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(it.dst) & 15) == 0);
  if (vec) {
    const int c4n = it.cols >> 2;
    const int64_t total = (int64_t)it.rows * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / c4n), c = (int)(i - (int64_t)r * c4n) * 4;
      const int64_t sr = it.idx ? it.idx[r] : r;
      *reinterpret_cast<float4*>(it.dst + (int64_t)r * it.dst_ld + c) = *reinterpret_cast<const float4*>(it.src + sr * it.src_ld + c);
    }
  } else {
    const int64_t total = (int64_t)it.rows * it.cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int r = (int)(i / it.cols), c = (int)(i - (int64_t)r * it.cols);
      const int64_t sr = it.idx ? it.idx[r] : r;
      it.dst[(int64_t)r * it.dst_ld + c] = it.src[sr * it.src_ld + c];
    }
  }
}
}  // namespace ase

extern "C" int ase_gather_rows(const AseGatherBatch* batch, void* stream) {
  ASE_CHECK_ARG(batch && batch->count >= 0 && batch->count <= ASE_GATHER_MAX, "ase_gather_rows: bad batch");
  if (batch->count == 0) return ASE_OK;
  int64_t big = 0;
  for (int i = 0; i < batch->count; ++i) {
    const AseGatherItem& it = batch->item[i];
    ASE_CHECK_ARG(it.src && it.dst && it.rows >= 0 && it.cols > 0 && it.src_ld >= it.cols && it.dst_ld >= it.cols, "ase_gather_rows: item %d", i);
    big = ase::imax64(big, (int64_t)it.rows * it.cols);
  }
  if (big == 0) return ASE_OK;
  dim3 grid((unsigned)ase::imin64((big / 4 + 255) / 256 + 1, 148 * 4), (unsigned)batch->count);
  ase::gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*batch);
  ASE_LAUNCH_OK();
  return ASE_OK;
}
