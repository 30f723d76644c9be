// Loss heads of calc_gradients with hand-derived gradients, the gradient-penalty element-wise pieces,
// bias-gradient column sums, weight regularisers and the fused Adam step.  All scalar statistics are
// accumulated as doubles in `acc` (one atomicAdd per block) and turned into the reference's train_result by
// finalize_scalars_kernel -- no host synchronisation anywhere (the reference .item()s the KL every minibatch).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

namespace ase {

__global__ void __launch_bounds__(256)
mask_sum_kernel(const float* __restrict__ mask, int rows, double* __restrict__ acc) {
  __shared__ double sm[32];
  double v[1] = {0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += gridDim.x * blockDim.x) v[0] += mask ? (double)mask[i] : 1.0;
  block_sum<1>(v, sm);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_MSUM], v[0]);
}

// ---------------------------------------------------------------------------------------------------------
// PPO head: Gaussian neglogp (rl_games ModelA2CContinuousLogStd.neglogp), clipped surrogate
// (common_agent.py:505-519), value MSE (:521-534, clip_value False), bound loss (:456-464), policy KL
// (torch_ext.policy_kl), latent-diversity loss (ase_agent.py:445-467); masked means ase_agent.py:236-241.
// One warp per row; lanes stride the action dimension.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ppo_head_kernel(PpoHeadArgs a) {
  __shared__ double sm[32 * 6];
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[6] = {0, 0, 0, 0, 0, 0};   // a_loss, c_loss, b_loss, clipped, kl, div
  if (row < a.B) {
    const int A = a.A;
    const float m = a.mask ? a.mask[row] : 1.0f;
    const float inv_msum = (float)(1.0 / a.acc[ACC_MSUM]);
    const float* mu = a.mu + (int64_t)row * a.ld_mu;
    const float* act = a.actions + (int64_t)row * A;
    const float* omu = a.old_mu + (int64_t)row * A;
    const float* osg = a.old_sigma + (int64_t)row * A;
    float s = 0.0f, sumlog = 0.0f, bl = 0.0f, kl = 0.0f;
    for (int j = lane; j < A; j += 32) {
      const float ls = a.logstd[j], sg = expf(ls);
      const float t = (act[j] - mu[j]) / sg;
      s += t * t; sumlog += ls;
      const float lo = fminf(mu[j] + 1.0f, 0.0f), hi = fmaxf(mu[j] - 1.0f, 0.0f);
      bl += lo * lo + hi * hi;
      const float c1 = logf(osg[j] / sg + 1e-5f);
      const float dm = omu[j] - mu[j];
      const float c2 = (sg * sg + dm * dm) / (2.0f * (osg[j] * osg[j] + 1e-5f));
      kl += c1 + c2 - 0.5f;
    }
    s = warp_sum(s); sumlog = warp_sum(sumlog); bl = warp_sum(bl); kl = warp_sum(kl);
    const float nlp = 0.5f * s + (float)(0.5 * 1.8378770664093453 * (double)A) + sumlog;   // log(2*pi)
    const float ratio = expf(a.old_logp[row] - nlp);
    const float adv = a.adv[row];
    const float lo_c = 1.0f - a.e_clip, hi_c = 1.0f + a.e_clip;
    const float t1 = -adv * ratio, t2 = -adv * fminf(fmaxf(ratio, lo_c), hi_c);
    const float a_loss = fmaxf(t1, t2);
    const float clipped = (fabsf(ratio - 1.0f) > a.e_clip) ? 1.0f : 0.0f;
    const float g1 = adv * ratio;                                         // d t1 / d nlp
    const float g2 = (ratio >= lo_c && ratio <= hi_c) ? g1 : 0.0f;        // d t2 / d nlp (clamp passes grad inside, inclusive)
    float g_nlp = (t1 > t2) ? g1 : ((t1 < t2) ? g2 : 0.5f * (g1 + g2));   // torch.max splits ties evenly
    const float wrow = m * inv_msum;
    g_nlp *= wrow;
    // latent diversity
    float coef_div = 0.0f, div_l = 0.0f;
    const float* mu2 = nullptr;
    if (a.has_div) {
      mu2 = a.mu + (int64_t)(a.B + row) * a.ld_mu;
      float sd = 0.0f, zz = 0.0f;
      for (int j = lane; j < A; j += 32) {
        const float d = fminf(fmaxf(mu[j], -1.0f), 1.0f) - fminf(fmaxf(mu2[j], -1.0f), 1.0f);
        sd += d * d;
      }
      for (int j = lane; j < a.Z; j += 32) zz += a.z2[(int64_t)row * a.Z + j] * a.z[(int64_t)row * a.Z + j];
      sd = warp_sum(sd); zz = warp_sum(zz);
      const float a_diff = sd / (float)A;
      const float z_diff = 0.5f - 0.5f * zz;
      const float bonus = a_diff / (z_diff + 1e-5f);
      div_l = (a.div_tar - bonus) * (a.div_tar - bonus);
      const float dL_da = -2.0f * (a.div_tar - bonus) / (z_diff + 1e-5f);
      coef_div = a.div_bonus * wrow * dL_da * (2.0f / (float)A);
    }
    float* dmu = a.dmu + (int64_t)row * a.ld_mu;
    float* dmu2 = a.has_div ? a.dmu + (int64_t)(a.B + row) * a.ld_mu : nullptr;
    const float cb = a.bounds_coef * wrow;
    for (int j = lane; j < A; j += 32) {
      const float sg = expf(a.logstd[j]);
      float g = g_nlp * (-(act[j] - mu[j]) / (sg * sg));
      g += cb * (2.0f * fminf(mu[j] + 1.0f, 0.0f) + 2.0f * fmaxf(mu[j] - 1.0f, 0.0f));
      if (a.has_div) {
        const float d = fminf(fmaxf(mu[j], -1.0f), 1.0f) - fminf(fmaxf(mu2[j], -1.0f), 1.0f);
        if (mu[j] >= -1.0f && mu[j] <= 1.0f) g += coef_div * d;
        dmu2[j] = (mu2[j] >= -1.0f && mu2[j] <= 1.0f) ? -coef_div * d : 0.0f;
      }
      if (a.mu_tanh) g *= (1.0f - mu[j] * mu[j]);
      dmu[j] = g;
    }
    if (lane == 0) {
      const float v = a.values[row], R = a.returns[row];
      a.dv[row] = a.critic_coef * 2.0f * (v - R) / (float)a.B;
      part[0] = (double)(m * a_loss); part[1] = (double)((R - v) * (R - v)); part[2] = (double)(m * bl);
      part[3] = (double)(m * clipped); part[4] = (double)kl; part[5] = (double)(m * div_l);
    }
  }
  block_sum<6>(part, sm);
  if (threadIdx.x == 0) {
    atomicAdd(&a.acc[ACC_ALOSS], part[0]); atomicAdd(&a.acc[ACC_CLOSS], part[1]); atomicAdd(&a.acc[ACC_BLOSS], part[2]);
    atomicAdd(&a.acc[ACC_CLIPPED], part[3]); atomicAdd(&a.acc[ACC_KL], part[4]); atomicAdd(&a.acc[ACC_DIV], part[5]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Discriminator BCE heads (amp_agent.py:442-446,481-489): rows [0,2Ba) agent+replay -> label 0, [2Ba,3Ba) demo -> 1
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(256)
disc_head_kernel(const float* __restrict__ logit, int Ba, float disc_coef, float* __restrict__ dlogit, double* __restrict__ acc,
                 float* __restrict__ out_agent, float* __restrict__ out_demo) {
  __shared__ double sm[32 * 6];
  double part[6] = {0, 0, 0, 0, 0, 0};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * Ba) {
    const float l = logit[i];
    const float c = disc_coef * 0.5f;
    if (i < 2 * Ba) {
      part[0] = (double)softplusf(l);
      dlogit[i] = c * (1.0f / (1.0f + expf(-l))) / (float)(2 * Ba);
      part[2] = (l < 0.0f) ? 1.0 : 0.0; part[4] = (double)l;
      if (out_agent) out_agent[i] = l;
    } else {
      part[1] = (double)softplusf(-l);
      dlogit[i] = -c * (1.0f / (1.0f + expf(l))) / (float)Ba;
      part[3] = (l > 0.0f) ? 1.0 : 0.0; part[5] = (double)l;
      if (out_demo) out_demo[i - 2 * Ba] = l;
    }
  }
  block_sum<6>(part, sm);
  if (threadIdx.x == 0) {
    atomicAdd(&acc[ACC_BCE_AGENT], part[0]); atomicAdd(&acc[ACC_BCE_DEMO], part[1]); atomicAdd(&acc[ACC_ACC_AGENT], part[2]);
    atomicAdd(&acc[ACC_ACC_DEMO], part[3]); atomicAdd(&acc[ACC_LOGIT_AGENT], part[4]); atomicAdd(&acc[ACC_LOGIT_DEMO], part[5]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Encoder head: F.normalize (eps 1e-12) + enc loss -mean(z^ . z) (ase_agent.py:413-443,469-472) and its gradient.
// One warp per row.  With de == nullptr it only normalises (inference, ase_network_builder.py:214-219).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
enc_head_kernel(const float* __restrict__ e, int rows, int Z, const float* __restrict__ z, float enc_coef,
                float* __restrict__ de, float* __restrict__ enc_pred, double* __restrict__ acc) {
  __shared__ double sm[32];
  const int lane = threadIdx.x & 31;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double part[1] = {0.0};
  if (row < rows) {
    const float* er = e + (int64_t)row * Z;
    float n2 = 0.0f;
    for (int j = lane; j < Z; j += 32) n2 += er[j] * er[j];
    n2 = warp_sum(n2);
    const float n = sqrtf(n2), d = fmaxf(n, 1e-12f);
    if (enc_pred) for (int j = lane; j < Z; j += 32) enc_pred[(int64_t)row * Z + j] = er[j] / d;
    if (de) {
      const float* zr = z + (int64_t)row * Z;
      const float gc = -enc_coef / (float)rows;
      float dot = 0.0f;
      for (int j = lane; j < Z; j += 32) dot += (er[j] / d) * zr[j];
      dot = warp_sum(dot);
      const float pg = gc * dot;    // sum_j zh_j * g_j with g = gc * z
      for (int j = lane; j < Z; j += 32) {
        const float zh = er[j] / d, g = gc * zr[j];
        de[(int64_t)row * Z + j] = (n > 1e-12f) ? (g - zh * pg) / d : g / d;
      }
      if (lane == 0) part[0] = (double)(-dot);
    }
  }
  if (acc) {
    block_sum<1>(part, sm);
    if (threadIdx.x == 0) atomicAdd(&acc[ACC_ENC], part[0]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Gradient penalty element-wise pieces (amp_agent.py:454-459; analytic double backward, DESIGN.md "GP")
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gp_u_last_kernel(const float* __restrict__ h, int64_t ldh, int rows, int cols, const float* __restrict__ w, float* __restrict__ u) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    u[i] = (h[(int64_t)r * ldh + c] > 0.0f) ? w[c] : 0.0f;
  }
}

__global__ void __launch_bounds__(256)
gp_scale_kernel(float* __restrict__ g, int64_t total, float scale, double* __restrict__ acc) {
  __shared__ double sm[32];
  double part[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    part[0] += (double)v * (double)v;
    g[i] = v * scale;
  }
  block_sum<1>(part, sm);
  if (threadIdx.x == 0) atomicAdd(&acc[ACC_GP], part[0]);
}

// db[n] += sum_m dz[m, n]
__global__ void __launch_bounds__(128)
colsum_kernel(const float* __restrict__ dz, int64_t ld, int rows, int cols, int rows_per_block, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  // fp64 partial sums: a 256-term fp32 chain loses ~1e-5 of sum|dz| per partial, which on bias gradients with heavy cancellation
  // (critic trunk at B = 16384) showed up as 1e-4 of max|db| against the reference's tree reduction; the adds hide under the loads
  double s = 0.0;
  for (int r = r0; r < r1; ++r) s += (double)dz[(int64_t)r * ld + c];
  atomicAdd(&db[c], (float)s);
}

// Narrow matrices (head gradients: 31 / 1 / 64 columns, contiguous rows): the column of a thread is fixed by making the grid stride a
// multiple of cols, every thread streams the flat array (fully coalesced), per-column partial sums meet in shared memory.
__global__ void __launch_bounds__(256)
colsum_narrow_kernel(const float* __restrict__ dz, int64_t total, int cols, int64_t stride, float* __restrict__ db) {
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = 0.0;
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < stride) {
    double s = 0.0;
    for (int64_t i = t; i < total; i += stride) s += (double)dz[i];
    atomicAdd(&sm[(int)(t % cols)], s);
  }
  __syncthreads();
  if (threadIdx.x < cols) atomicAdd(&db[threadIdx.x], (float)sm[threadIdx.x]);
}

// grads += coef * w ; acc[idx] += sum w^2   (logit-weight regulariser and disc weight decay, amp_agent.py:448-466)
__global__ void __launch_bounds__(256)
weight_reg_kernel(const float* __restrict__ w, float* __restrict__ g, int64_t n, float coef, double* __restrict__ acc, int idx, int idx2) {
  __shared__ double sm[32];
  double part[1] = {0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    part[0] += (double)v * (double)v;
    g[i] += coef * v;
  }
  block_sum<1>(part, sm);
  if (threadIdx.x == 0) {
    atomicAdd(&acc[idx], part[0]);
    if (idx2 >= 0) atomicAdd(&acc[idx2], part[0]);
  }
}

__global__ void finalize_scalars_kernel(FinalizeArgs f) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double* a = f.acc;
  float* o = f.out;
  const double msum = a[ACC_MSUM], B = (double)f.B, Ba = (double)f.Ba;
  const double actor = a[ACC_ALOSS] / msum, critic = a[ACC_CLOSS] / B, bl = a[ACC_BLOSS] / msum;
  double ent = 0.0;
  for (int j = 0; j < f.A; ++j) ent += (double)(0.5f + 0.5f * 1.8378770664093453f + f.logstd[j]);
  double total = actor + f.critic_coef * critic - f.entropy_coef * ent + f.bounds_coef * bl;
  o[ASE_TR_ACTOR_LOSS] = (float)actor; o[ASE_TR_CRITIC_LOSS] = (float)critic; o[ASE_TR_B_LOSS] = (float)bl;
  o[ASE_TR_ENTROPY] = (float)ent; o[ASE_TR_CLIP_FRAC] = (float)(a[ACC_CLIPPED] / msum); o[ASE_TR_KL] = (float)(a[ACC_KL] / B);
  for (int i = ASE_TR_DISC_LOSS; i < ASE_TR_COUNT; ++i) o[i] = 0.0f;
  if (f.kind != ASE_KIND_PPO) {
    const double gp = a[ACC_GP] / Ba;
    const double disc = 0.5 * (a[ACC_BCE_AGENT] / (2.0 * Ba) + a[ACC_BCE_DEMO] / Ba) + f.logit_reg * a[ACC_WLOGIT2] +
                        f.gp_coef * gp + f.weight_decay * a[ACC_WDISC2];
    o[ASE_TR_DISC_LOSS] = (float)disc; o[ASE_TR_DISC_GRAD_PENALTY] = (float)gp; o[ASE_TR_DISC_LOGIT_LOSS] = (float)a[ACC_WLOGIT2];
    o[ASE_TR_DISC_AGENT_ACC] = (float)(a[ACC_ACC_AGENT] / (2.0 * Ba)); o[ASE_TR_DISC_DEMO_ACC] = (float)(a[ACC_ACC_DEMO] / Ba);
    o[ASE_TR_DISC_AGENT_LOGIT_MEAN] = (float)(a[ACC_LOGIT_AGENT] / (2.0 * Ba)); o[ASE_TR_DISC_DEMO_LOGIT_MEAN] = (float)(a[ACC_LOGIT_DEMO] / Ba);
    total += f.disc_coef * disc;
  }
  if (f.kind == ASE_KIND_ASE) {
    const double enc = a[ACC_ENC] / Ba, div = a[ACC_DIV] / msum;
    o[ASE_TR_ENC_LOSS] = (float)enc; o[ASE_TR_DIVERSITY_LOSS] = (float)div;
    total += f.enc_coef * enc + f.div_bonus * div;
  }
  o[ASE_TR_TOTAL_LOSS] = (float)total;
}

// torch.optim.Adam single-tensor path (common_agent.py:45): denom = sqrt(v)/sqrt(bc2) + eps; p -= lr/bc1 * m/denom
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float grad_scale, float b1, float b2, float omb1, float omb2, float step_size, float bc2_sqrt, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = b1 * m[i] + omb1 * gi;             // exp_avg.mul_(beta1).add_(grad, alpha=1-beta1)
    const float vi = b2 * v[i] + omb2 * gi * gi;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// -------------------------------------------------------------------------------------------- launchers
static inline int ew_blocks(int64_t n) { return (int)imin64((n + 255) / 256, 148 * 16); }

int launch_mask_sum(const float* mask, int rows, double* acc, cudaStream_t st) {
  mask_sum_kernel<<<min(ceil_div(rows, 256), 148), 256, 0, st>>>(mask, rows, acc);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_ppo_head(const PpoHeadArgs& a, cudaStream_t st) {
  ppo_head_kernel<<<ceil_div((int64_t)a.B * 32, 256), 256, 0, st>>>(a);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_disc_head(const float* logit, int Ba, float disc_coef, float* dlogit, double* acc, float* out_agent, float* out_demo, cudaStream_t st) {
  disc_head_kernel<<<ceil_div(3 * Ba, 256), 256, 0, st>>>(logit, Ba, disc_coef, dlogit, acc, out_agent, out_demo);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_enc_head(const float* e, int rows, int Z, const float* z, float enc_coef, float* de, float* enc_pred, double* acc, cudaStream_t st) {
  enc_head_kernel<<<ceil_div((int64_t)rows * 32, 256), 256, 0, st>>>(e, rows, Z, z, enc_coef, de, enc_pred, acc);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_gp_u_last(const float* h, int64_t ldh, int rows, int cols, const float* w, float* u, cudaStream_t st) {
  gp_u_last_kernel<<<ew_blocks((int64_t)rows * cols), 256, 0, st>>>(h, ldh, rows, cols, w, u);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_gp_scale(float* g, int64_t total, float scale, double* acc, cudaStream_t st) {
  gp_scale_kernel<<<ew_blocks(total), 256, 0, st>>>(g, total, scale, acc);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_colsum(const float* dz, int64_t ld, int rows, int cols, float* db, cudaStream_t st) {
  if (cols <= 64 && ld == cols) {
    const int64_t total = (int64_t)rows * cols;
    const int blocks = (int)imin64((total + 256 * 8 - 1) / (256 * 8), 148 * 4);
    const int64_t stride = (int64_t)blocks * 256 / cols * cols;        // a multiple of cols: thread t always sees column t % cols
    colsum_narrow_kernel<<<blocks, 256, 0, st>>>(dz, total, cols, stride, db);
    ASE_LAUNCH_OK(); return ASE_OK;
  }
  const int rpb = 256;
  dim3 grid(ceil_div(cols, 128), ceil_div(rows, rpb));
  colsum_kernel<<<grid, 128, 0, st>>>(dz, ld, rows, cols, rpb, db);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_weight_reg(const float* w, float* g, int64_t n, float coef, double* acc, int idx, int idx2, cudaStream_t st) {
  weight_reg_kernel<<<ew_blocks(n), 256, 0, st>>>(w, g, n, coef, acc, idx, idx2);
  ASE_LAUNCH_OK(); return ASE_OK;
}
int launch_finalize(const FinalizeArgs& f, cudaStream_t st) {
  finalize_scalars_kernel<<<1, 32, 0, st>>>(f);
  ASE_LAUNCH_OK(); return ASE_OK;
}
AdamConsts adam_consts(float b1, float b2, float lr, float eps, int64_t step) {
  // torch.optim.Adam evaluates 1-beta, the bias corrections and lr/bc1 in Python doubles from the decimal hyper-parameters
  // (0.9, 0.999, 2e-5) and only then rounds to fp32: recover those decimals from the fp32 config values
  auto dec = [](float x) { char buf[32]; snprintf(buf, sizeof(buf), "%.7g", (double)x); return strtod(buf, nullptr); };
  const double db1 = dec(b1), db2 = dec(b2), dlr = dec(lr), deps = dec(eps);
  const double bc1 = 1.0 - pow(db1, (double)step), bc2 = 1.0 - pow(db2, (double)step);
  AdamConsts c;
  c.b1 = (float)db1; c.b2 = (float)db2; c.omb1 = (float)(1.0 - db1); c.omb2 = (float)(1.0 - db2);
  c.step_size = (float)(dlr / bc1); c.bc2_sqrt = (float)sqrt(bc2); c.eps = (float)deps;
  return c;
}
int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float grad_scale, float b1, float b2, float lr, float eps,
                int64_t step, cudaStream_t st) {
  const AdamConsts c = adam_consts(b1, b2, lr, eps, step);
  adam_kernel<<<ew_blocks(n), 256, 0, st>>>(p, g, m, v, n, grad_scale, c.b1, c.b2, c.omb1, c.omb2, c.step_size, c.bc2_sqrt, c.eps);
  ASE_LAUNCH_OK(); return ASE_OK;
}

}  // namespace ase
