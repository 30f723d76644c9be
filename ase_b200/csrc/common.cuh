// Shared helpers for the ase_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/ase_b200.h"

namespace ase {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ASE_CHECK_ARG(cond, ...)                         \
  do {                                                   \
    if (!(cond)) {                                       \
      ase::set_error(__VA_ARGS__);                       \
      return ASE_ERR_INVALID;                            \
    }                                                    \
  } while (0)

#define ASE_CUDA_OK(expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ase::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ASE_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

// call after every kernel launch: catches launch-configuration errors without synchronising
#define ASE_LAUNCH_OK()                                                              \
  do {                                                                               \
    ase::count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ase::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ASE_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline int64_t align_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of NV doubles per thread, result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* >= 32*NV doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[warp * NV + i] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = (lane < nwarps) ? smem[lane * NV + i] : 0.0;
      v[i] = warp_sum(x);
    }
  }
  __syncthreads();
}

struct Quat { float x, y, z, w; };
struct Vec3 { float x, y, z; };

// isaacgym.torch_utils.quat_rotate: v(2w^2-1) + 2w (q x v) + 2 q (q.v)
__device__ __forceinline__ Vec3 quat_rotate(const Quat q, const Vec3 v) {
  const float s = 2.0f * q.w * q.w - 1.0f;
  const float cx = q.y * v.z - q.z * v.y, cy = q.z * v.x - q.x * v.z, cz = q.x * v.y - q.y * v.x;
  const float d = q.x * v.x + q.y * v.y + q.z * v.z;
  Vec3 r;
  r.x = v.x * s + cx * q.w * 2.0f + q.x * d * 2.0f;
  r.y = v.y * s + cy * q.w * 2.0f + q.y * d * 2.0f;
  r.z = v.z * s + cz * q.w * 2.0f + q.z * d * 2.0f;
  return r;
}

// isaacgym.torch_utils.quat_mul (8-multiply factorisation, same operation order as the reference)
__device__ __forceinline__ Quat quat_mul(const Quat a, const Quat b) {
  const float ww = (a.z + a.x) * (b.x + b.y);
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
  Quat r;
  r.w = qq - ww + (a.z - a.y) * (b.y - b.z);
  r.x = qq - xx + (a.x + a.w) * (b.x + b.w);
  r.y = qq - yy + (a.w - a.x) * (b.y + b.z);
  r.z = qq - zz + (a.z + a.y) * (b.w - b.x);
  return r;
}

// quat_from_angle_axis(angle, axis) = quat_unit([normalize(axis) * sin(angle/2), cos(angle/2)])
__device__ __forceinline__ Quat quat_from_angle_axis(float angle, Vec3 axis) {
  const float theta = angle / 2.0f;
  float n = sqrtf(axis.x * axis.x + axis.y * axis.y + axis.z * axis.z);
  n = fmaxf(n, 1e-9f);
  const float s = sinf(theta);
  Quat q;
  q.x = axis.x / n * s; q.y = axis.y / n * s; q.z = axis.z / n * s; q.w = cosf(theta);
  float qn = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  qn = fmaxf(qn, 1e-9f);
  q.x /= qn; q.y /= qn; q.z /= qn; q.w /= qn;
  return q;
}

// utils/torch_utils.py:117-154  heading-inverse quaternion of a root rotation
__device__ __forceinline__ Quat calc_heading_quat_inv(const Quat q) {
  const Vec3 ex = {1.0f, 0.0f, 0.0f};
  const Vec3 d = quat_rotate(q, ex);
  const float heading = atan2f(d.y, d.x);
  const Vec3 ez = {0.0f, 0.0f, 1.0f};
  return quat_from_angle_axis(-heading, ez);
}

// utils/torch_utils.py:46-59  writes 6 floats: rot(q, x^), rot(q, z^)
__device__ __forceinline__ void quat_to_tan_norm(const Quat q, float* out) {
  const Vec3 ex = {1.0f, 0.0f, 0.0f}, ez = {0.0f, 0.0f, 1.0f};
  const Vec3 t = quat_rotate(q, ex), n = quat_rotate(q, ez);
  out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = n.x; out[4] = n.y; out[5] = n.z;
}

}  // namespace ase
