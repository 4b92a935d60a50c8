// Shared device/host helpers for libdstagnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dstagnn_b200.h"

namespace dstag {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DSTAG_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::dstag::set_error(__VA_ARGS__);        \
      return DSTAG_ERR_BAD_ARG;               \
    }                                         \
  } while (0)

// Call right after a kernel launch: returns the cudaError_t (positive) on failure.
#define DSTAG_RETURN_LAUNCH(name)                                                   \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::dstag::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return (int)e__;                                                              \
    }                                                                               \
    ::dstag::count_launch();                                                        \
    return 0;                                                                       \
  } while (0)

#define DSTAG_CHECK_LAUNCH(name)                                                    \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::dstag::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return (int)e__;                                                              \
    }                                                                               \
    ::dstag::count_launch();                                                        \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- warp / block reductions ----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` is a shared array of >= 33 floats.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` against a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// ---- counter-based RNG (Philox4x32-10) for dropout -------------------------------------
// Element e of a tensor draws word (e & 3) of philox(counter = e >> 2, key = seed).  The
// forward and backward kernels regenerate the same mask from (seed, e); nothing is stored.
__device__ __forceinline__ uint4 philox4x32(uint64_t ctr, uint64_t seed) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x2B992DDFu, c3 = 0xA0F2F5B1u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout key of a launch: the host-side per-site key, advanced by a DEVICE-resident step counter when one
// is given.  A training step captured in a CUDA graph replays the same kernel arguments every time; the
// counter (bumped once per step by dstag_tick, inside the graph) is what makes the masks differ per step.
__device__ __forceinline__ unsigned long long fold_seed(unsigned long long seed, const unsigned long long* seed_dev) {
  return seed_dev ? seed + 0x9E3779B97F4A7C15ull * (*seed_dev + 1ull) : seed;
}

// Inverted-dropout multiplier for element e: 0 with probability p, else 1/(1-p).
__device__ __forceinline__ float dropout_scale(uint64_t e, uint64_t seed, float p, float inv_keep) {
  const uint4 r = philox4x32(e >> 2, seed);
  const uint32_t w = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
  const float u = (float)(w >> 8) * (1.0f / 16777216.0f);  // [0,1)
  return u < p ? 0.f : inv_keep;
}

// Four consecutive elements e0 .. e0+3 with e0 % 4 == 0 share one Philox evaluation (same mask as four
// dropout_scale() calls at a quarter of the integer work -- the EmbedS LayerNorm was bound by it).
__device__ __forceinline__ void dropout_scale4(uint64_t e0, uint64_t seed, float p, float inv_keep, float (&m)[4]) {
  const uint4 r = philox4x32(e0 >> 2, seed);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = ((float)(w[j] >> 8) * (1.0f / 16777216.0f)) < p ? 0.f : inv_keep;
}

}  // namespace dstag
