// Device-side helpers: Philox4x32-10, uniform/normal transforms, reductions, fast sigmoid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sapt {

// ---- Philox4x32-10 (Salmon et al., SC'11). Integer-exact twin of oracle.sapt_oracle.philox4x32_10.
__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)c[0] * 0xD2511F53ull;
  const uint64_t p1 = (uint64_t)c[2] * 0xCD9E8D57ull;
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0;
  const uint32_t n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__host__ __device__ __forceinline__ void philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// (x + 0.5) * 2^-32, one rounding: in (0, 1]
__device__ __forceinline__ float u32_to_uniform(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// block of 4 values for (seed, kind, step, chain, block index)
__device__ __forceinline__ void philox_block(uint32_t seed, int kind, int step, uint32_t chain, uint32_t blk, uint32_t out[4]) {
  const uint32_t ctr[4] = {blk, (uint32_t)step, (uint32_t)kind, 0u};
  philox4x32_10(ctr, seed, chain, out);
}

__device__ __forceinline__ float philox_uniform(uint32_t seed, int kind, int step, uint32_t chain, uint32_t index) {
  uint32_t o[4];
  philox_block(seed, kind, step, chain, index >> 2, o);
  return u32_to_uniform(o[index & 3]);
}

// 4 standard normals from one Philox block (two Box-Muller pairs: r0cos, r0sin, r1cos, r1sin)
__device__ __forceinline__ void philox_normal4(uint32_t seed, int kind, int step, uint32_t chain, uint32_t blk, float z[4]) {
  uint32_t o[4];
  philox_block(seed, kind, step, chain, blk, o);
  const float u0 = u32_to_uniform(o[0]), u1 = u32_to_uniform(o[1]);
  const float u2 = u32_to_uniform(o[2]), u3 = u32_to_uniform(o[3]);
  const float r0 = sqrtf(-2.0f * logf(u0));
  const float r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0, s1, c1;
  sincospif(2.0f * u1, &s0, &c0);
  sincospif(2.0f * u3, &s1, &c1);
  z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

// ---- reductions -------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sums v over the block; result valid in every thread. scratch: >= 32 elements of T in shared.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T t = (lane < nw) ? scratch[lane] : T(0);
  t = warp_sum(t);
  return t;
}

// ---- math --------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigmoid of z given t = -z*log2(e):  1 / (1 + 2^t)   (2 MUFU + 1 FADD)
__device__ __forceinline__ float sigmoid_from_neg_log2e(float t) { return rcp_approx(1.0f + ex2_approx(t)); }

// prior: -(P/2) ln s2 - sum(w^2)/(2 s2)   (_revamp.py:404-410)
__device__ __forceinline__ double prior_from_sumsq(double sumsq, int P, double s2) {
  return -1.0 * ((double)P / 2.0) * log(s2) - (1.0 / (2.0 * s2)) * sumsq;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace sapt
