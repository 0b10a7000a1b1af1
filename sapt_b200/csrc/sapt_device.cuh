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

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 process two fp32 lanes per issue slot; the epilogues are bound by
// instruction issue and the SFU pipe, not by the FP32 pipe) ------------------------------------------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// Four sigmoids s_i = scale / (1 + 2^min(t_i, lim)) with one reciprocal, packed: the results come out as the pairs
// s10 = (s1, s0) and s32 = (s3, s2) (low lane first) — 4 FMNMX + 4 MUFU.EX2 + 2 FADD2 + 3 FMUL + MUFU.RCP + 3 FMUL2 (+1 FMUL for a scale)
// (e0_out: the first 2^t, for callers that order successive groups on it)
template <bool SCALED>
__device__ __forceinline__ void sigmoid4_packed_e(float t0, float t1, float t2, float t3, float lim, float scale, f32x2& s10, f32x2& s32, float& e0_out) {
  const float e0 = ex2_approx(fminf(t0, lim)), e1 = ex2_approx(fminf(t1, lim));
  e0_out = e0;
  const float e2 = ex2_approx(fminf(t2, lim)), e3 = ex2_approx(fminf(t3, lim));
  const f32x2 one = pack2(1.0f, 1.0f);
  const f32x2 a01 = add2(pack2(e0, e1), one), a23 = add2(pack2(e2, e3), one);
  float a0, a1, a2, a3;
  unpack2(a01, a0, a1); unpack2(a23, a2, a3);
  const float ab = a0 * a1, cd = a2 * a3;
  float r = rcp_approx(ab * cd);
  if constexpr (SCALED) r *= scale;
  float rcd, rab;
  unpack2(mul2(pack2(cd, ab), pack2(r, r)), rcd, rab);
  s10 = mul2(a01, pack2(rcd, rcd));   // (a0 r cd, a1 r cd) = (s1, s0)
  s32 = mul2(a23, pack2(rab, rab));   // (a2 r ab, a3 r ab) = (s3, s2)
}
template <bool SCALED>
__device__ __forceinline__ void sigmoid4_packed(float t0, float t1, float t2, float t3, float lim, float scale, f32x2& s10, f32x2& s32) {
  float e0;
  sigmoid4_packed_e<SCALED>(t0, t1, t2, t3, lim, scale, s10, s32, e0);
}

// four sigmoids with ONE reciprocal:  s_i = 1 / a_i,  a_i = 1 + 2^min(t_i, 30)  (the clamp costs < 1e-9 and keeps the product of
// four below 2^120):  1/a0 = (1/(a0 a1 a2 a3)) a1 a2 a3, ...  = 4 MUFU.EX2 + 1 MUFU.RCP + 9 FMUL instead of 8 MUFU
__device__ __forceinline__ void sigmoid4_from_neg_log2e(float t0, float t1, float t2, float t3, float& s0, float& s1, float& s2, float& s3) {
  const float a0 = 1.0f + ex2_approx(fminf(t0, 30.0f));
  const float a1 = 1.0f + ex2_approx(fminf(t1, 30.0f));
  const float a2 = 1.0f + ex2_approx(fminf(t2, 30.0f));
  const float a3 = 1.0f + ex2_approx(fminf(t3, 30.0f));
  const float ab = a0 * a1, cd = a2 * a3;
  const float r = rcp_approx(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  s0 = rab * a1; s1 = rab * a0; s2 = rcd * a3; s3 = rcd * a2;
}

// Two fp32 values -> two packed bf16 pairs (low half = a, high half = b) with v = hi + lo to 2^-17 |v|, integer
// instructions only: hi = v truncated to bf16 (one PRMT packs the pair), lo = the exact residual v - hi (same sign,
// < 2^-8 |v|) with its magnitude rounded half-up to bf16. The conversion instructions (F2FP, cvt.rn.bf16x2.f32) issue on
// the SFU pipe like MUFU (measured: sapt_measure_pipes), which is the pipe these kernels are short of.
__device__ __forceinline__ void split_pack_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  const uint32_t ra = __float_as_uint(a - __uint_as_float(ua & 0xffff0000u)) + 0x8000u;
  const uint32_t rb = __float_as_uint(b - __uint_as_float(ub & 0xffff0000u)) + 0x8000u;
  lo = __byte_perm(ra, rb, 0x7632);
}

// prior: -(P/2) ln s2 - sum(w^2)/(2 s2)   (_revamp.py:404-410)
__device__ __forceinline__ double prior_from_sumsq(double sumsq, int P, double s2) {
  return -1.0 * ((double)P / 2.0) * log(s2) - (1.0 / (2.0 * s2)) * sumsq;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace sapt
