// SM pipe micro-benchmarks: the measured denominators of the "governing" roofline (SURVEY.md 8d says the
// MUFU / FP32 pipe peaks "must be re-measured by microbenchmark"). Every kernel keeps kChains independent
// dependency chains per thread in registers, 1024 threads x 2 CTAs per SM, timed with CUDA events on the
// caller's stream; clock64 around the loop gives the SM clock during the measurement.
//
//   MUFU.EX2, MUFU.RCP : the two SFU opcodes of the likelihood kernels' sigmoids / softmax
//   FFMA (3 registers), FMUL, FADD2/FFMA2 (fma.rn.f32x2, two fp32 lanes per instruction slot)
//   IADD3/LOP3 (ALU pipe)
#include "sapt_device.cuh"
#include "sapt_internal.h"

namespace sapt {

namespace {
constexpr int kChains = 8;
constexpr int kInner = 64;

enum PipeOp { OP_EX2 = 0, OP_RCP = 1, OP_FFMA = 2, OP_FMUL = 3, OP_FFMA2 = 4, OP_LOP3 = 5, OP_MIX = 6, OP_F2FP = 7, OP_COUNT = 8 };

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int OP>
// cycles[0] = SM cycles, cycles[1] = nanoseconds (%globaltimer) block 0 spent in the loop: their ratio is the SM clock whatever
// the residency (the event time covers the whole grid, possibly several waves)
__global__ void __launch_bounds__(1024, 2) pipe_kernel(float* out, long long* cycles, int iters, float a, float b) {
  float v[kChains];
  unsigned long long w[kChains / 2];
  uint32_t q[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) { v[c] = 0.5f + 0.01f * (float)(threadIdx.x + c); q[c] = threadIdx.x * 7u + c; }
#pragma unroll
  for (int c = 0; c < kChains / 2; ++c) asm("mov.b64 %0, {%1, %2};" : "=l"(w[c]) : "f"(v[2 * c]), "f"(v[2 * c + 1]));
  // per-thread multiplier / addend: keeps them out of uniform registers (FFMA with a uniform or immediate operand
  // is a different, faster form than the three-register FFMA the kernels execute)
  a += 1e-9f * (float)threadIdx.x;
  b += 1e-9f * (float)threadIdx.x;
  unsigned long long ab, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ab) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  const unsigned long long g0 = global_timer_ns();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kInner; ++k) {
#pragma unroll
      for (int c = 0; c < kChains; ++c) {
        if constexpr (OP == OP_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));       // 2^x: kept in range by a: see host
        else if constexpr (OP == OP_RCP) { v[c] += b; asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[c])); }   // ptxas folds rcp(rcp(x)): one FADD (FMA pipe) in between
        else if constexpr (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b));
        else if constexpr (OP == OP_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a));
        else if constexpr (OP == OP_FFMA2) { if (c < kChains / 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[c]) : "l"(ab), "l"(bb)); }
        else if constexpr (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[c]) : "r"(q[(c + 1) % kChains]), "r"(0x9E3779B9u));
        else if constexpr (OP == OP_F2FP) {   // cvt.rn.bf16x2.f32 (F2FP.BF16.F32.PACK_AB): which pipe do conversions run on?
          uint32_t pk;
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(v[c]), "f"(v[(c + 1) % kChains]));
          v[c] = __uint_as_float(pk & 0x3f800000u | 0x3f000000u);
        }
        else {  // OP_MIX: one MUFU per four FFMA (roughly the likelihood epilogue's ratio), to see whether the pipes overlap
          if (c % 4 == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[c]));
          else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[c]) : "f"(a), "f"(b));
        }
      }
      if constexpr (OP == OP_EX2) {  // 2^x diverges: negate every round (a free source modifier would be folded; one FMUL per 8 MUFU is not timed out)
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = -v[c];
      }
    }
  }
  const long long t1 = clock64();
  const unsigned long long g1 = global_timer_ns();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kChains; ++c) s += v[c] + __uint_as_float(q[c] & 0x3fffffu);
#pragma unroll
  for (int c = 0; c < kChains / 2; ++c) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(w[c])); s += lo + hi; }
  if (s == 123.456f) out[0] = s;   // keeps the chains alive, practically never true
  if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = (long long)(g1 - g0); }
}

template <int OP>
int run_pipe(int num_sms, int iters, float* d_out, long long* d_cyc, cudaStream_t st, double* ops_per_s, double* sm_hz) {
  const int grid = num_sms * 2;
  cudaEvent_t e0, e1;
  SAPT_CUDA(cudaEventCreate(&e0));
  SAPT_CUDA(cudaEventCreate(&e1));
  pipe_kernel<OP><<<grid, 1024, 0, st>>>(d_out, d_cyc, iters / 8 + 1, 0.999f, 0.001f);   // warm-up
  SAPT_CHECK_LAUNCH();
  SAPT_CUDA(cudaEventRecord(e0, st));
  pipe_kernel<OP><<<grid, 1024, 0, st>>>(d_out, d_cyc, iters, 0.999f, 0.001f);
  SAPT_CHECK_LAUNCH();
  SAPT_CUDA(cudaEventRecord(e1, st));
  SAPT_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  SAPT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  long long cyc[2] = {0, 1};
  SAPT_CUDA(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double per_thread = (double)iters * kInner * ((OP == OP_FFMA2) ? kChains / 2 : kChains);
  *ops_per_s = per_thread * 1024.0 * grid / (ms * 1e-3);   // warp-lane operations (an f32x2 instruction counts once)
  *sm_hz = (double)cyc[0] / ((double)cyc[1] * 1e-9);        // SM cycles per second of block 0's loop (clock64 over %globaltimer)
  return SAPT_OK;
}
}  // namespace

}  // namespace sapt

extern "C" int sapt_measure_pipes(int device, double* out, int n_out, void* stream) {
  using namespace sapt;
  SAPT_REQUIRE(out != nullptr && n_out >= SAPT_PIPE_COUNT, "out must hold SAPT_PIPE_COUNT doubles");
  SAPT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SAPT_CUDA(cudaGetDeviceProperties(&prop, device));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* d_out = nullptr; long long* d_cyc = nullptr;
  SAPT_CUDA(cudaMalloc(&d_out, 64));
  SAPT_CUDA(cudaMalloc(&d_cyc, 64));
  const int sms = prop.multiProcessorCount;
  double ops[OP_COUNT], hz[OP_COUNT];
  int rc = SAPT_OK;
  const int iters = 400;   // a few ms per kernel
  if (!rc) rc = run_pipe<OP_EX2>(sms, iters, d_out, d_cyc, st, &ops[OP_EX2], &hz[OP_EX2]);
  if (!rc) rc = run_pipe<OP_RCP>(sms, iters, d_out, d_cyc, st, &ops[OP_RCP], &hz[OP_RCP]);
  if (!rc) rc = run_pipe<OP_FFMA>(sms, 4 * iters, d_out, d_cyc, st, &ops[OP_FFMA], &hz[OP_FFMA]);
  if (!rc) rc = run_pipe<OP_FMUL>(sms, 4 * iters, d_out, d_cyc, st, &ops[OP_FMUL], &hz[OP_FMUL]);
  if (!rc) rc = run_pipe<OP_FFMA2>(sms, 4 * iters, d_out, d_cyc, st, &ops[OP_FFMA2], &hz[OP_FFMA2]);
  if (!rc) rc = run_pipe<OP_LOP3>(sms, 4 * iters, d_out, d_cyc, st, &ops[OP_LOP3], &hz[OP_LOP3]);
  if (!rc) rc = run_pipe<OP_MIX>(sms, 2 * iters, d_out, d_cyc, st, &ops[OP_MIX], &hz[OP_MIX]);
  if (!rc) rc = run_pipe<OP_F2FP>(sms, iters, d_out, d_cyc, st, &ops[OP_F2FP], &hz[OP_F2FP]);
  cudaFree(d_out); cudaFree(d_cyc);
  if (rc) return rc;
  out[SAPT_PIPE_SMS] = sms;
  out[SAPT_PIPE_SM_HZ] = hz[OP_FFMA];
  out[SAPT_PIPE_EX2] = ops[OP_EX2];
  out[SAPT_PIPE_RCP] = ops[OP_RCP];
  out[SAPT_PIPE_FFMA] = ops[OP_FFMA];
  out[SAPT_PIPE_FMUL] = ops[OP_FMUL];
  out[SAPT_PIPE_FFMA2] = ops[OP_FFMA2];
  out[SAPT_PIPE_LOP3] = ops[OP_LOP3];
  out[SAPT_PIPE_MIX] = ops[OP_MIX];
  out[SAPT_PIPE_SM_HZ_MUFU] = hz[OP_EX2];
  out[SAPT_PIPE_F2FP] = ops[OP_F2FP];
  return SAPT_OK;
}
