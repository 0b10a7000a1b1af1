// Development probe (run on the GPU box): tensor-pipe cost of the tcgen05.mma forms the kernels issue — one CTA per SM, one thread issues
// REP x (a batch of identical MMAs + commit), clock64 around the batch after waiting for its completion. Operand contents are irrelevant.
//   K-major SS  (likelihood MMA-1, surrogate layer 1):   M=128, N in {32,48,64,128,208,256}, K=16, bf16 / fp16
//   TS          (likelihood MMA-2, A in tensor memory):  N in {32,48}
//   MN-major SS (batch-Langevin G1 / G2, K = data rows): N in {64,128}
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o devtools/mma_cost devtools/mma_cost.cu -lcuda ; run: devtools/mma_cost
#include <cstdio>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../sapt_b200/csrc/sapt_tc_ptx.cuh"
using namespace sapt;

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int kRep = 100;

// form: 0 = K-major SS bf16, 1 = K-major SS fp16, 2 = TS (A in TMEM), 3 = MN-major SS
// ISSUE: 0 = `if (tid == 0)` (how the kernels issued until round 2: ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY
//            loop, ~13 dependent instructions per MMA), 1 = warp 0 stays converged and an elected lane issues (elect.sync): the descriptors
//            stay on the uniform datapath and the UTCHMMA are issued back to back.
// NACC accumulators are used round-robin (consecutive MMAs into the same accumulator depend on one another).
// Everything is a compile-time constant and the batch is fully unrolled, so that the probe measures the tensor pipe and not its own loop
// (the first version of this probe — runtime form / batch — measured ~110 cycles per MMA for every shape: its own issue cost).
template <int FORM, int N, int BATCH, int ISSUE, int NACC>
__global__ void __launch_bounds__(128) cost_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // finite values in either format
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (ISSUE == 0 ? (tid == 0) : (warp == 0)) {
    constexpr uint32_t fmt = (FORM == 1) ? 0u : ((1u << 7) | (1u << 10));
    constexpr uint32_t mn = (FORM == 3) ? ((1u << 15) | (1u << 16)) : 0u;
    constexpr uint32_t idesc = (1u << 4) | fmt | mn | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 64 * 1024;
    long long best = 1ll << 60, tot = 0;
    for (int r = 0; r < kRep; ++r) {
      const long long t0 = clock64();
      if (ISSUE == 0 || elect_one()) {
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
          const uint32_t acc = tm + (uint32_t)((k % NACC) * (FORM == 2 ? 64 : (N > 128 ? 256 : 128)));
          if constexpr (FORM == 2) tc_mma_bf16_ts(acc, tm + 256 + (k & 7) * 8, umma_desc_sw128(b + (k & 3) * 32), idesc, 1u);
          else if constexpr (FORM == 3) tc_mma_bf16(acc, desc_mn(a + (k & 7) * 2048, 16384), desc_mn(b + (k & 7) * 2048, 16384), idesc, 1u);
          else tc_mma_bf16(acc, umma_desc_sw128(a + (k & 3) * 32), umma_desc_sw128(b + (k & 3) * 32), idesc, 1u);
        }
        tc_commit(&bar);
      }
      if (ISSUE) __syncwarp();
      mbar_wait(&bar, r & 1);
      const long long dt = clock64() - t0;
      tot += dt; if (dt < best) best = dt;
    }
    if (blockIdx.x == 0 && (tid & 31) == 0) { out[0] = best; out[1] = tot / kRep; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512));
}

static long long* d_out;
static const int kSmem = 161 * 1024 + 1024;
template <int FORM, int N, int ISSUE, int NACC>
static int run() {
  const char* names[4] = {"K-major SS bf16", "K-major SS fp16", "TS (A in TMEM)", "MN-major SS"};
  long long t[2] = {0, 0};
  cudaError_t e = cudaSuccess;
  for (int bi = 0; bi < 2; ++bi) {     // two batch sizes: the difference removes the commit / mbarrier / wake-up overhead
    if (bi) { cudaFuncSetAttribute(cost_kernel<FORM, N, 64, ISSUE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); cost_kernel<FORM, N, 64, ISSUE, NACC><<<148, 128, kSmem>>>(d_out); }
    else { cudaFuncSetAttribute(cost_kernel<FORM, N, 16, ISSUE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); cost_kernel<FORM, N, 16, ISSUE, NACC><<<148, 128, kSmem>>>(d_out); }
    e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
    t[bi] = h[0];
    if (e != cudaSuccess) break;
  }
  printf("%-12s acc=%d %-16s M=128 N=%3d K=16: %6.1f cycles per MMA (incremental: (T64 - T16) / 48), fixed overhead per batch %6.0f; T16 %lld T64 %lld; %s\n",
         ISSUE ? "elect.sync" : "if(lane==0)", NACC, names[FORM], N, (double)(t[1] - t[0]) / 48.0, (double)t[0] - 16.0 * (double)(t[1] - t[0]) / 48.0, t[0], t[1],
         cudaGetErrorString(e));
  return e != cudaSuccess;
}
template <int ISSUE, int NACC>
static int run_forms() {
  int rc = 0;
  rc |= run<0, 32, ISSUE, NACC>(); rc |= run<0, 64, ISSUE, NACC>(); rc |= run<0, 128, ISSUE, NACC>();
  if (NACC <= 2) rc |= run<0, 256, ISSUE, NACC>();
  rc |= run<1, 64, ISSUE, NACC>(); rc |= run<1, 128, ISSUE, NACC>();
  rc |= run<2, 32, ISSUE, NACC>(); rc |= run<2, 48, ISSUE, NACC>(); rc |= run<2, 64, ISSUE, NACC>();
  rc |= run<3, 64, ISSUE, NACC>(); rc |= run<3, 128, ISSUE, NACC>();
  return rc;
}

int main() {
  cudaMalloc(&d_out, 16);
  int rc = 0;
  rc |= run_forms<0, 1>();
  rc |= run_forms<1, 1>();
  rc |= run_forms<1, 2>();
  rc |= run_forms<1, 4>();
  return rc;
}
