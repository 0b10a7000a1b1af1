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
// issue: 0 = `if (tid == 0)` (how the kernels issued until round 2: ptxas wraps every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop),
//        1 = warp 0 stays converged and an elected lane issues (elect.sync): descriptors stay on the uniform datapath
__global__ void __launch_bounds__(128) cost_kernel(int form, int N, int kBatch, int issue, int nacc, long long* out) {
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
  if (issue == 0 ? (tid == 0) : (warp == 0)) {
    const uint32_t fmt = (form == 1) ? 0u : ((1u << 7) | (1u << 10));
    const uint32_t mn = (form == 3) ? ((1u << 15) | (1u << 16)) : 0u;
    const uint32_t idesc = (1u << 4) | fmt | mn | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a = smem_u32(sm), b = smem_u32(sm) + 64 * 1024;
    long long best = 1ll << 60, tot = 0;
    for (int r = 0; r < kRep; ++r) {
      const long long t0 = clock64();
      if (issue == 0 || elect_one()) {
        for (int k = 0; k < kBatch; ++k) {
          // nacc accumulators used round-robin: consecutive MMAs into the SAME accumulator are serialised by the dependency
          const uint32_t acc = tm + (uint32_t)((k % nacc) * (form == 2 ? 64 : 128));
          if (form == 2) tc_mma_bf16_ts(acc, tm + 256 + (k & 7) * 8, umma_desc_sw128(b + (k & 3) * 32), idesc, 1u);
          else if (form == 3) tc_mma_bf16(acc, desc_mn(a + (k & 7) * 2048, 16384), desc_mn(b + (k & 7) * 2048, 16384), idesc, 1u);
          else tc_mma_bf16(acc, umma_desc_sw128(a + (k & 3) * 32), umma_desc_sw128(b + (k & 3) * 32), idesc, 1u);
        }
        tc_commit(&bar);
      }
      if (issue) __syncwarp();
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

int main() {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 161 * 1024 + 1024;
  cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"K-major SS bf16", "K-major SS fp16", "TS (A in TMEM)", "MN-major SS"};
  const int Ns[6] = {32, 48, 64, 128, 208, 256};
  for (int issue = 0; issue < 2; ++issue)
  for (int nacc = 1; nacc <= 4; nacc *= 2)
  for (int form = 0; form < 4; ++form)
    for (int ni = 0; ni < 6; ++ni) {
      const int N = Ns[ni];
      if (form == 2 && N > 64) continue;
      if (form == 3 && N != 64 && N != 128) continue;
      if (nacc > 1 && (issue == 0 || N > 128 || N == 48)) continue;
      long long t[2] = {0, 0};
      cudaError_t e = cudaSuccess;
      for (int bi = 0; bi < 2; ++bi) {     // two batch sizes: the difference removes the commit / mbarrier / wake-up overhead
        cost_kernel<<<148, 128, smem>>>(form, N, bi ? 64 : 16, issue, nacc, d);
        e = cudaDeviceSynchronize();
        long long h[2] = {0, 0};
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        t[bi] = h[0];
        if (e != cudaSuccess) break;
      }
      printf("%-12s acc=%d %-16s M=128 N=%3d K=16: %6.1f cycles per MMA (incremental: (T64 - T16) / 48), fixed overhead per batch %6.0f; T16 %lld T64 %lld; %s\n",
             issue ? "elect.sync" : "if(lane==0)", nacc, names[form], N,
             (double)(t[1] - t[0]) / 48.0, (double)t[0] - 16.0 * (double)(t[1] - t[0]) / 48.0, t[0], t[1], cudaGetErrorString(e));
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
