// Development probe (run on the GPU box): checks the MN-major shared-memory operand descriptors used by the
// batch-Langevin contraction G[M, N] = A^T B with K = data rows: A stored [K rows][M], B stored [K rows][N], both
// 128 B per row segment, SWIZZLE_128B written by plain st.shared.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

#include "../sapt_b200/csrc/sapt_tc_ptx.cuh"
using namespace sapt;

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 64-element blocks along M / N
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride between groups of 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int K = 128, M = 128, N = 128;
constexpr int kBlk = K * 128;   // bytes per 64-column block

__global__ void __launch_bounds__(128) mn_kernel(const float* a, const float* b, float* out, int variant) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;                 // M/64 blocks
  uint8_t* sB = sm + (M / 64) * kBlk;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // thread = K row: writes its row of A and B, swizzled
  {
    const int r = tid;
    for (int m = 0; m < M; ++m) {
      const int blk = m / 64, c = (m % 64) / 8, e = m % 8;
      uint8_t* p = sA + blk * kBlk + r * 128 + ((c ^ (r & 7)) << 4) + e * 2;
      *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(a[r * M + m]);
    }
    for (int n = 0; n < N; ++n) {
      const int blk = n / 64, c = (n % 64) / 8, e = n % 8;
      uint8_t* p = sB + blk * kBlk + r * 128 + ((c ^ (r & 7)) << 4) + e * 2;
      *reinterpret_cast<__nv_bfloat16*>(p) = __float2bfloat16_rn(b[r * N + n]);
    }
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid == 0) {
    // bf16 in, f32 acc, a_major = b_major = MN (bits 15, 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t off = ks * 2048;
      tc_mma_bf16(tm, umma_desc_mn_sw128(smem_u32(sA) + off, kBlk), umma_desc_mn_sw128(smem_u32(sB) + off, kBlk), idesc, ks ? 1u : 0u);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(128));
}

int main() {
  std::vector<float> a(K * M), b(K * N), ref(M * N, 0.f), out(M * N);
  for (int r = 0; r < K; ++r) for (int m = 0; m < M; ++m) a[r * M + m] = (float)(((r * 3 + m * 5) % 7) - 3);
  for (int r = 0; r < K; ++r) for (int n = 0; n < N; ++n) b[r * N + n] = (float)(((r * 5 + n * 3) % 5) - 2);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int r = 0; r < K; ++r) s += a[r * M + m] * b[r * N + n]; ref[m * N + n] = s; }
  float *da, *db, *dout;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, out.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (M / 64 + N / 64) * kBlk + 2048;
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mn_kernel<<<1, 128, smem>>>(da, db, dout, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0; float maxd = 0;
  for (int i = 0; i < M * N; ++i) { float d = fabsf(out[i] - ref[i]); if (d > 1e-3f) ++bad; if (d > maxd) maxd = d; }
  printf("mismatches %d of %d, max diff %g; out[0..3] %g %g %g %g ref %g %g %g %g; out[64*N+65] %g ref %g\n", bad, M * N, maxd, out[0], out[1], out[2], out[3],
         ref[0], ref[1], ref[2], ref[3], out[64 * N + 65], ref[64 * N + 65]);
  return bad != 0;
}
