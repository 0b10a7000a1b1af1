// tcgen05 / TMEM / TMA surrogate predict (sm_100a only): replaces surrogate.predict (_revamp.py:297-322), the Keras MLP
// P -> h1 (relu) -> h2 (relu) -> 1 (sigmoid) of _revamp.py:217-233, for every chain that took the surrogate branch this step.
//
// Layer 1 is a real GEMM — [rows x P] . [P x h1], P up to 3 975, h1 up to 200 — and runs on the tensor cores:
//   pack  : the listed theta' rows -> fp16 image A [n_pad][2 KP] = [x_hi | x_lo]; W1 -> fp16 image B [N1][2 KP] = [w_hi | w_lo] (K-major),
//           x = x_hi + x_lo and w = w_hi + w_lo to 2^-22 (saturating at +-65504; fp16 subnormals keep small values to 2^-25 absolute)
//   GEMM  : D[128 rows, N1] += x_hi.w_hi + x_hi.w_lo + x_lo.w_hi over K blocks of 64 (TMA -> shared memory ring -> tcgen05.mma,
//           fp32 accumulator in TMEM); the dropped term is 2^-22 |x w|, below the fp32 rounding of the Keras-style reference
//   epilogue (thread = row): a1 = relu(D + b1) straight from TMEM, layer 2 (h1 x h2 <= 200 x 50) as FFMA against W2 in shared
//           memory (broadcast LDS.128), relu, layer 3 + sigmoid, s[row].
// One CTA per 128 listed rows: warp 4 = TMA producer, warp 5 = TMEM alloc + MMA issue, warps 0-3 = epilogue (TMEM lane quarters).
#include <cuda.h>
#include <cuda_fp16.h>

#include <map>
#include <mutex>
#include <utility>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

namespace sapt {

int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows);   // sapt_tc_host.cu

namespace st {
constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kStages = 2;
constexpr int kH2Max = 52;      // h2 padded to a multiple of 4

struct Params {
  const int32_t* rows; const int32_t* n_rows; int R;
  int P, h1, h2, N1, KB;        // N1 = h1 padded to 16, KB = K blocks of 64
  const float* b1; const float* W2; const float* b2; const float* W3; const float* b3;
  float* s;
};

// fp16 two-piece split with saturation
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  v = fminf(fmaxf(v, -65504.0f), 65504.0f);
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// A image row i = the i-th listed theta' row (zeros beyond the list): [x_hi (KP) | x_lo (KP)]
__global__ void pack_rows_kernel(const float* __restrict__ theta, int64_t ld, const int32_t* __restrict__ rows, const int32_t* __restrict__ n_rows,
                                 int R, int P, int KP, __half* __restrict__ img) {
  const int i = blockIdx.x;
  const int nact = n_rows ? *n_rows : R;
  if (i >= ((nact + kTileM - 1) / kTileM) * kTileM) return;      // tiles beyond the list are never read
  __half* dst = img + (int64_t)i * 2 * KP;
  const float* src = (i < nact) ? theta + (int64_t)(rows ? rows[i] : i) * ld : nullptr;
  for (int c = threadIdx.x; c < KP; c += blockDim.x) {
    __half hi = __float2half_rn(0.f), lo = hi;
    if (src && c < P) split_h(src[c], hi, lo);
    dst[c] = hi; dst[KP + c] = lo;
  }
}

// B image row n (hidden unit) = column n of W1[P][h1]: [w_hi (KP) | w_lo (KP)]
__global__ void pack_w1_kernel(const float* __restrict__ W1, int P, int h1, int N1, int KP, __half* __restrict__ img) {
  const int n = blockIdx.x;
  __half* dst = img + (int64_t)n * 2 * KP;
  for (int c = threadIdx.x; c < KP; c += blockDim.x) {
    __half hi = __float2half_rn(0.f), lo = hi;
    if (n < h1 && c < P) split_h(W1[(int64_t)c * h1 + n], hi, lo);
    dst[c] = hi; dst[KP + c] = lo;
  }
}

template <int H2Q>   // h2 padded / 4
__global__ void __launch_bounds__(kThreads, 1) surrogate_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                   Params p) {
  constexpr int H2P = 4 * H2Q;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N1 = p.N1, KB = p.KB;
  const uint32_t kABlk = kTileM * 128;                 // one piece of an A stage: 128 rows x 64 fp16
  const uint32_t kBBlk = (uint32_t)N1 * 128;           // one piece of a B stage
  const uint32_t kStage = 2 * kABlk + 2 * kBBlk;       // [A hi | A lo | B hi | B lo]
  uint8_t* sStage = smem;
  float* sW2 = reinterpret_cast<float*>(smem + kStages * ((kStage + 1023u) & ~1023u));   // [h1][H2P]
  float* sB1 = sW2 + p.h1 * H2P;                       // [N1]
  float* sB2 = sB1 + N1;                               // [H2P]
  float* sW3 = sB2 + H2P;                              // [H2P]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sW3 + H2P) + 15) & ~uintptr_t(15));
  uint64_t* full = bars;                  // [kStages]
  uint64_t* empty = full + kStages;       // [kStages]
  uint64_t* acc_full = empty + kStages;   // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  const int row0 = blockIdx.x * kTileM;
  if (row0 >= nact) return;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < N1) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // small operands of layers 2 and 3 (fp32, zero padded)
  for (int i = threadIdx.x; i < p.h1 * H2P; i += kThreads) { const int k = i / H2P, j = i - k * H2P; sW2[i] = (j < p.h2) ? p.W2[k * p.h2 + j] : 0.f; }
  for (int i = threadIdx.x; i < N1; i += kThreads) sB1[i] = (i < p.h1) ? p.b1[i] : 0.f;
  for (int i = threadIdx.x; i < H2P; i += kThreads) { sB2[i] = (i < p.h2) ? p.b2[i] : 0.f; sW3[i] = (i < p.h2) ? p.W3[i] : 0.f; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t stage_stride = (kStage + 1023u) & ~1023u;

  // Warps 4 (TMA) and 5 (MMA) stay converged; the single-thread instructions are issued by an elected lane (see elect_one).
  if (warp == 4) {
    const int KP = KB * 64;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % kStages;
      mbar_wait<200>(&empty[s], ((kb / kStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&full[s], kStage);
        uint8_t* st = sStage + s * stage_stride;
        tma_load_2d(&map_a, &full[s], st, kb * 64, row0);
        tma_load_2d(&map_a, &full[s], st + kABlk, KP + kb * 64, row0);
        tma_load_2d(&map_b, &full[s], st + 2 * kABlk, kb * 64, 0);
        tma_load_2d(&map_b, &full[s], st + 2 * kABlk + kBBlk, KP + kb * 64, 0);
      }
      __syncwarp();
    }
  } else if (warp == 5) {
    // kind::f16, A and B fp16 (format 0), D fp32 (bit 4), N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % kStages;
      mbar_wait(&full[s], (kb / kStages) & 1);
      tc_fence_after();
      const uint32_t a_hi = smem_u32(sStage + s * stage_stride), a_lo = a_hi + kABlk;
      const uint32_t b_hi = a_hi + 2 * kABlk, b_lo = b_hi + kBBlk;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {     // four K steps of 16 inside the 64-element block
          tc_mma_bf16(tmem_base, umma_desc_sw128(a_hi + k * 32), umma_desc_sw128(b_hi + k * 32), idesc, (kb | k) ? 1u : 0u);
          tc_mma_bf16(tmem_base, umma_desc_sw128(a_hi + k * 32), umma_desc_sw128(b_lo + k * 32), idesc, 1u);
          tc_mma_bf16(tmem_base, umma_desc_sw128(a_lo + k * 32), umma_desc_sw128(b_hi + k * 32), idesc, 1u);
        }
        tc_commit(&empty[s]);
        if (kb == KB - 1) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ================= epilogue: thread = listed row =================
    const int r = row0 + warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float acc2[H2P];
#pragma unroll
    for (int j = 0; j < H2P; ++j) acc2[j] = sB2[j];
    mbar_wait(acc_full, 0);
    tc_fence_after();
    for (int k0 = 0; k0 < N1; k0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + lane_base + (uint32_t)k0, v);
      tmem_ld_wait();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        const int k = k0 + kk;
        if (k < p.h1) {                                             // warp-uniform
          const float a1 = fmaxf(__uint_as_float(v[kk]) + sB1[k], 0.f);   // relu (RV:221-229)
          const float4* w = reinterpret_cast<const float4*>(sW2 + k * H2P);
#pragma unroll
          for (int q = 0; q < H2Q; ++q) {
            const float4 ww = w[q];
            acc2[4 * q] = fmaf(a1, ww.x, acc2[4 * q]); acc2[4 * q + 1] = fmaf(a1, ww.y, acc2[4 * q + 1]);
            acc2[4 * q + 2] = fmaf(a1, ww.z, acc2[4 * q + 2]); acc2[4 * q + 3] = fmaf(a1, ww.w, acc2[4 * q + 3]);
          }
        }
      }
    }
    float z = p.b3[0];
#pragma unroll
    for (int j = 0; j < H2P; ++j) z = fmaf(fmaxf(acc2[j], 0.f), sW3[j], z);
    if (r < nact) p.s[p.rows ? p.rows[r] : r] = 1.0f / (1.0f + expf(-z));      // sigmoid output (RV:230)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols));
  }
}

struct Workspace { void* buf = nullptr; size_t bytes = 0; };

static int get_workspace(size_t bytes, cudaStream_t stream, void** out) {
  static std::mutex mu;
  static std::map<std::pair<int, void*>, Workspace> table;   // per (device, stream), like the exchange sweep's scratch
  int dev = 0;
  SAPT_CUDA(cudaGetDevice(&dev));
  std::unique_lock<std::mutex> lock(mu);
  Workspace& w = table[std::make_pair(dev, (void*)stream)];
  lock.unlock();
  if (bytes > w.bytes) {      // growing synchronises this stream: happens on the first call of a given size, not capturable
    SAPT_CUDA(cudaStreamSynchronize(stream));
    if (w.buf) SAPT_CUDA(cudaFree(w.buf));
    w.buf = nullptr; w.bytes = 0;
    SAPT_CUDA(cudaMalloc(&w.buf, bytes + bytes / 4));
    w.bytes = bytes + bytes / 4;
  }
  *out = w.buf;
  return SAPT_OK;
}

}  // namespace st

bool surrogate_tc_supported(int P, int h1, int h2) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  return major == 10 && h1 >= 1 && h1 <= 256 && h2 >= 1 && h2 <= st::kH2Max && P >= 1;
}

int launch_surrogate_tc(const float* theta, int64_t ld, int R, int P, const int32_t* rows, const int32_t* n_rows, const float* W1, const float* b1,
                        int h1, const float* W2, const float* b2, int h2, const float* W3, const float* b3, float* s, cudaStream_t stream) {
  using namespace st;
  const int N1 = ((h1 + 15) / 16) * 16;
  const int KB = (P + 63) / 64, KP = KB * 64;
  const int n_pad = ((R + kTileM - 1) / kTileM) * kTileM;
  const int H2Q = (h2 + 3) / 4, H2P = 4 * H2Q;
  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t a_bytes = (size_t)n_pad * 2 * KP * 2, b_bytes = (size_t)N1 * 2 * KP * 2;
  void* ws = nullptr;
  int rc = get_workspace(up(a_bytes) + up(b_bytes), stream, &ws);
  if (rc) return rc;
  __half* aimg = reinterpret_cast<__half*>(ws);
  __half* bimg = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(ws) + up(a_bytes));
  pack_rows_kernel<<<n_pad, 256, 0, stream>>>(theta, ld, rows, n_rows, R, P, KP, aimg);
  SAPT_CHECK_LAUNCH();
  pack_w1_kernel<<<N1, 256, 0, stream>>>(W1, P, h1, N1, KP, bimg);
  SAPT_CHECK_LAUNCH();
  CUtensorMap ma, mb;
  if ((rc = tc_make_map_2d(&ma, aimg, (uint64_t)2 * KP, (uint64_t)n_pad, kTileM))) return rc;
  if ((rc = tc_make_map_2d(&mb, bimg, (uint64_t)2 * KP, (uint64_t)N1, (uint32_t)N1))) return rc;
  Params a{};
  a.rows = rows; a.n_rows = n_rows; a.R = R; a.P = P; a.h1 = h1; a.h2 = h2; a.N1 = N1; a.KB = KB;
  a.b1 = b1; a.W2 = W2; a.b2 = b2; a.W3 = W3; a.b3 = b3; a.s = s;
  const uint32_t stage = 2 * kTileM * 128 + 2 * (uint32_t)N1 * 128;
  const size_t smem = 1024 + kStages * (size_t)((stage + 1023u) & ~1023u) + (size_t)(h1 * H2P + N1 + 2 * H2P) * sizeof(float) + 256;
  if (smem > 227 * 1024) { set_error("sapt_surrogate_predict (tcgen05): h1=%d h2=%d needs %zu bytes of shared memory", h1, h2, smem); return SAPT_ERR_UNSUPPORTED; }
  const dim3 grid(n_pad / kTileM);
#define SAPT_ST(Q)                                                                                                           \
  if (H2Q == Q) {                                                                                                            \
    auto k = surrogate_tc_kernel<Q>;                                                                                         \
    SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                              \
    k<<<grid, kThreads, smem, stream>>>(ma, mb, a);                                                                          \
  } else
  SAPT_ST(1) SAPT_ST(2) SAPT_ST(3) SAPT_ST(4) SAPT_ST(5) SAPT_ST(6) SAPT_ST(7) SAPT_ST(8) SAPT_ST(9) SAPT_ST(10) SAPT_ST(11) SAPT_ST(12) SAPT_ST(13)
  { set_error("sapt_surrogate_predict (tcgen05): h2=%d outside the compiled range", h2); return SAPT_ERR_UNSUPPORTED; }
#undef SAPT_ST
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

}  // namespace sapt
