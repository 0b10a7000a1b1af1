// tcgen05 likelihood, version 2: BOTH layers on the tensor cores (sm_100a only).
//
// v1 (sapt_loglik_tc.cu) ran layer 2 (H x O per chain) as FFMA with weights fetched by warp-uniform
// LDS.128; ncu showed the shared-memory pipe at 93 % (a uniform LDS.128 still costs 4 wavefronts), i.e.
// the kernel was bound by weight fetch, not math. v2 follows the flash-attention-on-Blackwell pattern:
//
//   MMA-1  D1[128 rows, 4 chains x HP]  = A[128, 112] . B1            (smem x smem -> TMEM, as v1)
//   epilogue phase 1 (thread = row): tcgen05.ld D1 -> h = sigmoid -> split h = h1 + h2 (bf16) ->
//            tcgen05.st  HV[chain] = [h1 | h1 | h2 | h1]  (bf16 pairs, the A operand of MMA-2, in TMEM)
//   MMA-2  D2[128 rows, ON] = HV[chain] . [w1 ; w2 ; w1 ; w3]          (TMEM x smem -> TMEM, per chain),
//            w = -log2(e) W2 split into three bf16 pieces (exact); D2 overwrites the chain's dead D1 columns
//   epilogue phase 2: tcgen05.ld D2 (+ bias) -> sigmoid -> softmax / argmax -> fp64 log-likelihood sums
//
// Shared memory now only feeds the tensor cores (TMA-written, UMMA-read); the SIMT side touches it for
// 16 bias floats per evaluation. Dropped cross terms: h2.w2, h3.* (|.| <= 2^-17 per activation, row
// dependent rounding residue, i.e. unbiased noise ~1e-5 in z2) — parity tests hold the 1e-4 bound with
// 3 orders of magnitude to spare; the weights themselves are represented exactly.
//
// TMEM map (512 columns): [0,256) two D1/D2 stages of 128 | [256,512) four HV buffers of 64
// (HP=16: stages of 64 at 0/64, HV buffers of 32 from 256).
#include <algorithm>
#include <vector>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

namespace sapt {

namespace v2 {
constexpr int kTileM = 128;
constexpr int kKA = 128;        // bf16 per A/B1 row
constexpr int kK1 = 112;        // used K of MMA-1
constexpr int kUC = 4;          // chains per unit (one MMA-1)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr uint32_t kABytesPerKb = kTileM * 128;
constexpr uint32_t kAStageBytes = 2 * kABytesPerKb;
constexpr uint32_t kB1BytesPerKb = 256 * 128;
constexpr uint32_t kB1Bytes = 2 * kB1BytesPerKb;
}  // namespace v2

struct Tc2Params {
  const int32_t* y; int64_t n_pad;
  const float* bias2;            // [pos][ON] log2(e) * B2
  const int32_t* n_rows; int R;
  int O; int tiles_per_chunk; int64_t n_tiles;
  double* partL; int32_t* partNc; int32_t* partSq;
};

// ---- per-step operand packing: B1 rows, W2 rows (bf16 pieces), bias2 (fp32) ---------------------
// B1: [group][CH*HP][128]; W2: [group][CH*ON][K2] with K2 = 4*HP = [w1 | w2 | w1 | w3] per output row
__global__ void tc2_pack_kernel(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O,
                                int HP, int ON, int CH, __nv_bfloat16* b1img, __nv_bfloat16* w2img, float* bias2) {
  const int nact = n_rows ? *n_rows : R;
  const int pos = blockIdx.x;
  const int n_pos_pad = ((nact + CH - 1) / CH) * CH;
  if (pos >= n_pos_pad) return;
  const int grp = pos / CH, within = pos - grp * CH;
  const int K2 = 4 * HP;
  __nv_bfloat16* brow0 = b1img + ((int64_t)grp * 256 + (int64_t)within * HP) * v2::kKA;
  __nv_bfloat16* wrow0 = w2img + ((int64_t)grp * CH * ON + (int64_t)within * ON) * K2;
  float* bb = bias2 + (int64_t)pos * ON;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (pos >= nact) {
    for (int i = threadIdx.x; i < HP * v2::kKA; i += blockDim.x) brow0[i] = zero;
    for (int i = threadIdx.x; i < ON * K2; i += blockDim.x) wrow0[i] = zero;
    for (int i = threadIdx.x; i < ON; i += blockDim.x) bb[i] = 0.f;
    return;
  }
  const int row = rows ? rows[pos] : pos;
  const float* src = theta + (int64_t)row * ld;
  for (int i = threadIdx.x; i < HP * 8; i += blockDim.x) {
    const int h = i >> 3, c = i & 7;
    __nv_bfloat16* dst = brow0 + (int64_t)h * v2::kKA + c * 16;
    if (h >= H || c == 7) {
#pragma unroll
      for (int d = 0; d < 16; ++d) dst[d] = zero;
    } else if (c == 6) {
      __nv_bfloat16 p1, p2, p3;
      split3(kLog2e * src[D * H + H * O + h], p1, p2, p3);
      dst[0] = p1; dst[1] = p2; dst[2] = p3;
#pragma unroll
      for (int d = 3; d < 16; ++d) dst[d] = zero;
    } else {
      const int piece = (c == 0 || c == 2 || c == 5) ? 0 : ((c == 1 || c == 4) ? 1 : 2);
      for (int d = 0; d < 16; ++d) {
        __nv_bfloat16 p[3] = {zero, zero, zero};
        if (d < D) split3(-kLog2e * src[d * H + h], p[0], p[1], p[2]);
        dst[d] = p[piece];
      }
    }
  }
  for (int i = threadIdx.x; i < ON * HP; i += blockDim.x) {
    const int o = i / HP, h = i - o * HP;
    __nv_bfloat16 p[3] = {zero, zero, zero};
    if (o < O && h < H) split3(-kLog2e * src[D * H + h * O + o], p[0], p[1], p[2]);
    __nv_bfloat16* dst = wrow0 + (int64_t)o * K2;
    dst[h] = p[0]; dst[HP + h] = p[1]; dst[2 * HP + h] = p[0]; dst[3 * HP + h] = p[2];
  }
  for (int i = threadIdx.x; i < ON; i += blockDim.x) bb[i] = (i < O) ? kLog2e * src[D * H + H * O + H + i] : 0.f;
}

template <int HP, int ON, int STAGES>
__global__ void __launch_bounds__(v2::kThreads, 1) loglik_tc2_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                     const __grid_constant__ CUtensorMap map_b1,
                                                                     const __grid_constant__ CUtensorMap map_w2, Tc2Params p) {
  using namespace v2;
  constexpr int CH = 256 / HP;              // chains per CTA
  constexpr int UPT = CH / kUC;             // units (MMA-1 issues) per row tile
  constexpr int N1 = kUC * HP;              // MMA-1 N
  constexpr int K2 = 4 * HP;                // MMA-2 K
  constexpr int KB2 = K2 / 64;              // K blocks of the W2 operand
  constexpr int HVC = K2 / 2;               // TMEM columns of one HV buffer
  constexpr uint32_t kW2BytesPerKb = CH * ON * 128;
  constexpr uint32_t kW2Bytes = KB2 * kW2BytesPerKb;
  static_assert(ON <= HP, "D2 must fit in the chain's D1 columns");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB1 = smem;
  uint8_t* sA = sB1 + kB1Bytes;
  uint8_t* sW2 = sA + STAGES * kAStageBytes;
  float* sBias = reinterpret_cast<float*>(sW2 + kW2Bytes);          // [CH][ON]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + CH * ON);
  uint64_t* full_a = bars;                  // [STAGES]
  uint64_t* empty_a = full_a + STAGES;      // [STAGES]
  uint64_t* b_full = empty_a + STAGES;      // [1]
  uint64_t* d1_full = b_full + 1;           // [2]
  uint64_t* d1_empty = d1_full + 2;         // [2]
  uint64_t* hv_full = d1_empty + 2;         // [kUC]
  uint64_t* d2_full = hv_full + kUC;        // [kUC]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d2_full + kUC);
  double* red = reinterpret_cast<double*>(tmem_ptr + 2);  // [CH][4 quarters][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  const int g0 = blockIdx.y * CH;
  if (g0 >= nact) return;
  const int64_t t_begin = (int64_t)blockIdx.x * p.tiles_per_chunk;
  const int64_t t_end = min(t_begin + (int64_t)p.tiles_per_chunk, p.n_tiles);
  const int n_my_tiles = (int)max((int64_t)0, t_end - t_begin);
  const int n_units = n_my_tiles * UPT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&d1_full[s], 1); mbar_init(&d1_empty[s], kEpiWarps); }
    for (int c = 0; c < kUC; ++c) { mbar_init(&hv_full[c], 4); mbar_init(&d2_full[c], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    const float* src = p.bias2 + (int64_t)g0 * ON;
    const int ng = min(CH, nact - g0);
    for (int i = threadIdx.x; i < CH * ON; i += kThreads) sBias[i] = (i < ng * ON) ? src[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_hv = tmem_base + 256;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(b_full, kB1Bytes + kW2Bytes);
      tma_load_2d(&map_b1, b_full, sB1, 0, blockIdx.y * 256);
      tma_load_2d(&map_b1, b_full, sB1 + kB1BytesPerKb, 64, blockIdx.y * 256);
#pragma unroll
      for (int kb = 0; kb < KB2; ++kb) tma_load_2d(&map_w2, b_full, sW2 + kb * kW2BytesPerKb, kb * 64, blockIdx.y * CH * ON);
      for (int it = 0; it < n_my_tiles; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_a[s], ph ^ 1);
        mbar_expect_tx(&full_a[s], kAStageBytes);
        const int row0 = (int)((t_begin + it) * kTileM);
        tma_load_2d(&map_a, &full_a[s], sA + s * kAStageBytes, 0, row0);
        tma_load_2d(&map_a, &full_a[s], sA + s * kAStageBytes + kABytesPerKb, 64, row0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ON >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t b1_addr = smem_u32(sB1);
      const uint32_t w2_addr = smem_u32(sW2);
      mbar_wait(b_full, 0);
      auto issue_mma1 = [&](int u) {
        const int it = u / UPT, uu = u - it * UPT;
        const int s = it % STAGES;
        const int ds = u & 1;
        mbar_wait(&d1_empty[ds], ((u >> 1) & 1) ^ 1);
        if (uu == 0) mbar_wait(&full_a[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * kAStageBytes);
        const uint32_t d_tmem = tmem_base + (uint32_t)(ds * N1);
#pragma unroll
        for (int k = 0; k < kK1 / 16; ++k) {
          const uint32_t kb = k >> 2, kin = k & 3;
          const uint64_t ad = umma_desc_sw128(a_addr + kb * kABytesPerKb + kin * 32);
          const uint64_t bd = umma_desc_sw128(b1_addr + kb * kB1BytesPerKb + (uint32_t)(uu * N1) * 128 + kin * 32);
          tc_mma_bf16(d_tmem, ad, bd, idesc1, k > 0 ? 1u : 0u);
        }
        if (uu == UPT - 1) tc_commit(&empty_a[s]);
        tc_commit(&d1_full[ds]);
      };
      if (n_units > 0) issue_mma1(0);
      for (int u = 0; u < n_units; ++u) {
        const int uu = u % UPT;
        const int ds = u & 1;
        const uint32_t uph = u & 1;
#pragma unroll
        for (int ci = 0; ci < kUC; ++ci) {
          // chains are finished by the epilogue warps in the order 0,2 (first chain of each warp) then 1,3
          const int c = (ci < 2) ? (2 * ci) : (2 * (ci - 2) + 1);
          if (ci == 2 && u + 1 < n_units) issue_mma1(u + 1);
          mbar_wait(&hv_full[c], uph);
          tc_fence_after();
          const uint32_t d2 = tmem_base + (uint32_t)(ds * N1 + c * HP);
          const uint32_t hv = tmem_hv + (uint32_t)(c * HVC);
          const int chain = uu * kUC + c;
#pragma unroll
          for (int j = 0; j < K2 / 16; ++j) {
            const uint32_t kb = j >> 2, kin = j & 3;
            const uint64_t bd = umma_desc_sw128(w2_addr + kb * kW2BytesPerKb + (uint32_t)(chain * ON) * 128 + kin * 32);
            tc_mma_bf16_ts(d2, hv + (uint32_t)(j * 8), bd, idesc2, j > 0 ? 1u : 0u);
          }
          tc_commit(&d2_full[c]);
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int e = ew >> 2;                    // this warp owns chains 2e, 2e+1 of every unit
    const int O = p.O;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    double accL[UPT][2];
    int accNc[UPT][2], accSq[UPT][2];
#pragma unroll
    for (int a = 0; a < UPT; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) { accL[a][b] = 0.0; accNc[a][b] = 0; accSq[a][b] = 0; }
    for (int it = 0; it < n_my_tiles; ++it) {
      const int64_t row = (t_begin + it) * kTileM + quarter * 32 + lane;
      const int y = __ldg(p.y + row);
#pragma unroll
      for (int uu = 0; uu < UPT; ++uu) {
        const int u = it * UPT + uu;
        const int ds = u & 1;
        mbar_wait(&d1_full[ds], (u >> 1) & 1);
        tc_fence_after();
        // ---- phase 1: hidden activations -> bf16 pieces -> TMEM (A operand of MMA-2) ----
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * e + cc;
          uint32_t zr[HP];
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(ds * N1 + c * HP);
          if constexpr (HP == 32) tmem_ld32(taddr, zr); else tmem_ld16(taddr, zr);
          tmem_ld_wait();
          uint32_t h1p[HP / 2], h2p[HP / 2];
#pragma unroll
          for (int j = 0; j < HP / 2; ++j) {
            const float ha = sigmoid_from_neg_log2e(__uint_as_float(zr[2 * j]));
            const float hb = sigmoid_from_neg_log2e(__uint_as_float(zr[2 * j + 1]));
            const uint32_t hi = pack_bf16x2(ha, hb);
            const float ra = ha - __uint_as_float(hi << 16);
            const float rb = hb - __uint_as_float(hi & 0xffff0000u);
            h1p[j] = hi;
            h2p[j] = pack_bf16x2(ra, rb);
          }
          // HV[c] = [h1 | h1 | h2 | h1], HP/2 packed columns per chunk
          const uint32_t hv = tmem_hv + lane_base + (uint32_t)(c * HVC);
          if constexpr (HP == 32) {
            uint32_t w[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) { w[j] = h1p[j]; w[16 + j] = h1p[j]; }
            tmem_st32(hv, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) { w[j] = h2p[j]; w[16 + j] = h1p[j]; }
            tmem_st32(hv + 32, w);
          } else {
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { w[j] = h1p[j]; w[8 + j] = h1p[j]; }
            tmem_st16(hv, w);
#pragma unroll
            for (int j = 0; j < 8; ++j) { w[j] = h2p[j]; w[8 + j] = h1p[j]; }
            tmem_st16(hv + 16, w);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&hv_full[c]);
        }
        // ---- phase 2: output layer from D2 ----
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * e + cc;
          mbar_wait(&d2_full[c], u & 1);
          tc_fence_after();
          uint32_t dr[ON];
          const uint32_t taddr = tmem_base + lane_base + (uint32_t)(ds * N1 + c * HP);
          if constexpr (ON == 32) tmem_ld32(taddr, dr); else tmem_ld16(taddr, dr);
          tmem_ld_wait();
          if (cc == 1) {  // both of this warp's D2 blocks are in registers: the stage may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[ds]);
          }
          const float* bias = sBias + (uu * kUC + c) * ON;
          float tmin = 0.f; int pm = 0; float sum = 0.f, oy = 0.f;
#pragma unroll
          for (int o = 0; o < ON; ++o) {
            if (o < O) {
              const float t = __uint_as_float(dr[o]) + bias[o];
              if (o == 0 || t < tmin) { tmin = t; pm = o; }
              const float s = sigmoid_from_neg_log2e(t);
              sum += ex2_approx(s * kLog2e);
              if (o == y) oy = s;
            }
          }
          if (y >= 0) {
            accL[uu][cc] += (double)(oy - kLn2 * lg2_approx(sum));
            accNc[uu][cc] += (pm == y);
            const int d = pm - y;
            accSq[uu][cc] += d * d;
          }
        }
      }
    }
#pragma unroll
    for (int uu = 0; uu < UPT; ++uu)
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int chain = uu * kUC + 2 * e + cc;
        const double l = warp_sum(accL[uu][cc]);
        const int a = warp_sum(accNc[uu][cc]);
        const int b = warp_sum(accSq[uu][cc]);
        if (lane == 0) {
          double* r = red + (chain * 4 + quarter) * 3;
          r[0] = l; r[1] = (double)a; r[2] = (double)b;
        }
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
  if (threadIdx.x < CH && g0 + (int)threadIdx.x < nact) {
    const double* r = red + threadIdx.x * 12;
    const int64_t o = (int64_t)blockIdx.x * p.R + (g0 + threadIdx.x);
    p.partL[o] = (r[0] + r[3]) + (r[6] + r[9]);
    p.partNc[o] = (int32_t)(r[1] + r[4] + r[7] + r[10]);
    p.partSq[o] = (int32_t)(r[2] + r[5] + r[8] + r[11]);
  }
}

// ---- host ------------------------------------------------------------------------------------------
int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows);  // sapt_loglik_tc.cu

template <int HP, int ON, int STAGES>
static int launch_tc2(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mw2, const Tc2Params& a, dim3 grid, cudaStream_t st) {
  constexpr int CH = 256 / HP;
  constexpr int K2 = 4 * HP;
  const size_t smem = 1024 + v2::kB1Bytes + (size_t)STAGES * v2::kAStageBytes + (size_t)(K2 / 64) * CH * ON * 128 + (size_t)CH * ON * 4 +
                      (2 * STAGES + 1 + 4 + 2 * v2::kUC) * 8 + 16 + (size_t)CH * 12 * 8 + 64;
  auto k = loglik_tc2_kernel<HP, ON, STAGES>;
  SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, v2::kThreads, smem, st>>>(ma, mb1, mw2, a);
  SAPT_CHECK_LAUNCH();
  return SAPT_OK;
}

static int gcd_int2(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

int launch_loglik_tc2(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                      double* L, int32_t* nc, double* sq, cudaStream_t st) {
  const DataSet& ds = p->set[which];
  if (R <= 0 || ds.n <= 0) return SAPT_OK;
  const int HP = p->H <= 16 ? 16 : 32;
  const int ON = (p->O <= 16) ? 16 : 32;
  if (ON > HP) { set_error("tcgen05 v2: O=%d needs H > 16", p->O); return SAPT_ERR_UNSUPPORTED; }
  const int CH = 256 / HP;
  const int K2 = 4 * HP;
  const int n_groups = (R + CH - 1) / CH;
  const int64_t n_tiles = ds.n_pad / v2::kTileM;
  const int base = p->num_sms / gcd_int2(n_groups, p->num_sms);
  int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_tiles / 16, ((int64_t)p->num_sms * 8 + n_groups - 1) / n_groups));
  if (n_chunks >= base) n_chunks = (n_chunks / base) * base;
  n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, n_tiles));
  const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  const size_t b1_bytes = (size_t)n_groups * 256 * v2::kKA * 2;
  const size_t w2_bytes = (size_t)n_groups * CH * ON * K2 * 2;
  const size_t bias_bytes = (size_t)n_groups * CH * ON * sizeof(float);
  const size_t part_bytes = (size_t)n_chunks * R * (sizeof(double) + 2 * sizeof(int32_t));
  int rc = ensure_workspace(p, up(b1_bytes) + up(w2_bytes) + up(bias_bytes) + up(part_bytes));
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->ws);
  __nv_bfloat16* b1img = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* w2img = reinterpret_cast<__nv_bfloat16*>(ws + up(b1_bytes));
  float* bias2 = reinterpret_cast<float*>(ws + up(b1_bytes) + up(w2_bytes));
  double* partL = reinterpret_cast<double*>(ws + up(b1_bytes) + up(w2_bytes) + up(bias_bytes));
  int32_t* partNc = reinterpret_cast<int32_t*>(partL + (size_t)n_chunks * R);
  int32_t* partSq = partNc + (size_t)n_chunks * R;

  tc2_pack_kernel<<<n_groups * CH, 128, 0, st>>>(theta, ld, rows, n_rows, R, p->D, p->H, p->O, HP, ON, CH, b1img, w2img, bias2);
  SAPT_CHECK_LAUNCH();
  CUtensorMap mb1, mw2;
  rc = tc_make_map_2d(&mb1, b1img, v2::kKA, (uint64_t)n_groups * 256, 256);
  if (rc) return rc;
  rc = tc_make_map_2d(&mw2, w2img, (uint64_t)K2, (uint64_t)n_groups * CH * ON, (uint32_t)(CH * ON));
  if (rc) return rc;
  Tc2Params a{};
  a.y = ds.y; a.n_pad = ds.n_pad; a.bias2 = bias2; a.n_rows = n_rows; a.R = R; a.O = p->O;
  a.tiles_per_chunk = tiles_per_chunk; a.n_tiles = n_tiles; a.partL = partL; a.partNc = partNc; a.partSq = partSq;
  dim3 grid((unsigned)n_chunks, (unsigned)n_groups);
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(ds.tmap);
  if (HP == 16) rc = launch_tc2<16, 16, 3>(ma, mb1, mw2, a, grid, st);
  else if (ON == 16) rc = launch_tc2<32, 16, 3>(ma, mb1, mw2, a, grid, st);
  else rc = launch_tc2<32, 32, 2>(ma, mb1, mw2, a, grid, st);
  if (rc) return rc;
  return launch_loglik_finalize(partL, partNc, partSq, (int)n_chunks, R, rows, n_rows, L, nc, sq, st);
}

}  // namespace sapt
