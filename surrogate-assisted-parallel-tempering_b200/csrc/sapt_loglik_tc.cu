// tcgen05 / TMEM / TMA path of the batched likelihood (sm_100a only).
//
// Layer 1 of every chain is ONE stacked GEMM per row tile:  Z[128 rows, NG chains x HP hidden]
// = A[128, K] . B[K, NG*HP], issued as tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in
// TMEM). fp32-grade accuracy is kept by splitting every fp32 value into three bf16 pieces
// (v = v1 + v2 + v3 exactly) and stacking the six significant cross terms along K:
//     A row  = [ x1 | x1 | x2 | x1 | x2 | x3 | 1 1 1 0.. ]          (7 chunks of 16 = K 112, padded to 128)
//     B col  = [ w1 ; w2 ; w1 ; w3 ; w2 ; w1 ; b1 b2 b3 0.. ]       w = -log2(e) W1[:,h],  b = +log2(e) B1[h]
// so the accumulator is t = -log2(e) (x.W1 - B1) and a sigmoid is ex2 + add + rcp. Dropped terms are
// O(2^-24); the tensor pipe has >5x headroom over the SFU-bound epilogue, so the 7 K-steps are free.
// The A image is built once per data set (operand of every step), B once per step from theta'.
//
// CTA = (row chunk, group of NG chains). Warp 0 lane 0: TMA producer (A tiles, 3-stage mbarrier ring;
// B panel once). Warp 1: TMEM alloc + single-thread MMA issue into a 2-stage TMEM accumulator ring.
// Warps 2..: epilogue straight out of TMEM (tcgen05.ld 32x32b, thread = data row): sigmoid ->
// block-diagonal H x O layer from shared memory (broadcast LDS.128) -> sigmoid -> softmax / argmax ->
// fp64 per-thread log-likelihood sums. Hidden activations never leave the SM.
#include <cuda.h>
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "sapt_device.cuh"
#include "sapt_internal.h"
#include "sapt_tc_ptx.cuh"

namespace sapt {

constexpr int kTileM = 128;
constexpr int kKA = 128;           // bf16 per A/B row (two 128-byte swizzle rows)
constexpr int kKUsed = 112;        // 6 x 16 cross-term chunks + one bias chunk
constexpr int kUmmaK = 16;
constexpr int kStages = 3;
constexpr int kAcc = 2;
constexpr int kUmmaN = 256;
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 32 * (2 + kEpiWarps);
constexpr uint32_t kABytesPerKb = kTileM * 128;      // 16 KB: one K block of an A tile
constexpr uint32_t kAStageBytes = 2 * kABytesPerKb;  // 32 KB
constexpr uint32_t kBBytesPerKb = kUmmaN * 128;      // 32 KB
constexpr uint32_t kBBytes = 2 * kBBytesPerKb;       // 64 KB

bool tc_topology_supported(int D, int H, int O) { return D <= 16 && H <= 32 && O <= 20; }

// ---- per-step operand packing -----------------------------------------------------------------
// B image rows: [group][NG*HP][128] bf16; W2 image: [pos][HP*OP + OP] fp32 (pre-scaled by -/+ log2 e)
__global__ void tc_pack_kernel(const float* theta, int64_t ld, const int32_t* rows, const int32_t* n_rows, int R, int D, int H, int O,
                               int HP, int OP, int NG, __nv_bfloat16* bimg, float* w2img) {
  const int nact = n_rows ? *n_rows : R;
  const int pos = blockIdx.x;
  const int n_pos_pad = ((nact + NG - 1) / NG) * NG;
  if (pos >= n_pos_pad) return;
  const int grp = pos / NG, within = pos - grp * NG;
  __nv_bfloat16* brow0 = bimg + ((int64_t)grp * kUmmaN + (int64_t)within * HP) * kKA;
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  if (pos >= nact) {
    for (int i = threadIdx.x; i < HP * kKA; i += blockDim.x) brow0[i] = zero;
    return;
  }
  const int row = rows ? rows[pos] : pos;
  const float* src = theta + (int64_t)row * ld;
  for (int i = threadIdx.x; i < HP * 8; i += blockDim.x) {  // (h, chunk) pairs; chunk 7 is the zero pad
    const int h = i >> 3, c = i & 7;
    __nv_bfloat16* dst = brow0 + (int64_t)h * kKA + c * 16;
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
      // chunk -> which piece of w pairs with the A chunk [x1|x1|x2|x1|x2|x3]
      const int piece = (c == 0 || c == 2 || c == 5) ? 0 : ((c == 1 || c == 4) ? 1 : 2);
      for (int d = 0; d < 16; ++d) {
        __nv_bfloat16 p[3] = {zero, zero, zero};
        if (d < D) split3(-kLog2e * src[d * H + h], p[0], p[1], p[2]);
        dst[d] = p[piece];
      }
    }
  }
  float* w2 = w2img + (int64_t)pos * (HP * OP + OP);
  for (int i = threadIdx.x; i < HP * OP + OP; i += blockDim.x) {
    float v = 0.f;
    if (i < HP * OP) {
      const int h = i / OP, o = i - h * OP;
      if (h < H && o < O) v = -kLog2e * src[D * H + h * O + o];
    } else {
      const int o = i - HP * OP;
      if (o < O) v = kLog2e * src[D * H + H * O + H + o];
    }
    w2[i] = v;
  }
}

struct TcParams {
  const int32_t* y; int64_t n_pad;
  const float* w2img;
  const int32_t* n_rows; int R;
  int O; int tiles_per_chunk; int64_t n_tiles;
  double* partL; int32_t* partNc; int32_t* partSq;
};

template <int HP, int OP>
__global__ void __launch_bounds__(kTcThreads, 1) loglik_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b, TcParams p) {
  constexpr int NG = kUmmaN / HP;                       // chains per CTA
  constexpr int CPW = NG / (kEpiWarps / 4);             // chains per epilogue warp
  constexpr int W2S = HP * OP + OP;                     // floats per chain in the W2 image
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B 64K][A stages 3 x 32K][W2 image][barriers][tmem ptr][reduction scratch]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = sB + kBBytes;
  float* sW2 = reinterpret_cast<float*>(sA + kStages * kAStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW2 + NG * W2S + ((NG * W2S) & 1));
  uint64_t* full_a = bars;               // [kStages]
  uint64_t* empty_a = bars + kStages;    // [kStages]
  uint64_t* b_full = bars + 2 * kStages; // [1]
  uint64_t* tm_full = b_full + 1;        // [kAcc]
  uint64_t* tm_empty = tm_full + kAcc;   // [kAcc]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tm_empty + kAcc);
  double* red = reinterpret_cast<double*>(tmem_ptr + 2);  // [NG][4 quarters][3]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nact = p.n_rows ? *p.n_rows : p.R;
  const int g0 = blockIdx.y * NG;
  if (g0 >= nact) return;  // uniform across the CTA: nothing allocated yet
  const int64_t t_begin = (int64_t)blockIdx.x * p.tiles_per_chunk;
  const int64_t t_end = min(t_begin + (int64_t)p.tiles_per_chunk, p.n_tiles);
  const int n_my_tiles = (int)max((int64_t)0, t_end - t_begin);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    mbar_init(b_full, 1);
    for (int a = 0; a < kAcc; ++a) { mbar_init(&tm_full[a], 1); mbar_init(&tm_empty[a], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 512 columns = 2 accumulator stages x 256 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // W2 image of this CTA's chains -> shared memory (epilogue operand)
  {
    const float* src = p.w2img + (int64_t)g0 * W2S;
    const int ng = min(NG, nact - g0);
    for (int i = threadIdx.x; i < NG * W2S; i += kTcThreads) sW2[i] = (i < ng * W2S) ? src[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      mbar_expect_tx(b_full, kBBytes);
      tma_load_2d(&map_b, b_full, sB, 0, blockIdx.y * kUmmaN);
      tma_load_2d(&map_b, b_full, sB + kBBytesPerKb, 64, blockIdx.y * kUmmaN);
      for (int it = 0; it < n_my_tiles; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
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
      // kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kUmmaN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      mbar_wait(b_full, 0);
      const uint32_t b_addr = smem_u32(sB);
      for (int it = 0; it < n_my_tiles; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        const int acc = it % kAcc;
        const uint32_t aph = (it / kAcc) & 1;
        mbar_wait(&tm_empty[acc], aph ^ 1);
        mbar_wait(&full_a[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA + s * kAStageBytes);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kUmmaN);
#pragma unroll
        for (int k = 0; k < kKUsed / kUmmaK; ++k) {
          const uint32_t kb = k >> 2, kin = k & 3;  // 4 K-steps of 32 bytes per 128-byte swizzle row
          const uint64_t ad = umma_desc_sw128(a_addr + kb * kABytesPerKb + kin * 32);
          const uint64_t bd = umma_desc_sw128(b_addr + kb * kBBytesPerKb + kin * 32);
          tc_mma_bf16(d_tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        tc_commit(&empty_a[s]);   // smem slot reusable once these MMAs retire
        tc_commit(&tm_full[acc]); // accumulator ready for the epilogue
      }
    }
  } else {
    // ================= epilogue warps =================
    const int ew = warp - 2;
    const int quarter = warp & 3;            // TMEM lanes this warp may touch: 32*quarter .. +31
    const int cgrp = ew >> 2;                // which block of CPW chains
    const int O = p.O;
    double accL[CPW];
    int accNc[CPW], accSq[CPW];
#pragma unroll
    for (int c = 0; c < CPW; ++c) { accL[c] = 0.0; accNc[c] = 0; accSq[c] = 0; }
    for (int it = 0; it < n_my_tiles; ++it) {
      const int acc = it % kAcc;
      const uint32_t aph = (it / kAcc) & 1;
      const int64_t row = (t_begin + it) * kTileM + quarter * 32 + lane;
      const int y = __ldg(p.y + row);
      mbar_wait(&tm_full[acc], aph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        const int chain = cgrp * CPW + c;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kUmmaN + chain * HP);
        uint32_t zr[HP];
        if constexpr (HP == 32) tmem_ld32(taddr, zr); else tmem_ld16(taddr, zr);
        tmem_ld_wait();
        if (c == CPW - 1) {  // all of this warp's columns are in registers: hand the accumulator back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tm_empty[acc]);
        }
        const float* w2 = sW2 + chain * W2S;
        float out[OP];
#pragma unroll
        for (int o4 = 0; o4 < OP; o4 += 4) {
          const float4 b = *reinterpret_cast<const float4*>(w2 + HP * OP + o4);
          out[o4] = b.x; out[o4 + 1] = b.y; out[o4 + 2] = b.z; out[o4 + 3] = b.w;
        }
#pragma unroll
        for (int h = 0; h < HP; ++h) {
          const float hv = sigmoid_from_neg_log2e(__uint_as_float(zr[h]));
#pragma unroll
          for (int o4 = 0; o4 < OP; o4 += 4) {
            const float4 w = *reinterpret_cast<const float4*>(w2 + h * OP + o4);
            out[o4] = fmaf(hv, w.x, out[o4]);
            out[o4 + 1] = fmaf(hv, w.y, out[o4 + 1]);
            out[o4 + 2] = fmaf(hv, w.z, out[o4 + 2]);
            out[o4 + 3] = fmaf(hv, w.w, out[o4 + 3]);
          }
        }
        float tmin = out[0]; int pm = 0; float sum = 0.f, oy = 0.f;
#pragma unroll
        for (int o = 0; o < OP; ++o) {
          if (o < O) {
            const float t = out[o];
            if (t < tmin) { tmin = t; pm = o; }
            const float s = sigmoid_from_neg_log2e(t);
            sum += ex2_approx(s * kLog2e);
            if (o == y) oy = s;
          }
        }
        if (y >= 0) {
          accL[c] += (double)(oy - kLn2 * lg2_approx(sum));
          accNc[c] += (pm == y);
          const int e = pm - y;
          accSq[c] += e * e;
        }
      }
    }
    // chunk partials: warp shuffle, then the 4 quarter-warps of a chain meet in shared memory
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      const int chain = cgrp * CPW + c;
      const double l = warp_sum(accL[c]);
      const int a = warp_sum(accNc[c]);
      const int b = warp_sum(accSq[c]);
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
  if (threadIdx.x < NG && g0 + (int)threadIdx.x < nact) {
    const double* r = red + threadIdx.x * 12;
    const int64_t o = (int64_t)blockIdx.x * p.R + (g0 + threadIdx.x);
    p.partL[o] = (r[0] + r[3]) + (r[6] + r[9]);
    p.partNc[o] = (int32_t)(r[1] + r[4] + r[7] + r[10]);
    p.partSq[o] = (int32_t)(r[2] + r[5] + r[8] + r[11]);
  }
}


// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// bf16 [rows][inner_elems] row-major, box = 64 elements (one 128-byte swizzle row) x box_rows
int tc_make_map_2d(CUtensorMap* map, void* base, uint64_t inner_elems, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return SAPT_ERR_CUDA; }
  const cuuint64_t dims[2] = {(cuuint64_t)inner_elems, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)inner_elems * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SAPT_ERR_CUDA; }
  return SAPT_OK;
}
static int make_map(CUtensorMap* map, void* base, uint64_t rows, uint32_t box_rows) { return tc_make_map_2d(map, base, kKA, rows, box_rows); }

static inline uint16_t f2bf(float v) {  // round-to-nearest-even, host side
  uint32_t u; memcpy(&u, &v, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float v; memcpy(&v, &u, 4); return v; }

// A image: [n_pad][128] bf16 = [x1|x1|x2|x1|x2|x3|1 1 1 0..|0..]; x_host is the padded fp32 image [n_pad][DP]
int tc_prepare_dataset(sapt_problem* p, int which, const float* x_host, int64_t n) {
  DataSet& ds = p->set[which];
  std::vector<uint16_t> img((size_t)ds.n_pad * kKA, 0);
  const uint16_t one = f2bf(1.0f);
  static const int piece_of_chunk[6] = {0, 0, 1, 0, 1, 2};
  for (int64_t i = 0; i < ds.n_pad; ++i) {
    uint16_t* row = &img[(size_t)i * kKA];
    if (i < n) {
      for (int d = 0; d < p->D; ++d) {
        const float v = x_host[(size_t)i * p->DP + d];
        uint16_t pc[3];
        pc[0] = f2bf(v); const float r1 = v - bf2f(pc[0]);
        pc[1] = f2bf(r1); const float r2 = r1 - bf2f(pc[1]);
        pc[2] = f2bf(r2);
        for (int c = 0; c < 6; ++c) row[c * 16 + d] = pc[piece_of_chunk[c]];
      }
    }
    row[96] = one; row[97] = one; row[98] = one;
  }
  SAPT_CUDA(cudaMalloc(&ds.a_bf16, img.size() * 2));
  SAPT_CUDA(cudaMemcpy(ds.a_bf16, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  ds.tmap = aligned_alloc(64, sizeof(CUtensorMap));
  return make_map(reinterpret_cast<CUtensorMap*>(ds.tmap), ds.a_bf16, (uint64_t)ds.n_pad, kTileM);
}

template <int HP>
static int launch_tc(int OP, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& a, dim3 grid, size_t smem, cudaStream_t st) {
#define SAPT_TC(OPV)                                                                                          \
  case OPV: {                                                                                                 \
    auto k = loglik_tc_kernel<HP, OPV>;                                                                       \
    SAPT_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    k<<<grid, kTcThreads, smem, st>>>(ma, mb, a);                                                             \
    SAPT_CHECK_LAUNCH();                                                                                      \
    return SAPT_OK;                                                                                           \
  }
  switch (OP) {
    SAPT_TC(4) SAPT_TC(8) SAPT_TC(12) SAPT_TC(16) SAPT_TC(20)
    default: set_error("tcgen05 path: O padded to %d unsupported", OP); return SAPT_ERR_UNSUPPORTED;
  }
#undef SAPT_TC
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

int launch_loglik_tc2(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                      double* L, int32_t* nc, double* sq, cudaStream_t st);  // sapt_loglik_tc2.cu

static int tc_version() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SAPT_TC_VERSION"); v = (e && e[0] == '1') ? 1 : 2; }
  return v;
}

int launch_loglik_tc(sapt_problem* p, int which, const float* theta, int64_t ld, int R, const int32_t* rows, const int32_t* n_rows,
                     double* L, int32_t* nc, double* sq, cudaStream_t st) {
  const DataSet& ds = p->set[which];
  if (R <= 0 || ds.n <= 0) return SAPT_OK;
  // v2 (both layers on the tensor cores) is the product path; v1 (layer 2 on the FMA pipe) is kept for A/B runs
  if (tc_version() == 2 && !(p->O > 16 && p->H <= 16)) return launch_loglik_tc2(p, which, theta, ld, R, rows, n_rows, L, nc, sq, st);
  const int HP = p->H <= 16 ? 16 : 32;
  const int OP = p->OP;
  const int NG = kUmmaN / HP;
  const int n_groups = (R + NG - 1) / NG;
  const int64_t n_tiles = ds.n_pad / kTileM;
  // CTAs = chunks x groups, one CTA per SM: make the total a multiple of the SM count when the
  // problem is large enough, keeping >= 16 tiles per chunk to amortise the B panel load
  const int base = p->num_sms / gcd_int(n_groups, p->num_sms);
  int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_tiles / 16, ((int64_t)p->num_sms * 8 + n_groups - 1) / n_groups));
  if (n_chunks >= base) n_chunks = (n_chunks / base) * base;
  n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, n_tiles));
  const int tiles_per_chunk = (int)((n_tiles + n_chunks - 1) / n_chunks);
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;

  const size_t w2s = (size_t)HP * OP + OP;
  const size_t b_bytes = (size_t)n_groups * kUmmaN * kKA * 2;
  const size_t w2_bytes = (size_t)n_groups * NG * w2s * sizeof(float);
  const size_t part_bytes = (size_t)n_chunks * R * (sizeof(double) + 2 * sizeof(int32_t));
  auto up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  int rc = ensure_workspace(p, up(b_bytes) + up(w2_bytes) + up(part_bytes));
  if (rc) return rc;
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->ws);
  __nv_bfloat16* bimg = reinterpret_cast<__nv_bfloat16*>(ws);
  float* w2img = reinterpret_cast<float*>(ws + up(b_bytes));
  double* partL = reinterpret_cast<double*>(ws + up(b_bytes) + up(w2_bytes));
  int32_t* partNc = reinterpret_cast<int32_t*>(partL + (size_t)n_chunks * R);
  int32_t* partSq = partNc + (size_t)n_chunks * R;

  tc_pack_kernel<<<n_groups * NG, 128, 0, st>>>(theta, ld, rows, n_rows, R, p->D, p->H, p->O, HP, OP, NG, bimg, w2img);
  SAPT_CHECK_LAUNCH();

  CUtensorMap mb;
  rc = make_map(&mb, bimg, (uint64_t)n_groups * kUmmaN, kUmmaN);
  if (rc) return rc;
  TcParams a{};
  a.y = ds.y; a.n_pad = ds.n_pad; a.w2img = w2img; a.n_rows = n_rows; a.R = R; a.O = p->O;
  a.tiles_per_chunk = tiles_per_chunk; a.n_tiles = n_tiles; a.partL = partL; a.partNc = partNc; a.partSq = partSq;
  const size_t smem = 1024 + kBBytes + kStages * kAStageBytes + (size_t)NG * w2s * 4 + 8 + (2 * kStages + 1 + 2 * kAcc) * 8 + 16 +
                      (size_t)NG * 12 * 8 + 64;
  dim3 grid((unsigned)n_chunks, (unsigned)n_groups);
  const CUtensorMap& ma = *reinterpret_cast<const CUtensorMap*>(ds.tmap);
  rc = (HP == 16) ? launch_tc<16>(OP, ma, mb, a, grid, smem, st) : launch_tc<32>(OP, ma, mb, a, grid, smem, st);
  if (rc) return rc;
  return launch_loglik_finalize(partL, partNc, partSq, (int)n_chunks, R, rows, n_rows, L, nc, sq, st);
}

}  // namespace sapt
